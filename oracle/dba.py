"""ctypes wrapper of oracle/dba.c -- the CPU oracle of the DTW-barycentre-averaging step.
TEST INFRASTRUCTURE ONLY (see the header of oracle/dba.c for what is pinned and what is not).

``build()`` compiles the C restatement with gcc into ``oracle/_build/libdba_oracle.so``
(git-ignored, travels to the GPU box); ``__graft_entry__.build()`` calls it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "dba.c")
LIB = os.path.join(_HERE, "_build", "libdba_oracle.so")
TIE_TSLEARN, TIE_DTWA = 0, 1

_lib = None


def build(verbose: bool = False) -> str:
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        P, I, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        lib.be_oracle_squared_dtw.restype = D
        lib.be_oracle_squared_dtw.argtypes = [P, I, P, I, I]
        lib.be_oracle_dtw_path.restype = I
        lib.be_oracle_dtw_path.argtypes = [P, I, P, I, I, P, P]
        lib.be_oracle_dba_subgradient.restype = I
        lib.be_oracle_dba_subgradient.argtypes = [P, I, I, I, D, D, D, P, P, P]
        lib.be_oracle_perform_dba.restype = I
        lib.be_oracle_perform_dba.argtypes = [P, I, I, I, P]
        _lib = lib
    return _lib


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def squared_dtw(s, t, tie=TIE_DTWA) -> float:
    """ensembles/dtwa.py:48-75 ``squared_DTW`` (tie rule irrelevant for the value)."""
    s, ps = _c(s)
    t, pt = _c(t)
    return float(_load().be_oracle_squared_dtw(ps, s.size, pt, t.size, tie))


def dtw_path(s, t, tie=TIE_TSLEARN):
    """tslearn ``dtw_path(s, t)`` -> (list of (i, j), sqrt of the accumulated cost)."""
    s, ps = _c(s)
    t, pt = _c(t)
    path = np.zeros(2 * (s.size + t.size), dtype=np.int32)
    cost = ctypes.c_double()
    n = _load().be_oracle_dtw_path(ps, s.size, pt, t.size, tie, path.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.cast(ctypes.byref(cost), ctypes.c_void_p))
    return path[: 2 * n].reshape(n, 2), float(np.sqrt(cost.value))


def dba_subgradient(X, max_iter=30, initial_step_size=0.05, final_step_size=0.005, tol=1e-5, init_barycenter=None):
    """tslearn 0.5.1.0 ``dtw_barycenter_averaging_subgradient`` for X [R, T]; the reference calls it
    with max_iter=50, tol=1e-3 (ensembles/models.py:176-178).  -> (barycentre [T], n_iter, last cost)."""
    X, px = _c(X)
    R, T = X.shape
    out = np.empty(T)
    cost = ctypes.c_double()
    pi = None
    if init_barycenter is not None:
        init, pi = _c(np.asarray(init_barycenter).ravel())
    n = _load().be_oracle_dba_subgradient(px, R, T, int(max_iter), float(initial_step_size), float(final_step_size),
                                          float(tol), pi, out.ctypes.data_as(ctypes.c_void_p),
                                          ctypes.cast(ctypes.byref(cost), ctypes.c_void_p))
    return out, int(n), float(cost.value)


def perform_dba(X, n_iterations=10):
    """ensembles/dtwa.py:6-20 ``performDBA`` for <= 50 equal-length series -> (centre [T], medoid index)."""
    X, px = _c(X)
    R, T = X.shape
    assert R <= 50
    out = np.empty(T)
    med = _load().be_oracle_perform_dba(px, R, T, int(n_iterations), out.ctypes.data_as(ctypes.c_void_p))
    return out, int(med)
