"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Reader for the reference's pickled ``ModelCollection`` fixtures
(``/root/reference/experiments/pre_fit_models/*.pkl``, written by
``ensembles/data.py:397-404``) that needs neither xarray nor distrax.

The pickles hold xarray / pandas / distrax objects; every global that is not
numpy or a builtin is replaced by a stub that only remembers its state, which
is enough to dig out the fp64 arrays we need:

* ``model_data`` realisations ``[R, T]``          (input of ``GPDTW1D.fit``, models.py:175)
* ``mu [T]``, ``covariance [T, T]``               (output of ``GPDTW1D.fit``, models.py:217-229)
* ``_scale_tri [T, T]``                           (distrax Cholesky factor, data.py:38-39)
"""
from __future__ import annotations

import io
import pickle
from dataclasses import dataclass

import numpy as np


class _Stub:
    """Stands in for any class the pickle references that is not importable here."""

    def __init__(self, *a, **k):
        self._args = a
        self._kwargs = k
        self._state = None

    def __setstate__(self, state):
        self._state = state

    def __call__(self, *a, **k):  # some reconstructors are called again
        return _Stub(*a, **k)

    def __reduce_ex__(self, proto):  # pragma: no cover - never re-pickled
        raise TypeError("stub objects are read-only")


def _make_stub_class(module: str, name: str):
    return type(name, (_Stub,), {"__module__": module})


def _make_stub_function(module: str, name: str):
    def _f(*a, **k):
        s = _Stub(*a, **k)
        s._func = f"{module}.{name}"
        return s

    return _f


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("numpy"):
            module = module.replace("numpy.core", "numpy._core")
            return super().find_class(module, name)
        if module in ("builtins", "collections", "copyreg", "datetime"):
            return super().find_class(module, name)
        if name[:1].isupper():
            return _make_stub_class(module, name)
        return _make_stub_function(module, name)


@dataclass
class FittedMember:
    name: str
    realisations: np.ndarray  # [R, T]
    mu: np.ndarray  # [T]
    covariance: np.ndarray  # [T, T]
    scale_tri: np.ndarray  # [T, T]


def _find_arrays(obj, want_ndim, depth=0, seen=None):
    """Depth-first search for the first float64 ndarray with ``want_ndim`` dims."""
    if seen is None:
        seen = set()
    if id(obj) in seen or depth > 12:
        return None
    seen.add(id(obj))
    if isinstance(obj, np.ndarray):
        if obj.dtype == np.float64 and obj.ndim == want_ndim:
            return obj
        return None
    if isinstance(obj, dict):
        it = obj.values()
    elif isinstance(obj, (list, tuple)):
        it = obj
    elif isinstance(obj, _Stub):
        it = [getattr(obj, '_state', None), getattr(obj, '_args', None), getattr(obj, '_kwargs', None)]
    else:
        return None
    for v in it:
        r = _find_arrays(v, want_ndim, depth + 1, seen)
        if r is not None:
            return r
    return None


def _state_dict(stub):
    st = getattr(stub, '_state', None)
    if isinstance(st, tuple):  # (dict, slots) form
        for s in st:
            if isinstance(s, dict):
                return s
    return st


def load_fitted_collection(path: str) -> list[FittedMember]:
    with open(path, "rb") as f:
        mc = _StubUnpickler(io.BytesIO(f.read())).load()
    models = _state_dict(mc)["models"]
    out = []
    for pm in models:
        st = _state_dict(pm)
        reals = _find_arrays(st["model_data"], 2)
        dist = _state_dict(st["_distribution"])
        mu = np.asarray(dist["mu"], dtype=np.float64)
        cov = np.asarray(dist["covariance"], dtype=np.float64)
        inner = _state_dict(dist["_dist"])
        tri = np.asarray(inner["_scale_tri"], dtype=np.float64)
        out.append(FittedMember(str(st["model_name"]), np.array(reals), mu, cov, tri))
    return out
