"""ctypes binding of the C ABI declared in ``include/be_b200.h``.

The shared library ``libbe_b200.so`` is built in-tree by ``__graft_entry__.build()``
(nvcc, sm_100a).  There is no CPU fallback: if the library is missing, or a call is
made without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BE_B200_LIB") or os.path.join(_HERE, "libbe_b200.so")  # override: A/B experiments only
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
]

c_double_p = ctypes.c_void_p  # device pointers travel as integers
c_int_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/be_b200.h one to one
_I, _D, _P, _Z = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t
SIGNATURES = {
    "be_version": (_I, []),
    "be_ctx_create": (_I, [_I, _P, ctypes.POINTER(_P)]),
    "be_ctx_set_stream": (_I, [_P, _P]),
    "be_ctx_destroy": (_I, [_P]),
    "be_ctx_sync": (_I, [_P]),
    "be_ctx_last_error": (ctypes.c_char_p, [_P]),
    "be_ctx_launch_count": (ctypes.c_longlong, [_P]),
    "be_ctx_profile_enable": (_I, [_P, _I]),
    "be_ctx_profile_reset": (_I, [_P]),
    "be_ctx_profile_families": (_I, []),
    "be_ctx_profile_get": (_I, [_P, _I, ctypes.c_char_p, _Z, ctypes.POINTER(_D), ctypes.POINTER(ctypes.c_longlong),
                                ctypes.POINTER(_D), ctypes.POINTER(_D)]),
    "be_gpdtw1d_inputs": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "be_matern32_gram": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "be_potrf_workspace_bytes": (_Z, [_I, _I]),
    "be_potrf_batched": (_I, [_P, _P, _I, _I, _P, _P, _P, _Z]),
    "be_gp_posterior_workspace_bytes": (_Z, [_I, _I, _I]),
    "be_gp_posterior": (_I, [_P, _P, _P, _P, _P, _P, _D, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z]),
    "be_gp_posterior_factored_workspace_bytes": (_Z, [_I, _I, _I]),
    "be_gp_posterior_factored": (_I, [_P, _P, _P, _P, _P, _P, _D, _I, _I, _I, _P, _P, _P, _P, _P, _P, _Z]),
    "be_vgp_fit_workspace_bytes": (_Z, [_I, _I, _I]),
    "be_vgp_fit": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _D, _D, _I, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z]),
    "be_mvn_from_cov_workspace_bytes": (_Z, [_I, _I]),
    "be_mvn_from_cov": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _Z]),
    "be_loglik_weights_mvn": (_I, [_P, _P, _P, _I, _I, _I, _I, _D, _P, _P, _P]),
    "be_mvn_constvec_logprob": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "be_mvn_log_prob": (_I, [_P, _P, _P, _P, _I, _I, _D, _P]),
    "be_normal_logprob": (_I, [_P, _P, _P, _P, _Z, _P]),
    "be_loglik_weights_normal": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _D, _P, _P, _P]),
    "be_weights_time_mean": (_I, [_P, _P, _I, _I, _I, _P]),
    "be_weights_normalise": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "be_barycentre_1d": (_I, [_P, _P, _P, _P, _I, _I, _I, _D, _D, _I, _P, _P, _P]),
    "be_barycentre_1d_partial": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "be_barycentre_1d_finish": (_I, [_P, _P, _I, _I, _D, _D, _I, _P, _P, _P]),
    "be_sqrtm_psd_workspace_bytes": (_Z, [_I, _I]),
    "be_sqrtm_psd": (_I, [_P, _P, _I, _I, _D, _I, _P, _P, _P, _P, _P, _Z]),
    "be_w2_distance_workspace_bytes": (_Z, [_I, _I]),
    "be_w2_distance": (_I, [_P, _P, _P, _P, _P, _I, _I, _D, _I, _P, _P, _P, _Z]),
    "be_w2_distance_diag": (_I, [_P, _P, _P, _P, _P, _I, _I, _P]),
    "be_barycentre_fullcov_workspace_bytes": (_Z, [_I, _I, _I]),
    "be_crps_weights": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "be_ksd_weights": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "be_w2_collapse": (_I, [_P, _P, _I, _I, _I, _P]),
    "be_similarity_weights_pointwise": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "be_dtw_dba_workspace_bytes": (_Z, [_I, _I, _I]),
    "be_dtw_barycenter_averaging_subgradient": (_I, [_P, _P, _I, _I, _I, _I, _D, _D, _D, _P, _P, _P, _P, _P, _Z]),
    "be_perform_dba": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _Z]),
    "be_dtw_squared": (_I, [_P, _P, _P, _I, _I, _P]),
    "be_svgp_fit_workspace_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "be_svgp_fit": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _I, _D, _D, _I, _D, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z]),
    "be_barycentre_fullcov": (_I, [_P, _P, _P, _P, _I, _I, _I, _D, _D, _I, _D, _I, _P, _P, _P, _P, _P, _Z]),
}

BE_OK = 0
BE_ERR_CUDA = 1000
BE_ERR_WORKSPACE = 1001
BE_ERR_UNSUPPORTED = 1002


class BackendError(RuntimeError):
    """The CUDA library is missing or a call failed at the CUDA level."""


def build_library(verbose: bool = False) -> str:
    """Compiles csrc/be_api.cu into libbe_b200.so for sm_100a (cross-compiles without a GPU)."""
    src = os.path.join(CSRC, "be_api.cu")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(INCLUDE, "be_b200.h"))
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-o", LIB_PATH, src]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BackendError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "bayesian_ensembling_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx_handle, rc: int, what: str):
    if rc == BE_OK:
        return
    if rc < 0:
        raise ValueError(f"{what}: bad argument #{-rc}")
    lib = load_library()
    msg = lib.be_ctx_last_error(ctx_handle).decode() if ctx_handle else ""
    if rc == BE_ERR_WORKSPACE:
        raise BackendError(f"{what}: workspace too small")
    raise BackendError(f"{what}: CUDA failure ({rc}) {msg}")
