"""The C-ABI library: builds for sm_100a without a GPU, exports exactly what include/be_b200.h
declares, and the Python binding table covers it.  No compute calls (CPU only)."""
import ctypes
import os
import re
import subprocess

import pytest

from bayesian_ensembling_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "be_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(be_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    _lib.build_library()
    return _lib.load_library()


def test_header_compiles_as_c():
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in be_b200.h but not exported by libbe_b200.so"


def test_binding_table_matches_header(lib):
    assert sorted(_lib.SIGNATURES) == _declared()


def test_no_torch_or_cuda_types_in_signatures():
    src = open(HEADER).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "cudaStream_t" not in code and "torch" not in code and "at::" not in code
    assert 'extern "C"' in code


def test_version_and_argument_checking_without_gpu(lib):
    assert lib.be_version() >= 100
    # NULL ctx is argument #1 for every entry point: rejected before any CUDA call
    assert lib.be_ctx_sync(None) == -1
    assert lib.be_barycentre_1d(None, None, None, None, 1, 1, 1, 1e-6, 1.0, 200, None, None, None) == -1
    assert lib.be_ctx_launch_count(None) == 0


def test_sass_is_sm100a_with_fp64_tensor_instructions():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN2be13k_chol_updateEPdiiii", _lib.LIB_PATH],
                          capture_output=True, text=True).stdout
    if "DMMA" not in sass:  # symbol name may change with the signature: scan everything
        sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "DMMA" in sass, "FP64 tensor-core instructions missing from the factorisation kernels"
    assert "LDGSTS" in sass, "cp.async staging missing"


def test_product_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from bayesian_ensembling_b200.backend import Backend

    with pytest.raises(_lib.BackendError):
        Backend.get()
    from bayesian_ensembling_b200 import gaussian_barycentre

    with pytest.raises(_lib.BackendError):
        gaussian_barycentre([0.0], [1.0], [1.0])
