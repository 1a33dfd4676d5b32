"""The fp64-equivalent GEMM on the int8 tensor cores (be_dgemm_nt_i8tc: tcgen05.mma kind::i8, Ozaki scheme) against
fp64 references.  Needs a B200: ``-m gpu``.  The error measure is |C - C_ref| / sum_k |a_ik b_jk| -- the scale of an
fp64 dot product's own rounding bound -- with C_ref from NumPy's longdouble (80-bit) accumulation."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]  # mbarrier pipelines: a regression must fail, not hang


def _t(backend, a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=backend.device)


def _operands(rng, M, N, K, spread):
    A = rng.uniform(-0.5, 0.5, size=(M, K)) * np.exp(spread * rng.uniform(-0.5, 0.5, size=(M, K)))
    B = rng.uniform(-0.5, 0.5, size=(N, K)) * np.exp(spread * rng.uniform(-0.5, 0.5, size=(N, K)))
    A *= np.ldexp(1.0, (np.arange(M) * 7) % 40 - 20)[:, None]   # rows of very different scale
    B *= np.ldexp(1.0, (np.arange(N) * 5) % 30 - 15)[:, None]
    return A, B


@pytest.mark.parametrize("M,N,K,spread", [(128, 256, 32, 0.0), (128, 256, 64, 4.0), (256, 512, 1024, 4.0), (384, 256, 3008, 6.0)])
def test_dgemm_nt_i8tc_matches_long_double(backend, M, N, K, spread):
    rng = np.random.default_rng(M + N + K)
    A, B = _operands(rng, M, N, K, spread)
    C = backend.dgemm_nt_i8tc(_t(backend, A), _t(backend, B)).cpu().numpy()
    rows = rng.choice(M, size=min(M, 24), replace=False)
    Al = A[rows].astype(np.longdouble)
    ref = Al @ B.astype(np.longdouble).T
    mag = np.abs(Al) @ np.abs(B.astype(np.longdouble)).T
    err = float((np.abs(C[rows].astype(np.longdouble) - ref) / mag).max())
    plain = float((np.abs((A[rows] @ B.T).astype(np.longdouble) - ref) / mag).max())
    assert err < 1e-15, (err, plain)


def test_dgemm_nt_i8tc_special_values_and_errors(backend):
    rng = np.random.default_rng(3)
    A, B = _operands(rng, 128, 256, 96, 2.0)
    A[5] = 0.0                      # an all-zero row: exponent 0, slices 0
    B[7, :] = 0.0
    A[9, :] = 2.0 ** -300           # tiny but representable rows keep their relative accuracy
    C = backend.dgemm_nt_i8tc(_t(backend, A), _t(backend, B)).cpu().numpy()
    assert np.all(C[5] == 0.0) and np.all(C[:, 7] == 0.0)
    ref = A[9].astype(np.longdouble) @ B.astype(np.longdouble).T
    mag = np.abs(A[9]).astype(np.longdouble) @ np.abs(B).astype(np.longdouble).T
    ok = mag > 0
    assert float((np.abs(C[9].astype(np.longdouble) - ref)[ok] / mag[ok]).max()) < 1e-15
    import torch

    with pytest.raises(ValueError):
        backend.dgemm_nt_i8tc(torch.zeros(100, 32, dtype=torch.float64), torch.zeros(256, 32, dtype=torch.float64))
    with pytest.raises(ValueError):
        backend.dgemm_nt_i8tc(torch.zeros(128, 32, dtype=torch.float64), torch.zeros(256, 64, dtype=torch.float64))
