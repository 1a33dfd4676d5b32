"""Error behaviour of the C ABI on a device (SURVEY 8b "Errors"): 0 = ok, < 0 = index of the bad argument
(-> ValueError in the mirror), BE_ERR_WORKSPACE / BE_ERR_UNSUPPORTED (-> BackendError); numerical conditions
are NOT errors: per-problem ``info`` and NaN propagation.  Needs a B200: ``-m gpu``."""
import ctypes

import numpy as np
import pytest

from bayesian_ensembling_b200 import _lib

pytestmark = pytest.mark.gpu


def _t(backend, a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=backend.device)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def test_negative_codes_name_the_bad_argument(backend):
    import torch

    lib, ctx = backend.lib, backend.ctx
    x = torch.zeros(2, 3, 5, dtype=torch.float64, device=backend.device)
    out = torch.zeros(2, 5, dtype=torch.float64, device=backend.device)
    assert lib.be_gpdtw1d_inputs(None, _p(x), 2, 3, 5, None, _p(out), _p(out)) == -1
    assert lib.be_gpdtw1d_inputs(ctx, None, 2, 3, 5, None, _p(out), _p(out)) == -2
    assert lib.be_dtw_squared(ctx, _p(out), _p(out), 0, 5, _p(out)) == -4
    assert lib.be_dtw_squared(ctx, _p(out), _p(out), 2, 5, None) == -6
    assert lib.be_ksd_weights(ctx, _p(x), _p(x), _p(x), 2, 0, 3, 5, _p(x), None) == -6
    with pytest.raises(ValueError, match="bad argument #4"):
        _lib.check(ctx, -4, "be_dtw_squared")


def test_workspace_too_small_and_unsupported_shapes(backend):
    import torch

    lib, ctx = backend.lib, backend.ctx
    B, R, T = 2, 3, 40
    x = torch.zeros(B, R, T, dtype=torch.float64, device=backend.device)
    bary = torch.zeros(B, T, dtype=torch.float64, device=backend.device)
    ws = torch.zeros(64, dtype=torch.uint8, device=backend.device)
    rc = lib.be_dtw_barycenter_averaging_subgradient(ctx, _p(x), B, R, T, 5, 0.05, 0.005, 1e-3, None, _p(bary), None, None,
                                                     _p(ws), 64)
    assert rc == _lib.BE_ERR_WORKSPACE
    with pytest.raises(_lib.BackendError, match="workspace"):
        _lib.check(ctx, rc, "dba")
    assert lib.be_dtw_dba_workspace_bytes(1, 2, 4097) == 0  # T > 4096: no kernel shape
    rc = lib.be_dtw_squared(ctx, _p(bary), _p(bary), 1, 5000, _p(bary))
    assert rc == _lib.BE_ERR_UNSUPPORTED
    # more than 50 series: the reference's medoid search turns random (dtwa.py:26) -> argument error, not a guess
    assert lib.be_perform_dba(ctx, _p(x), 1, 51, 4, 1, _p(bary), None, _p(ws), 64) == -4
    post_ws = lib.be_gp_posterior_factored_workspace_bytes(B, T, R)
    assert post_ws > lib.be_gp_posterior_workspace_bytes(B, T, R) > 0


def test_numerical_conditions_are_info_not_errors(backend):
    """A member whose realisations are all identical has y_var = 0: M = K + 1e-6 I with K rank one is still
    positive definite, the posterior is finite; a NaN input poisons that problem only."""
    rng = np.random.default_rng(0)
    M, R, T = 3, 4, 30
    reals = rng.normal(size=(M, R, T)).cumsum(axis=2) * 0.1
    reals[1, :, :] = reals[1, :1, :]          # zero across-realisation variance
    reals[2, 0, 7] = np.nan                   # NaN input
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    post = backend.gp_posterior(X, ym, yv, np.full(M, 0.5), np.full(M, 6.0), want_cov=False, want_scale_tri=False)
    mu = post.mu.cpu().numpy()
    assert np.isfinite(mu[0]).all()
    assert int(post.info_fit[0]) == 0 and int(post.info_dist[0]) == 0
    assert np.isnan(mu[2]).any()              # propagates, as jnp.linalg.cholesky would
    assert np.isfinite(mu[0]).all()           # and stays inside its own problem
    fac = backend.gp_posterior_factored(X, ym, yv, np.full(M, 0.5), np.full(M, 6.0))
    assert np.isfinite(fac.mu[0].cpu().numpy()).all() and np.isnan(fac.mu[2].cpu().numpy()).any()
