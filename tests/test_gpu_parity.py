"""Parity of the CUDA path (through the C ABI) against the CPU oracle, the reference's golden
fixtures, and size-independent properties at BASELINE sizes.  Needs a B200: ``-m gpu``.

Tolerances (BASELINE.json north_star): <= 1e-8 relative on posterior mean / variance,
<= 1e-6 on normalised weights and barycentre moments.  Most checks are far tighter and say so.
"""
import warnings

import numpy as np
import pytest

from bayesian_ensembling_b200 import synthetic
from oracle import reference_path as rp
from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL_POSTERIOR = 1e-8
TOL_WEIGHTS = 1e-6


def _t(backend, a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=backend.device)


def _cell(M, R, T, Ro, seed, monthly=False):
    cfg = synthetic.Config("t", 9, 1, M, R, T, Ro, monthly, "")
    reals, obs = synthetic.make_cells(cfg, seed=seed)
    return reals[0], obs[0]


def _nan_equal_close(got, want, tol, name=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    assert (np.isnan(got) == np.isnan(want)).all(), f"{name}: NaN pattern differs"
    ok = ~np.isnan(want)
    if ok.any():
        err = np.abs(got[ok] - want[ok]).max() / max(np.abs(want[ok]).max(), 1e-300)
        assert err <= tol, (name, err)


# ------------------------------------------------------------------------------------ a1 inputs / gram
@pytest.mark.parametrize("R,T", [(1, 5), (3, 24), (5, 251), (25, 165)])
def test_gpdtw1d_inputs(backend, R, T):
    reals, _ = _cell(3, R, T, 2, seed=R * 100 + T)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    for m in range(3):
        Xo, yo, so = rp.gpdtw1d_inputs(reals[m])
        assert np.array_equal(X[m].cpu().numpy(), Xo)
        assert rel_err(ym[m].cpu().numpy(), yo) < 1e-15
        assert np.abs(yv[m].cpu().numpy() - so).max() < 1e-17 + 1e-14 * so.max()


@pytest.mark.parametrize("R,T", [(1, 7), (3, 24), (5, 130), (10, 251), (25, 300)])
def test_matern32_gram(backend, R, T):
    reals, _ = _cell(2, R, T, 2, seed=T)
    X, _, _ = backend.gpdtw1d_inputs(_t(backend, reals))
    var = np.array([0.5, 1.3])
    ls = np.array([6.0, 0.9])
    K = backend.matern32_gram(X, var, ls).cpu().numpy()
    for m in range(2):
        Xo, _, _ = rp.gpdtw1d_inputs(reals[m])
        assert rel_err(K[m], rp.matern32_gram(Xo, var[m], ls[m])) < 1e-12
        # (-2 x.y + |x|^2) + |y|^2 is not symmetric in rounding (neither is GPflow's): 1 ulp of r2
        assert np.abs(K[m] - K[m].T).max() < 1e-13


# ------------------------------------------------------------------------------------ a3 Cholesky
@pytest.mark.parametrize("T", [1, 2, 15, 16, 17, 86, 127, 128, 129, 165, 251, 300, 515])
def test_potrf_vs_lapack(backend, T):
    rng = np.random.default_rng(T)
    B = 3
    A = rng.normal(size=(B, T, T + 3))
    A = A @ A.transpose(0, 2, 1) / T + 0.5 * np.eye(T)
    L, info = backend.potrf(_t(backend, A))
    assert info.cpu().numpy().tolist() == [0] * B
    L = L.cpu().numpy()
    for b in range(B):
        assert rel_err(L[b], np.linalg.cholesky(A[b])) < 1e-12
        assert np.abs(np.triu(L[b], 1)).max() == 0.0


@pytest.mark.parametrize("T", [16, 100, 128, 165, 251, 300])
def test_potrf_many_problems_uses_the_two_cta_diag_kernel(backend, T):
    """With >= 2 x (SM count) problems the diagonal-block kernel runs in its 100 KB form (trapezoidal
    storage, in-place recursive-doubling inverse), two CTAs per SM: same factors, same inverse (through
    the posterior), same non-PD reports."""
    rng = np.random.default_rng(T + 7)
    B = 320
    A = rng.normal(size=(B, T, T + 3))
    A = A @ A.transpose(0, 2, 1) / T + 0.5 * np.eye(T)
    bad = 5
    if T >= 16:
        A[bad, 9, 9] = -1.0  # leading minor of order 10 is not positive definite
    L, info = backend.potrf(_t(backend, A))
    info = info.cpu().numpy()
    assert info[bad] == 10 and np.count_nonzero(info) == 1
    L = L.cpu().numpy()
    ok = np.arange(B) != bad
    assert rel_err(L[ok], np.linalg.cholesky(A[ok])) < 1e-12
    assert np.abs(np.triu(L[ok], 1)).max() == 0.0
    # the inverse of the diagonal blocks enters trtri / the posterior: a batch of B member posteriors
    reals, _ = _cell(4, 3, T, 2, seed=T)
    rb = np.tile(reals, (B // 4, 1, 1))
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, rb))
    post = backend.gp_posterior(X, ym, yv, np.full(B, 0.5), np.full(B, 6.0), want_scale_tri=False)
    for m in range(4):
        Xo, yo, so = rp.gpdtw1d_inputs(reals[m])
        mu_o, cov_o = rp.gp_posterior_closed_form(Xo, yo, so, 0.5, 6.0)
        for k in (m, B - 4 + m):
            assert rel_err(post.mu[k].cpu().numpy(), mu_o) <= TOL_POSTERIOR
            assert rel_err(post.cov[k].cpu().numpy(), cov_o) <= TOL_POSTERIOR


def test_potrf_reports_non_pd_like_lapack(backend):
    T = 200
    rng = np.random.default_rng(0)
    A = rng.normal(size=(2, T, T))
    A = A @ A.transpose(0, 2, 1) + T * np.eye(T)
    A[1, 140, 140] = -1.0  # leading minor 141 is not PD
    _, info = backend.potrf(_t(backend, A))
    assert info.cpu().numpy().tolist() == [0, 141]


@pytest.mark.parametrize("T,R,M", [(3, 2, 2), (24, 3, 5), (126, 4, 2), (165, 10, 4), (251, 3, 3), (300, 5, 2)])
def test_gp_posterior_factored_vs_dense_and_oracle(backend, T, R, M):
    """be_gp_posterior_factored (Woodbury form, no dense covariance) returns the same mean, variance and
    constant-vector statistics as the dense path and the oracle."""
    reals, obs = _cell(M, R, T, 4, seed=T * 3 + R)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    var, ls = np.full(M, 0.5), np.linspace(4.0, 8.0, M)
    dense = backend.gp_posterior(X, ym, yv, var, ls, want_cov=False, want_scale_tri=False)
    fac = backend.gp_posterior_factored(X, ym, yv, var, ls)
    assert int(fac.info_fit.abs().sum()) == 0 and int(fac.info_dist.abs().sum()) == 0
    assert rel_err(fac.mu.cpu().numpy(), dense.mu.cpu().numpy()) <= 1e-13
    assert rel_err(fac.var_diag.cpu().numpy(), dense.var_diag.cpu().numpy()) <= 1e-10
    sd, sf = dense.mvn_stats.cpu().numpy(), fac.mvn_stats.cpu().numpy()
    assert np.abs(sf - sd).max() <= 1e-9 * np.abs(sd).max()
    ob = _t(backend, obs[None])
    wd = backend.loglik_weights_mvn(dense.mvn_stats, ob, M).cpu().numpy()
    wf = backend.loglik_weights_mvn(fac.mvn_stats, ob, M).cpu().numpy()
    _nan_equal_close(wf, wd, TOL_WEIGHTS, "weights from factored statistics")
    for m in range(M):
        Xo, yo, so = rp.gpdtw1d_inputs(reals[m])
        mu_o, cov_o = rp.gp_posterior_closed_form(Xo, yo, so, var[m], ls[m])
        assert rel_err(fac.mu[m].cpu().numpy(), mu_o) <= TOL_POSTERIOR
        assert rel_err(fac.var_diag[m].cpu().numpy(), np.diag(cov_o)) <= TOL_POSTERIOR


def test_grid_factored_posterior_equals_dense(backend):
    from bayesian_ensembling_b200 import grid

    cfg = synthetic.Config("t", 9, 3, 4, 3, 140, 3, False, "")
    reals, obs = synthetic.make_cells(cfg, seed=5)
    a = grid.fit_weight_barycentre(reals, obs, 0.5, 6.0)
    b = grid.fit_weight_barycentre(reals, obs, 0.5, 6.0, posterior="factored")
    _nan_equal_close(b.weights.cpu().numpy(), a.weights.cpu().numpy(), TOL_WEIGHTS, "weights")
    _nan_equal_close(b.bary_mu.cpu().numpy(), a.bary_mu.cpu().numpy(), TOL_WEIGHTS, "bary mu")
    _nan_equal_close(b.bary_std.cpu().numpy(), a.bary_std.cpu().numpy(), TOL_WEIGHTS, "bary std")
    assert rel_err(b.var_diag.cpu().numpy(), a.var_diag.cpu().numpy()) <= 1e-10
    with pytest.raises(ValueError):
        grid.fit_weight_barycentre(reals, obs, 0.5, 6.0, posterior="factored", keep_posteriors=True)


def test_golden_scale_tri(backend, golden_members):
    """a3 against the reference's stored distrax factor (pinned parity)."""
    for T in (86, 165):
        ms = [m for m in golden_members if m.mu.shape[0] == T]
        cov = np.stack([m.cov for m in ms])
        mu = np.stack([m.mu for m in ms])
        tri, var_diag, stats, info = backend.mvn_from_cov(_t(backend, mu), _t(backend, cov))
        assert int(info.abs().sum()) == 0
        tri = tri.cpu().numpy()
        for k, m in enumerate(ms):
            assert np.abs(tri[k] - m.scale_tri).max() <= 1e-13, m.key
            assert np.array_equal(var_diag[k].cpu().numpy(), np.diag(m.cov))


# ------------------------------------------------------------------------------------ a1 posterior
@pytest.mark.parametrize("T,R,M", [(3, 2, 2), (24, 3, 5), (86, 5, 3), (126, 4, 2), (128, 4, 2), (165, 10, 4),
                                   (251, 3, 10), (300, 5, 3), (600, 5, 2)])
def test_gp_posterior_vs_oracle(backend, T, R, M):
    reals, _ = _cell(M, R, T, 2, seed=T + R)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    var = np.linspace(0.3, 1.1, M)
    ls = np.linspace(4.0, 8.0, M)
    post = backend.gp_posterior(X, ym, yv, var, ls)
    assert int(post.info_fit.abs().sum()) == 0 and int(post.info_dist.abs().sum()) == 0
    for m in range(M):
        Xo, yo, so = rp.gpdtw1d_inputs(reals[m])
        mean, cov = rp.gp_posterior_closed_form(Xo, yo, so, var[m], ls[m])
        assert rel_err(post.mu[m].cpu().numpy(), mean) <= TOL_POSTERIOR
        assert np.abs(post.var_diag[m].cpu().numpy() / np.diag(cov) - 1).max() <= TOL_POSTERIOR
        assert rel_err(post.cov[m].cpu().numpy(), cov) <= TOL_POSTERIOR
        assert rel_err(post.scale_tri[m].cpu().numpy(), np.linalg.cholesky(cov)) <= TOL_POSTERIOR
        c = post.cov[m].cpu().numpy()
        assert np.array_equal(c, c.T)


def test_gp_posterior_on_golden_inputs(backend, golden_members):
    """The reference's own inputs with the hyper-parameters recovered from its fits: the
    CUDA posterior reproduces the reference's stored covariance to the structural-pin level
    (~1e-6 abs) and the oracle to 1e-8."""
    for T in (86, 165):
        for m in [g for g in golden_members if g.mu.shape[0] == T]:
            X, ym, yv = backend.gpdtw1d_inputs(_t(backend, m.realisations[None]))
            post = backend.gp_posterior(X, ym, yv, [m.variance], [m.lengthscale])
            cov = post.cov[0].cpu().numpy()
            assert np.abs(cov - m.cov).max() <= 6e-6, m.key
            Xo, yo, so = rp.gpdtw1d_inputs(m.realisations)
            mean_o, cov_o = rp.gp_posterior_closed_form(Xo, yo, so, m.variance, m.lengthscale)
            assert rel_err(cov, cov_o) <= TOL_POSTERIOR and rel_err(post.mu[0].cpu().numpy(), mean_o) <= TOL_POSTERIOR


# ------------------------------------------------------------------------------------ a4 weights
def test_constvec_logprob_and_weights_on_golden(backend, golden_members):
    """Reference posteriors + leave-one-out pseudo-observations (utils.py:196-200): log-probs,
    exp underflow and the 0/0 -> NaN columns (Q-EXP) must match the oracle exactly in pattern."""
    group = [m for m in golden_members if m.tag == "histssp460"]
    obs = group[2].realisations  # [10,165]
    members = group[:2]
    mus = np.stack([m.mu for m in members])
    covs = np.stack([m.cov for m in members])
    tri, _, stats, _ = backend.mvn_from_cov(_t(backend, mus), _t(backend, covs))
    w_o, e_o, l_o = rp.loglik_weights_mvn(mus, tri.cpu().numpy(), obs)
    ll = backend.mvn_constvec_logprob(stats, _t(backend, obs[None]), 2)[0].cpu().numpy()  # [M,Ro,T]
    for m in range(2):
        for r in range(obs.shape[0]):
            want = rp.mvn_log_prob(mus[m], members[m].scale_tri, obs[r][:, None])
            assert rel_err(ll[m, r], want) < 1e-10
    w, le, lm = backend.loglik_weights_mvn(stats, _t(backend, obs[None]), 2, want_lls=True)
    _nan_equal_close(lm[0].cpu().numpy(), l_o, 1e-10, "lls_mean")
    _nan_equal_close(le[0].cpu().numpy(), e_o, 1e-7, "lls_exp")
    _nan_equal_close(w[0].cpu().numpy(), w_o, TOL_WEIGHTS, "weights")
    assert np.isnan(w_o).any()


@pytest.mark.parametrize("M,Ro", [(2, 1), (5, 2), (10, 5), (10, 10)])
def test_weights_reference_test_shapes(backend, M, Ro):
    """Shapes of the reference's tests/test_weights.py:71-101 (24 monthly steps)."""
    reals, obs = _cell(M, 3, 24, Ro, seed=M * 10 + Ro, monthly=True)
    o = rp.cell_pipeline_L1(reals, obs, 0.5, 6.0)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    post = backend.gp_posterior(X, ym, yv, np.full(M, 0.5), np.full(M, 6.0))
    w = backend.loglik_weights_mvn(post.mvn_stats, _t(backend, obs[None]), M)[0].cpu().numpy()
    assert w.shape == (M, 24)
    _nan_equal_close(w, o["weights"], TOL_WEIGHTS, "weights")
    ok = ~np.isnan(w).any(axis=0)
    assert np.allclose(w[:, ok].sum(axis=0), 1.0, atol=1e-6)


def test_normal_branch_moment_form_and_special_scales(backend):
    """k_loglik_weights_normal takes the mean log-density from the observations' mean and centred second moment;
    zero / negative / subnormal / infinite / NaN scales keep the per-realisation form: all against the oracle,
    including observations tightly clustered far from zero with the model mean inside the cluster."""
    rng = np.random.default_rng(15)
    C, M, Ro, N = 1, 9, 10, 80
    loc = 288.0 + 1e-3 * rng.normal(size=(C, M, N))
    obs = 288.0 + 1e-3 * rng.normal(size=(C, Ro, N))
    scale = rng.uniform(5e-4, 5e-3, size=(C, M, N))
    scale[0, 0, :6] = [0.0, -0.2, 1e-320, np.inf, np.nan, 1e-200]
    w, le, lm = backend.loglik_weights_normal(_t(backend, loc), _t(backend, scale), _t(backend, obs), want_lls=True)
    with np.errstate(all="ignore"):
        wo, eo, lo = rp.loglik_weights_normal(loc[0], scale[0], obs[0])
    got = lm[0].cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(lo))
    inf = np.isinf(lo)
    assert np.array_equal(np.isinf(got), inf) and np.array_equal(got[inf], lo[inf])
    fin = np.isfinite(lo)
    assert np.abs(got[fin] - lo[fin]).max() <= 1e-12 * np.abs(lo[fin]).max()
    _nan_equal_close(w[0].cpu().numpy(), wo, 1e-9, "weights")  # exp amplifies the 1e-12 of exponents up to ~1e3


def _stats_for_exponent(backend, x, T):
    """mvn statistics (|a|^2, a.b, |b|^2, sum log diag L) under which a zero observation's mean log-density is x"""
    import torch

    st = torch.zeros(len(x), 4, dtype=torch.float64, device=backend.device)
    st[:, 3] = -torch.as_tensor(x, dtype=torch.float64, device=backend.device) - 0.5 * T * np.log(2 * np.pi)
    return st


def test_weights_exponential_whole_range(backend):
    """exp_tab16 of k_loglik_weights_mvn_tab (be_kernels.cuh) against the oracle's np.exp over the exponent
    range of the fast path and beyond it (overflow, gradual underflow), for M % 4 != 0 as well."""
    import torch

    rng = np.random.default_rng(11)
    # (M, T): rows shorter and longer than the CTA (128 threads to M = 32, 64 beyond), M % 4 != 0, and an M whose
    # staging does not fit shared memory (the library-exp kernel with the output array as staging)
    for M, T in ((24, 96), (7, 96), (40, 300), (33, 50), (24, 129), (200, 70)):
        C = 16
        x = rng.uniform(-760.0, 720.0, size=C * M)
        x[:6] = [0.0, 709.7, -745.0, 699.999999, -700.0, 1e-9]
        obs = torch.zeros(C, 2, T, dtype=torch.float64, device=backend.device)
        w, le, lm = backend.loglik_weights_mvn(_stats_for_exponent(backend, x, T), obs, M, want_lls=True)
        lm, le, w = lm.cpu().numpy(), le.cpu().numpy(), w.cpu().numpy()
        with np.errstate(over="ignore", under="ignore", invalid="ignore"):
            want = np.exp(lm)
            pos = np.isfinite(want) & (want > 0)
            assert (np.abs(le - want)[pos] <= 4e-16 * want[pos] + 1e-323).all()
            assert np.array_equal(le[~pos], want[~pos])  # overflow -> inf, underflow -> 0
            w_o = le / le.sum(axis=1, keepdims=True)
        _nan_equal_close(w, w_o, 1e-14, "weights")


@pytest.mark.parametrize("x", [-800.0, -740.0, -700.0, 690.0, 720.0, float("nan")])
def test_weights_normaliser_special_cases(backend, x):
    """weights.py:122-123 divides by the sum over models: 0/0 (every member underflows), a subnormal sum, a
    sum near the overflow threshold, inf/inf and NaN must come out as the division gives them."""
    import torch

    M, T = 10, 40
    xs = np.full(M, x)
    xs[0] = x - 3.0
    obs = torch.zeros(1, 3, T, dtype=torch.float64, device=backend.device)
    st = _stats_for_exponent(backend, xs, T)
    w, le, _ = backend.loglik_weights_mvn(st, obs, M, want_lls=True)
    assert torch.equal(w.nan_to_num(nan=-1.0), backend.loglik_weights_mvn(st, obs, M).nan_to_num(nan=-1.0))
    w, le = w.cpu().numpy(), le.cpu().numpy()
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        w_o = le / le.sum(axis=1, keepdims=True)
    _nan_equal_close(w, w_o, 1e-13, "weights")


def test_normal_branch(backend):
    rng = np.random.default_rng(5)
    C, M, Ro, N = 2, 4, 3, 77
    loc = rng.normal(size=(C, M, N))
    scale = rng.uniform(0.05, 0.4, size=(C, M, N))
    obs = rng.normal(size=(C, Ro, N))
    w, le, lm = backend.loglik_weights_normal(_t(backend, loc), _t(backend, scale), _t(backend, obs), want_lls=True)
    for c in range(C):
        wo, eo, lo = rp.loglik_weights_normal(loc[c], scale[c], obs[c])
        _nan_equal_close(lm[c].cpu().numpy(), lo, 1e-12, "lls_mean")
        _nan_equal_close(w[c].cpu().numpy(), wo, TOL_WEIGHTS, "weights")
    ll = backend.normal_logprob(_t(backend, loc), _t(backend, scale), _t(backend, loc + 0.1)).cpu().numpy()
    assert rel_err(ll, rp.normal_log_prob(loc, scale, loc + 0.1)) < 1e-13


def test_weights_time_mean_skips_nan(backend):
    rng = np.random.default_rng(6)
    w = rng.uniform(size=(2, 3, 50))
    w[0, :, 5:9] = np.nan
    w[1, 2, :] = np.nan
    out = backend.weights_time_mean(_t(backend, w)).cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.broadcast_to(np.nanmean(w, axis=2)[..., None], w.shape)
    _nan_equal_close(out, want, 1e-14, "time mean")


# ------------------------------------------------------------------------------------ a5/a6 barycentre
def test_barycentre_vs_oracle_all_regimes(backend):
    rng = np.random.default_rng(7)
    C, M, N = 2, 6, 40
    means = rng.normal(size=(C, M, N))
    variances = rng.uniform(0.001, 0.05, size=(C, M, N))
    variances[1] = rng.uniform(1.5, 30.0, size=(M, N))  # S > 1: the iteration really runs
    w = rng.uniform(size=(C, M, N))
    w /= w.sum(axis=1, keepdims=True)
    w[0, :, 3] = np.nan  # a 0/0 column from the weights stage
    mu, sd, it = backend.barycentre_1d(_t(backend, means), _t(backend, variances), _t(backend, w))
    for c in range(C):
        bmu, bsd, bit = rp.barycentre_points(means[c], variances[c], w[c])
        _nan_equal_close(mu[c].cpu().numpy(), bmu, 1e-13, "bary mu")
        _nan_equal_close(sd[c].cpu().numpy(), bsd, 1e-13, "bary sd")
        assert np.array_equal(it[c].cpu().numpy(), bit)
    assert int(it[0].max()) == 201 and int(it[1].min()) > 3


def test_single_point_gaussian_barycentre(backend):
    from bayesian_ensembling_b200 import gaussian_barycentre

    for stds in ([0.1, 0.3], [2.0, 4.0]):
        mu, sd = gaussian_barycentre([1.0, 3.0], stds, [0.25, 0.75])
        mo, so, _ = rp.gaussian_barycentre([1.0, 3.0], stds, [0.25, 0.75])
        assert abs(mu - mo) < 1e-14 and abs(sd - so) < 1e-14


def test_member_sharded_partials_equal_direct(backend):
    """Emulates 2 ranks on one GPU: partial sums over each half of the members, summed, give
    the single-GPU weights / barycentre (the NCCL all-reduce is a plain sum of these buffers)."""
    reals, obs = _cell(6, 3, 60, 4, seed=8)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    post = backend.gp_posterior(X, ym, yv, np.full(6, 0.5), np.full(6, 6.0), want_cov=False, want_scale_tri=False)
    ob = _t(backend, obs[None])
    w, le, _ = backend.loglik_weights_mvn(post.mvn_stats, ob, 6, want_lls=True)
    mu_d, sd_d, _ = backend.barycentre_1d(post.mu[None], post.var_diag[None], w)
    parts = []
    for lo, hi in ((0, 3), (3, 6)):
        _, le_l, _ = backend.loglik_weights_mvn(post.mvn_stats[lo:hi], ob, hi - lo, want_lls=True)
        parts.append(backend.barycentre_1d_partial(post.mu[None, lo:hi], post.var_diag[None, lo:hi], le_l))
    total = parts[0] + parts[1]
    mu_s, sd_s, _ = backend.barycentre_1d_finish(total)
    w_s = backend.weights_normalise(le[:, 0:3].contiguous(), total[0])
    _nan_equal_close(mu_s.cpu().numpy(), mu_d.cpu().numpy(), 1e-12, "mu")
    _nan_equal_close(sd_s.cpu().numpy(), sd_d.cpu().numpy(), 1e-12, "sd")
    _nan_equal_close(w_s.cpu().numpy(), w[:, 0:3].cpu().numpy(), 1e-12, "w")


# ------------------------------------------------------------------------------------ whole path
@pytest.mark.parametrize("time_mean", [False, True])
def test_cfg1_pipeline_vs_oracle(backend, time_mean):
    """BASELINE config 1: 10 models x 3 realisations x 251 annual steps."""
    from bayesian_ensembling_b200 import grid

    cfg = synthetic.CONFIGS["cfg1"]
    reals, obs = synthetic.make_cells(cfg)
    res = grid.fit_weight_barycentre(reals, obs, synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE,
                                     keep_posteriors=True, time_mean_weights=time_mean)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = rp.cell_pipeline_L1(reals[0], obs[0], synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE,
                                time_mean_weights=time_mean)
    assert rel_err(res.mu[0].cpu().numpy(), o["mu"]) <= TOL_POSTERIOR
    assert rel_err(res.cov[0].cpu().numpy(), o["cov"]) <= TOL_POSTERIOR
    _nan_equal_close(res.weights[0].cpu().numpy(), o["weights"], TOL_WEIGHTS, "weights")
    _nan_equal_close(res.bary_mu[0].cpu().numpy(), o["bary_mu"], TOL_WEIGHTS, "bary_mu")
    _nan_equal_close(res.bary_std[0].cpu().numpy(), o["bary_std"], TOL_WEIGHTS, "bary_std")


def test_multi_cell_waves_equal_single_wave(backend):
    from bayesian_ensembling_b200 import grid

    cfg = synthetic.Config("t", 9, 5, 3, 3, 90, 2, False, "")
    reals, obs = synthetic.make_cells(cfg, seed=77)
    a = grid.fit_weight_barycentre(reals, obs, 0.5, 6.0)
    b = grid.fit_weight_barycentre(reals, obs, 0.5, 6.0, cells_per_wave=2)
    for name in ("weights", "bary_mu", "bary_std", "mu", "var_diag"):
        x, y = getattr(a, name).cpu().numpy(), getattr(b, name).cpu().numpy()
        assert np.array_equal(x, y, equal_nan=True), name


def test_full_size_properties_T3012(backend):
    """BASELINE config 2 size (T=3012): the oracle takes seconds per member here, so one member
    is compared directly and the batch is checked through size-independent identities:
    (K+E) * Minv == I with Minv recovered from cov, scale_tri scale_tri^T == cov, symmetry."""
    import torch

    cfg = synthetic.CONFIGS["cfg2"]
    reals, obs = synthetic.make_cells(cfg)
    M = 3
    r = _t(backend, reals[0, :M])
    X, ym, yv = backend.gpdtw1d_inputs(r)
    var = torch.full((M,), synthetic.L1_VARIANCE, dtype=torch.float64, device=backend.device)
    ls = torch.full((M,), synthetic.L1_LENGTHSCALE, dtype=torch.float64, device=backend.device)
    post = backend.gp_posterior(X, ym, yv, var, ls)
    assert int(post.info_fit.abs().sum()) == 0 and int(post.info_dist.abs().sum()) == 0
    K = backend.matern32_gram(X, var, ls)
    E = yv + 1e-6
    P = post.cov - torch.diag_embed(yv)  # = E - E Minv E
    Minv = (torch.diag_embed(E) - P) / E[:, :, None] / E[:, None, :]
    I = torch.eye(cfg.steps, dtype=torch.float64, device=backend.device)
    resid = ((K + torch.diag_embed(E)) @ Minv - I).abs().max().item()
    assert resid < 1e-8, resid
    llt = post.scale_tri @ post.scale_tri.transpose(1, 2)
    assert ((llt - post.cov).abs().max() / post.cov.abs().max()).item() < 1e-13
    assert torch.equal(post.cov, post.cov.transpose(1, 2))
    # one member against the oracle at full size
    Xo, yo, so = rp.gpdtw1d_inputs(reals[0, 0])
    mean, cov = rp.gp_posterior_closed_form(Xo, yo, so, synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE)
    assert rel_err(post.mu[0].cpu().numpy(), mean) <= TOL_POSTERIOR
    assert rel_err(post.cov[0].cpu().numpy(), cov) <= TOL_POSTERIOR
    assert np.abs(post.var_diag[0].cpu().numpy() / np.diag(cov) - 1).max() <= TOL_POSTERIOR


# ------------------------------------------------------------------------------------ the reference's API
def test_reference_api_end_to_end(backend):
    """ModelCollection.fit(GPDTW1D) -> LogLikelihoodWeight -> Barycentre with the reference's
    call signatures (tests/test_weights.py:86-101, utils.py:102-135)."""
    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200.labelled import DataArray

    M, R, T, Ro = 5, 3, 24, 2
    reals, obs = _cell(M, R, T, Ro, seed=21, monthly=True)
    time = np.arange(T)
    pms = [es.ProcessModel(DataArray(reals[m], ("realisation", "time"),
                                     {"realisation": np.arange(R), "time": time}), f"model{m}") for m in range(M)]
    obs_pm = es.ProcessModel(DataArray(obs, ("realisation", "time"), {"realisation": np.arange(Ro), "time": time}),
                             "obs")
    mc = es.ModelCollection(pms)
    mc.fit(es.GPDTW1D(hyperparameters=(0.5, 6.0), y_mean="mean"), compile_objective=True, n_optim_nits=2,
           progress_bar=False)
    weights = es.LogLikelihoodWeight()(mc, obs_pm)
    assert weights.shape == (M, T) and weights.dims == ("model", "time")
    o = rp.cell_pipeline_L1(reals, obs, 0.5, 6.0)
    _nan_equal_close(weights.values, o["weights"], TOL_WEIGHTS, "weights")
    for m in range(M):
        d = mc[m].distribution
        assert rel_err(d.mean.values, o["mu"][m]) <= TOL_POSTERIOR
        assert rel_err(d.variance.values, np.diag(o["cov"][m])) <= TOL_POSTERIOR
        assert rel_err(d._dist.covariance(), o["cov"][m]) <= TOL_POSTERIOR
    bary = es.Barycentre()(mc, weights)
    _nan_equal_close(bary.mean.values, o["bary_mu"], TOL_WEIGHTS, "bary mean")
    # Q-SCALE: Distribution(mu, covariance=sd**2, MultivariateNormalDiag) => variance == sd**4
    _nan_equal_close(bary.variance.values, o["bary_std"] ** 4, TOL_WEIGHTS, "bary variance")
    # distribution.log_prob of a genuine [T] vector and of the [T,1] quirk input
    ll = mc[0].distribution._dist.log_prob(obs[0][:, None])
    assert rel_err(ll, rp.mvn_log_prob(o["mu"][0], o["scale_tri"][0], obs[0][:, None])) < 1e-9
    ll1 = mc[0].distribution._dist.log_prob(obs[0])
    assert abs(ll1 - rp.mvn_log_prob(o["mu"][0], o["scale_tri"][0], obs[0])) <= 1e-9 * abs(ll1)


# ------------------------------------------------------------------------------------ a1 L2: the training loop
L2_TOL = 1e-7  # measured: see DESIGN.md "L2 parity"; the loop amplifies rounding through Adam


@pytest.mark.parametrize("T,R,M,n_it,train", [(24, 3, 3, 2, True), (24, 3, 2, 25, True), (86, 5, 2, 40, True),
                                              (130, 4, 2, 10, True), (165, 10, 2, 12, True), (40, 3, 2, 30, False),
                                              (24, 3, 2, 0, True), (86, 5, 2, 400, True), (251, 3, 1, 60, True)])
def test_vgp_fit_vs_oracle(backend, T, R, M, n_it, train):
    """GPDTW1D.fit below the DBA step (models.py:179-220): natgrad(0.5) + Adam(0.01) from GPflow's
    initial state, then predict_f(full_cov=True) + diag(y_var)."""
    reals, _ = _cell(M, R, T, 2, seed=1000 + T + n_it)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    post, var, ls = backend.vgp_fit(X, ym, yv, n_it, train_hypers=train)
    assert int(post.info_fit.abs().sum()) == 0 and int(post.info_dist.abs().sum()) == 0
    worst = 0.0
    for m in range(M):
        mu_o, cov_o, st = rp.gpdtw1d_fit(reals[m], n_optim_nits=n_it, train_hypers=train, return_state=True)
        assert abs(float(var[m]) / st["variance"] - 1) <= L2_TOL and abs(float(ls[m]) / st["lengthscale"] - 1) <= L2_TOL
        e_mu = rel_err(post.mu[m].cpu().numpy(), mu_o)
        e_cov = rel_err(post.cov[m].cpu().numpy(), cov_o)
        e_var = float(np.abs(post.var_diag[m].cpu().numpy() / np.diag(cov_o) - 1).max())
        e_tri = rel_err(post.scale_tri[m].cpu().numpy(), np.linalg.cholesky(cov_o))
        worst = max(worst, e_mu, e_cov, e_var, e_tri)
        assert max(e_mu, e_cov, e_var, e_tri) <= L2_TOL, (e_mu, e_cov, e_var, e_tri)
    print(f"vgp_fit T={T} n_it={n_it}: worst rel err {worst:.2e}")


def test_vgp_fit_converges_to_closed_form(backend):
    """With frozen hyper-parameters the loop converges geometrically to the closed-form
    posterior be_gp_posterior computes."""
    reals, _ = _cell(2, 4, 60, 2, seed=5)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    post, var, ls = backend.vgp_fit(X, ym, yv, 60, train_hypers=False, init_variance=0.5, init_lengthscale=6.0)
    ref = backend.gp_posterior(X, ym, yv, [0.5, 0.5], [6.0, 6.0])
    assert rel_err(post.mu.cpu().numpy(), ref.mu.cpu().numpy()) < 1e-6
    assert rel_err(post.cov.cpu().numpy(), ref.cov.cpu().numpy()) < 1e-6


def test_reference_api_default_fit_runs_training_loop(backend):
    """tests/test_weights.py:90 verbatim: fit(GPDTW1D(), compile_objective=True, n_optim_nits=2)."""
    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200.labelled import DataArray

    M, R, T, Ro = 2, 3, 24, 2
    reals, obs = _cell(M, R, T, Ro, seed=31, monthly=True)
    pms = [es.ProcessModel(DataArray(reals[m], ("realisation", "time")), f"model{m}") for m in range(M)]
    mc = es.ModelCollection(pms)
    mc.fit(model=es.GPDTW1D(), compile_objective=True, n_optim_nits=2, progress_bar=False)
    from oracle import dba

    for m in range(M):
        y_dba = dba.dba_subgradient(reals[m], max_iter=50, tol=1e-3)[0]  # models.py:176-178
        mu_o, cov_o = rp.gpdtw1d_fit(reals[m], n_optim_nits=2, y_mean=y_dba)
        assert rel_err(mc[m].distribution.mean.values, mu_o) <= L2_TOL
        assert rel_err(mc[m].distribution._dist.covariance(), cov_o) <= L2_TOL
    obs_pm = es.ProcessModel(DataArray(obs, ("realisation", "time")), "obs")
    w = es.LogLikelihoodWeight()(mc, obs_pm)
    assert w.shape == (M, T)
    ok = ~np.isnan(w.values).any(axis=0)
    assert np.allclose(w.values[:, ok].sum(axis=0), 1.0, atol=1e-6)


# ------------------------------------------------------------------------------------ a7 / a8 / a9
def _posterior_covs(M, R, T, seed):
    reals, _ = _cell(M, R, T, 2, seed=seed)
    mus, covs = [], []
    for m in range(M):
        X, y, s = rp.gpdtw1d_inputs(reals[m])
        mu, cov = rp.gp_posterior_closed_form(X, y, s, 0.5, 6.0)
        mus.append(mu)
        covs.append(cov)
    return np.asarray(mus), np.asarray(covs)


@pytest.mark.parametrize("T", [1, 2, 17, 86, 128, 165, 251, 300])
def test_sqrtm_vs_svd_oracle(backend, T):
    """wasserstein.py:10-13 on posterior covariances (what the path feeds it) and on a random SPD matrix."""
    _, covs = _posterior_covs(2, 4, T, seed=T)
    rng = np.random.default_rng(T)
    G = rng.standard_normal((T, T + 3))
    A = np.concatenate([covs, (G @ G.T / T + 1e-3 * np.eye(T))[None]])
    out, inv, iters, info = backend.sqrtm_psd(_t(backend, A), want_inverse=True)
    assert int(info.abs().sum()) == 0 and 1 <= iters <= 25
    worst = 0.0
    for b in range(A.shape[0]):
        want = rp.sqrtm_svd(A[b])
        got = out[b].cpu().numpy()
        worst = max(worst, rel_err(got, want))
        assert rel_err(got @ got, A[b]) < 1e-11
        assert np.abs(got @ inv[b].cpu().numpy() - np.eye(T)).max() < 1e-9
        assert np.array_equal(got, got.T)
    print(f"sqrtm T={T}: {iters} iterations, worst rel err {worst:.2e}")
    assert worst < 1e-9  # tolerance of the north star for barycentre moments is 1e-6


def test_sqrtm_reports_non_spd(backend):
    A = np.eye(5)
    A[3, 3] = -1.0
    _, _, _, info = backend.sqrtm_psd(_t(backend, A[None]))
    assert int(info[0]) == 4


@pytest.mark.parametrize("T", [3, 86, 165])
def test_w2_distance_vs_oracle(backend, T):
    mus, covs = _posterior_covs(4, 5, T, seed=100 + T)
    pairs = [(0, 1), (1, 2), (2, 3), (3, 0), (1, 1)]
    i, j = [p[0] for p in pairs], [p[1] for p in pairs]
    w2, info = backend.w2_distance(_t(backend, mus[i]), _t(backend, covs[i]), _t(backend, mus[j]), _t(backend, covs[j]))
    assert int(info.abs().sum()) == 0
    want = np.array([rp.gaussian_w2_distance(mus[a], covs[a], mus[b], covs[b]) for a, b in pairs])
    assert np.abs(w2.cpu().numpy() - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    # full_cov=False branch: variances on a diagonal
    var = np.asarray([np.diag(c) for c in covs])
    w2d = backend.w2_distance_diag(_t(backend, mus[i]), _t(backend, var[i]), _t(backend, mus[j]), _t(backend, var[j]))
    wantd = np.array([rp.gaussian_w2_distance(mus[a], np.diag(var[a]), mus[b], np.diag(var[b])) for a, b in pairs])
    assert np.abs(w2d.cpu().numpy() - wantd).max() <= 1e-10 * max(1.0, np.abs(wantd).max())


def test_w2_python_api(backend):
    from bayesian_ensembling_b200 import dists, gaussian_w2_distance_distrax, sqrtm, wasserstien_distance

    mus, covs = _posterior_covs(2, 5, 40, seed=7)
    a = dists.MultivariateNormalFullCovariance(mus[0], covs[0])
    b = dists.MultivariateNormalFullCovariance(mus[1], covs[1])
    assert abs(gaussian_w2_distance_distrax(a, b) - rp.gaussian_w2_distance(mus[0], covs[0], mus[1], covs[1])) < 1e-9
    va, vb = np.diag(covs[0]), np.diag(covs[1])
    assert abs(gaussian_w2_distance_distrax(a, b, full_cov=False)
               - rp.gaussian_w2_distance(mus[0], np.diag(va), mus[1], np.diag(vb))) < 1e-10
    assert rel_err(sqrtm(covs[0]), rp.sqrtm_svd(covs[0])) < 1e-9
    z = np.zeros(40)
    assert abs(wasserstien_distance(covs[0], covs[1]) - rp.gaussian_w2_distance(z, covs[0], z, covs[1])) < 1e-9
    with pytest.raises(ValueError):
        sqrtm(-np.eye(3))


@pytest.mark.parametrize("T,M,scale", [(1, 3, 1.0), (1, 3, 2000.0), (24, 3, 1.0), (86, 4, 1.0), (40, 3, 2000.0), (130, 2, 500.0)])
def test_fullcov_barycentre_vs_oracle(backend, T, M, scale):
    """a9 against the oracle's definition; scale > 1 pushes tr(S) above init_var so that the fixed
    point iterates (with degC-anomaly covariances it exits at iteration 0, like the 1-D rule)."""
    mus, covs = _posterior_covs(M, 4, T, seed=300 + T)
    covs = covs * scale
    rng = np.random.default_rng(T)
    w = rng.uniform(0.2, 1.0, M)
    w /= w.sum()
    mu, S, iters, info = backend.barycentre_fullcov(_t(backend, mus[None]), _t(backend, covs[None]), _t(backend, w[None]))
    mo, So, ito = rp.fullcov_barycentre(mus, covs, w)
    assert int(info.abs().sum()) == 0
    assert iters[0] == ito, (iters, ito)
    if scale > 1.0:
        assert ito > 0
    assert rel_err(mu[0].cpu().numpy(), mo) < 1e-12
    assert rel_err(S[0].cpu().numpy(), So) < TOL_WEIGHTS
    print(f"fullcov barycentre T={T} M={M}: {ito} iterations, rel err {rel_err(S[0].cpu().numpy(), So):.2e}")


def test_fullcov_barycentre_reduces_to_1d_kernel(backend):
    """T = 1: the matrix fixed point and its stop rule ARE gaussian_barycentre (wasserstein.py:61-100)."""
    rng = np.random.default_rng(5)
    for scale in (0.05, 30.0):
        M = 5
        mus = rng.normal(size=(M, 1))
        var = rng.uniform(0.5, 2.0, size=(M, 1, 1)) * scale
        w = rng.uniform(0.1, 1.0, M)
        w /= w.sum()
        mu, S, iters, _ = backend.barycentre_fullcov(_t(backend, mus[None]), _t(backend, var[None]), _t(backend, w[None]))
        bm, bs, bi = backend.barycentre_1d(_t(backend, mus.reshape(1, M, 1)), _t(backend, var.reshape(1, M, 1)),
                                           _t(backend, w.reshape(1, M, 1)))
        assert iters[0] == int(bi.item())
        assert abs(float(S.item()) - float(bs.item()) ** 2) < 1e-13 * max(1.0, float(S.item()))
        assert abs(float(mu.item()) - float(bm.item())) < 1e-14


def test_fullcov_barycentre_cells_batched(backend):
    """Several cells in one call == one call per cell (cells converge at different iterations)."""
    C, M, T = 3, 3, 30
    mus, covs = _posterior_covs(C * M, 4, T, seed=11)
    mus, covs = mus.reshape(C, M, T), covs.reshape(C, M, T, T).copy()
    covs[1] *= 50.0
    covs[2] *= 400.0
    w = np.full((C, M), 1.0 / M)
    mu, S, iters, _ = backend.barycentre_fullcov(_t(backend, mus), _t(backend, covs), _t(backend, w))
    assert len(set(iters)) > 1
    for c in range(C):
        mu1, S1, it1, _ = backend.barycentre_fullcov(_t(backend, mus[c:c + 1]), _t(backend, covs[c:c + 1]),
                                                     _t(backend, w[c:c + 1]))
        assert it1[0] == iters[c]
        assert rel_err(S[c].cpu().numpy(), S1[0].cpu().numpy()) < 1e-12
        assert rel_err(mu[c].cpu().numpy(), mu1[0].cpu().numpy()) < 1e-14


# ------------------------------------------------------------------------------------ SURVEY 8f "next": CRPS / similarity weights
@pytest.mark.parametrize("M", [24, 40, 200])
def test_weight_kernels_staging_regimes(backend, M):
    """The normalising kernels stage un-normalised weights in shared memory ([M][128] up to M = 32, [M][64]
    up to 96 KB, the output array itself beyond): all three regimes against the oracle."""
    rng = np.random.default_rng(M)
    C, Ro, N = 2, 3, 151
    loc = rng.normal(size=(C, M, N))
    scale = rng.uniform(0.2, 0.6, size=(C, M, N))
    obs = rng.normal(size=(C, Ro, N))
    # points 0..9: every member's log-likelihood is hugely negative but representable (total below 2^-500:
    # the plain-division path); points 10..19: exp underflows to 0 for every member -> 0/0 = NaN (quirk
    # Q-EXP); points 20..29: member 0 alone is far off (its weight is tiny or denormal next to a normal total)
    obs[:, :, :10] += 17.0
    obs[:, :, 10:20] += 300.0
    loc[:, 0, 20:30] += 12.0
    w = backend.loglik_weights_normal(_t(backend, loc), _t(backend, scale), _t(backend, obs))
    assert bool(np.isnan(w.cpu().numpy()[:, :, 10:20]).all()) and not bool(np.isnan(w.cpu().numpy()[:, :, 20:]).any())
    wc = backend.crps_weights(_t(backend, loc), _t(backend, scale), _t(backend, obs))
    ws = backend.similarity_weights_pointwise(_t(backend, loc), _t(backend, scale * scale))  # variance() = scale**2
    a2 = rng.uniform(0.5, 1.5, size=C * M)
    stats = np.stack([a2, a2 * rng.uniform(0.9, 1.1, C * M), a2 * rng.uniform(1.0, 1.2, C * M),
                      -0.5 * N * 1.8378770664093453 + rng.random(C * M)], axis=1)
    wm = backend.loglik_weights_mvn(_t(backend, stats), _t(backend, obs), M).cpu().numpy()
    for c in range(C):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            w_o = rp.loglik_weights_normal(loc[c], scale[c], obs[c])[0]
            e_o = rp.loglik_weights_normal(loc[c], scale[c], obs[c])[1]
        # columns whose largest un-normalised weight is itself denormal carry only a few bits on either side
        ok = (e_o.max(axis=0) > 1e-290) | (e_o.max(axis=0) == 0.0)
        assert ok[20:].all() and ok[:10].sum() >= 3
        _nan_equal_close(w[c].cpu().numpy()[:, ok], w_o[:, ok], 1e-12, "normal")
        assert rel_err(wc[c].cpu().numpy(), rp.crps_weights(loc[c], scale[c], obs[c])[0]) < 1e-12
        if M <= 24:  # the oracle's triple Python loop is O(M^2 N); beyond that its vectorised statement
            ws_o = rp.model_similarity_weights_temporal(loc[c], scale[c])[0]
        else:
            v = scale[c] ** 2
            ri = np.sqrt(v)
            d = np.abs(loc[c][:, None] - loc[c][None]) + ((v[:, None] + v[None]) - 2.0 * np.sqrt(ri[:, None] * v[None] * ri[:, None]))
            ws_o = d.mean(axis=1) / d.mean(axis=1).sum(axis=0)
        assert rel_err(ws[c].cpu().numpy(), ws_o) < 1e-10
        st = stats[c * M:(c + 1) * M]
        m1, m2 = obs[c].mean(axis=0), (obs[c] ** 2).mean(axis=0)
        ll = -0.5 * (m2[None] * st[:, :1] - 2.0 * m1[None] * st[:, 1:2] + st[:, 2:3]) - 0.5 * N * np.log(2 * np.pi) - st[:, 3:4]
        e = np.exp(ll)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            wm_o = e / e.sum(axis=0)
        okm = (e.max(axis=0) > 1e-290) | (e.max(axis=0) == 0.0)
        _nan_equal_close(wm[c][:, okm], wm_o[:, okm], 1e-10, "mvn")
        fin = ~np.isnan(wm_o).any(axis=0) & okm
        assert fin.sum() > 100 and np.abs(wm[c][:, fin].sum(axis=0) - 1.0).max() < 1e-12


def test_crps_weights_vs_oracle(backend):
    rng = np.random.default_rng(3)
    C, M, Ro, N = 2, 5, 3, 57
    loc = rng.normal(size=(C, M, N))
    var = rng.uniform(0.01, 0.5, size=(C, M, N))  # passed as the SCALE (quirk Q-SCALE)
    obs = rng.normal(size=(C, Ro, N))
    w, cm = backend.crps_weights(_t(backend, loc), _t(backend, var), _t(backend, obs), want_crps=True)
    for c in range(C):
        wo, co = rp.crps_weights(loc[c], var[c], obs[c])
        assert rel_err(cm[c].cpu().numpy(), co) < 1e-13
        assert rel_err(w[c].cpu().numpy(), wo) < 1e-12
        assert np.abs(w[c].cpu().numpy().sum(axis=0) - 1.0).max() < 1e-12


def test_crps_special_scales_and_far_observations(backend):
    """Scales that are zero, subnormal, infinite or NaN keep the reference's division (inf / NaN results equal
    to the oracle's); observations hundreds of scales away take the library exp; many realisations."""
    rng = np.random.default_rng(8)
    C, M, Ro, N = 1, 8, 10, 64
    loc = rng.normal(size=(C, M, N))
    scale = rng.uniform(0.002, 2.0, size=(C, M, N))
    scale[0, 0, :4] = [0.0, 1e-320, np.inf, np.nan]
    scale[0, 1, 4:8] = [1e-200, 1e200, 1e-3, 50.0]
    obs = rng.normal(size=(C, Ro, N))
    _, cm = backend.crps_weights(_t(backend, loc), _t(backend, scale), _t(backend, obs), want_crps=True)
    with np.errstate(all="ignore"):
        _, co = rp.crps_weights(loc[0], scale[0], obs[0])
    got = cm[0].cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(co))
    assert np.array_equal(np.isinf(got), np.isinf(co)) and np.array_equal(got[np.isinf(co)], co[np.isinf(co)])
    fin = np.isfinite(co)
    assert np.abs(got[fin] / co[fin] - 1.0).max() < 1e-13


def test_ksd_weights_vs_oracle(backend):
    """KSDWeight (weights.py:336-441): IMQ kernel Stein discrepancy per model and point."""
    rng = np.random.default_rng(4)
    C, M, Ro, N = 2, 5, 6, 43
    loc = rng.normal(size=(C, M, N))
    var = rng.uniform(0.05, 0.6, size=(C, M, N))  # passed as the SCALE (quirk Q-SCALE)
    obs = rng.normal(size=(C, Ro, N))
    w, k = backend.ksd_weights(_t(backend, loc), _t(backend, var), _t(backend, obs), want_ksd=True)
    for c in range(C):
        wo, ko = rp.ksd_weights(loc[c], var[c], obs[c])
        assert rel_err(k[c].cpu().numpy(), ko) < 1e-12
        assert rel_err(w[c].cpu().numpy(), wo) < 1e-12
        assert np.abs(w[c].cpu().numpy().sum(axis=0) - 1.0).max() < 1e-12
    # tightly clustered samples with the model mean inside the cluster: the regime where the kernel's factored
    # form (moments of the pair kernel about the sample mean, weights_next_kernels.cuh) cancels the most
    obs_c = 0.8 + 1e-3 * rng.normal(size=(C, 10, N))
    loc_c = 0.8 + 1e-3 * rng.normal(size=(C, M, N))
    w, k = backend.ksd_weights(_t(backend, loc_c), _t(backend, var), _t(backend, obs_c), want_ksd=True)
    for c in range(C):
        wo, ko = rp.ksd_weights(loc_c[c], var[c], obs_c[c])
        assert np.abs(k[c].cpu().numpy() / ko - 1.0).max() < 1e-11
        assert rel_err(w[c].cpu().numpy(), wo) < 1e-11
    # a single observation realisation: k0(a, a) = g^2 + 1
    w1, k1 = backend.ksd_weights(_t(backend, loc), _t(backend, var), _t(backend, obs[:, :1]), want_ksd=True)
    g = -(obs[:, :1] - loc) / var ** 2
    assert rel_err(k1.cpu().numpy(), np.sqrt(g * g + 1.0)) < 1e-14


def test_similarity_weights_vs_oracle(backend):
    mus, covs = _posterior_covs(4, 5, 30, seed=17)
    M = 4
    # mode "single": M*M full-covariance W2 distances, nanmean, normalise
    ii, jj = np.divmod(np.arange(M * M), M)
    w2, info = backend.w2_distance(_t(backend, mus[ii]), _t(backend, covs[ii]), _t(backend, mus[jj]), _t(backend, covs[jj]))
    w = backend.w2_collapse(w2.reshape(1, M, M, 1))[0, :, 0].cpu().numpy()
    wo, w2o = rp.model_similarity_weights_single(mus, covs)
    assert np.abs(w2.cpu().numpy().reshape(M, M) - w2o).max() < 1e-9
    assert rel_err(w, wo) < 1e-9
    # mode "temporal": 1-D W2 per time step with variance**2 (dx.Normal(mean, variance))
    var = np.asarray([np.diag(c) for c in covs])
    wt, w2t = backend.similarity_weights_pointwise(_t(backend, mus[None]), _t(backend, (var * var)[None]), want_w2=True)
    wto, w2to = rp.model_similarity_weights_temporal(mus, var)
    assert np.abs(w2t[0].cpu().numpy() - w2to).max() < 1e-13
    assert rel_err(wt[0].cpu().numpy(), wto) < 1e-12
    # without the pair distances the kernel takes its M-square-roots form wherever the point is clean; a NaN mean,
    # a negative variance or an infinite one at a point sends that point through the NaN-skipping pair loop
    wt_fast = backend.similarity_weights_pointwise(_t(backend, mus[None]), _t(backend, (var * var)[None]))
    assert rel_err(wt_fast[0].cpu().numpy(), wto) < 1e-12
    rng = np.random.default_rng(23)
    Mb, Nb = 24, 200
    mu_b, v_b = rng.normal(size=(Mb, Nb)), rng.uniform(0.05, 0.5, size=(Mb, Nb))
    mu_b[3, 5] = np.nan
    v_b[7, 9] = -0.1
    v_b[2, 11] = np.inf
    got = backend.similarity_weights_pointwise(_t(backend, mu_b[None]), _t(backend, (v_b * np.abs(v_b))[None]))[0].cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            r = np.sqrt(v_b * np.abs(v_b))
            dist = np.abs(mu_b[:, None] - mu_b[None]) + ((v_b * np.abs(v_b))[:, None] + (v_b * np.abs(v_b))[None]
                                                         - 2.0 * np.sqrt(r[:, None] * (v_b * np.abs(v_b))[None] * r[:, None]))
            mean_d = np.nanmean(dist, axis=1)
            want = mean_d / np.nansum(mean_d, axis=0)  # weights.py:331: xarray .sum('model') skips NaN
    _nan_equal_close(got, want, 1e-12, "temporal similarity weights")
    # NaN distances are skipped by the nanmean
    d = np.arange(1.0, 1.0 + M * M * 3).reshape(1, M, M, 3)
    d[0, 1, 2, 0] = np.nan
    got = backend.w2_collapse(_t(backend, d))[0].cpu().numpy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = np.nanmean(d[0], axis=1)
    assert rel_err(got, m / m.sum(axis=0)) < 1e-14


def test_reference_weight_classes_shapes_and_sums(backend):
    """The reference's own weight test (tests/test_weights.py:71-101): every weight class returns a
    labelled array of shape (M,) + obs.mean('realisation').shape that sums to 1 over models."""
    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200.labelled import DataArray

    M, R, T, Ro = 5, 3, 24, 2
    reals, obs = _cell(M, R, T, Ro, seed=31, monthly=True)
    time = np.arange(T)
    pms = [es.ProcessModel(DataArray(reals[m], ("realisation", "time"),
                                     {"realisation": np.arange(R), "time": time}), f"model{m}") for m in range(M)]
    obs_pm = es.ProcessModel(DataArray(obs, ("realisation", "time"), {"realisation": np.arange(Ro), "time": time}), "obs")
    mc = es.ModelCollection(pms)
    mc.fit(es.GPDTW1D(), compile_objective=True, n_optim_nits=2, progress_bar=False)
    mus = np.stack([m.distribution._dist.mean() for m in mc])
    covs = np.stack([m.distribution._dist.covariance() for m in mc])
    var = np.stack([m.distribution._dist.variance() for m in mc])
    for cls, kwargs in ((es.InverseSquareWeight, {}), (es.UniformWeight, {}), (es.CRPSWeight, {}), (es.KSDWeight, {}),
                        (es.ModelSimilarityWeight, {"mode": "temporal"})):
        w = cls()(mc, obs_pm, **kwargs)
        assert w.shape == (M, T), cls
        assert np.abs(np.nansum(w.values, axis=0) - 1.0).max() < 1e-6, cls
    wc = es.CRPSWeight()(mc, obs_pm).values
    assert rel_err(wc, rp.crps_weights(mus, var, obs)[0]) < 1e-10
    ws = es.ModelSimilarityWeight()(mc, mode="temporal").values
    assert rel_err(ws, rp.model_similarity_weights_temporal(mus, var)[0]) < 1e-10
    wk = es.KSDWeight()(mc, obs_pm).values
    assert rel_err(wk, rp.ksd_weights(mus, var, obs)[0]) < 1e-10
    w1 = es.ModelSimilarityWeight()(mc, mode="single")
    assert w1.shape == (M, 1) and w1.dims == ("model", "time")
    assert rel_err(w1.values[:, 0], rp.model_similarity_weights_single(mus, covs)[0]) < 1e-8
    wi = es.InverseSquareWeight()(mc, obs_pm).values
    assert rel_err(wi, rp.inverse_square_weights(reals.mean(axis=1), obs.mean(axis=0))) < 1e-12
    with pytest.raises(ValueError):
        es.ModelSimilarityWeight()(mc, mode="nope")


# ------------------------------------------------------------------------------------ checkpoints
def test_model_collection_npz_round_trip(backend, tmp_path):
    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200 import utils
    from bayesian_ensembling_b200.labelled import DataArray

    M, R, T = 3, 3, 30
    reals, obs = _cell(M, R, T, 2, seed=41)
    time = 1850 + np.arange(T)
    pms = [es.ProcessModel(DataArray(reals[m], ("realisation", "time"), {"realisation": np.arange(R), "time": time}),
                           f"model{m}") for m in range(M)]
    mc = es.ModelCollection(pms)
    mc.fit(es.GPDTW1D(hyperparameters=(0.5, 6.0)), progress_bar=False)
    path = str(tmp_path / "mc.npz")
    mc.save(path)
    mc2 = utils.load_model_collection(path)
    assert mc2.model_names == mc.model_names
    for a, b in zip(mc, mc2):
        assert np.array_equal(a.model_data.values, b.model_data.values)
        assert np.array_equal(a.time.values, b.time.values)
        assert np.array_equal(a.distribution._dist.mean(), b.distribution._dist.mean())
        assert np.array_equal(a.distribution._dist.covariance(), b.distribution._dist.covariance())
        assert rel_err(b.distribution._dist.scale_tri, a.distribution._dist.scale_tri) < 1e-13
    obs_pm = es.ProcessModel(DataArray(obs, ("realisation", "time"), {"realisation": np.arange(2), "time": time}), "obs")
    w1, w2 = es.LogLikelihoodWeight()(mc, obs_pm), es.LogLikelihoodWeight()(mc2, obs_pm)
    _nan_equal_close(w2.values, w1.values, 1e-10, "weights after reload")


def test_golden_members_through_checkpoint_loader(backend, golden_members, tmp_path):
    """The reference's fitted members (tests/golden) written in the reference's pickle LAYOUT and read
    back by load_reference_pickle: posteriors are rebuilt on the device (data.py:38-39 -> be_mvn_from_cov)."""
    import pickle
    import sys
    import types

    from bayesian_ensembling_b200 import utils

    # a throw-away module tree that pickles with the same shape as the reference's objects
    mod = types.ModuleType("fake_ensembles_data")
    for name in ("ModelCollection", "ProcessModel", "Distribution", "DataArray", "Variable", "MVN"):
        setattr(mod, name, type(name, (), {"__module__": "fake_ensembles_data"}))
    sys.modules["fake_ensembles_data"] = mod
    try:
        members = golden_members[:2]
        pms = []
        for g in members:
            var = mod.Variable()
            var.__dict__.update(_dims=("realisation", "time"), _data=g.realisations)
            da = mod.DataArray()
            da.__dict__.update(_name="tas", _variable=var)
            inner = mod.MVN()
            inner.__dict__.update(_loc=g.mu, _covariance_matrix=g.cov, _scale_tri=g.scale_tri)
            dist = mod.Distribution()
            dist.__dict__.update(mu=g.mu, covariance=g.cov, _dist=inner)
            pm = mod.ProcessModel()
            pm.__dict__.update(model_data=da, model_name=g.name, idx=0, _distribution=dist)
            pms.append(pm)
        mc = mod.ModelCollection()
        mc.__dict__.update(models=pms, idx=0)
        path = str(tmp_path / "ref_layout.pkl")
        with open(path, "wb") as f:
            pickle.dump(mc, f, protocol=4)
    finally:
        del sys.modules["fake_ensembles_data"]
    loaded = utils.load_reference_pickle(path)
    for pm, g in zip(loaded, members):
        assert pm.model_name == g.name
        assert np.array_equal(pm.model_data.values, g.realisations)
        assert rel_err(pm.distribution._dist.scale_tri, g.scale_tri) < 1e-12  # pinned by the reference's own factor
        assert np.array_equal(pm.distribution._dist.mean(), g.mu)


def test_cfg5_properties_at_size(backend):
    """BASELINE config 5 shape at T=515 (the oracle's SVD-based fixed point takes minutes at 3012; the
    full size is run by tools/run_cfg5.py -> profiles/): with degC-anomaly covariances the signed stop
    rule exits at iteration 0, where S = sum_m w_m sqrtm(Sigma_m) exactly; S is symmetric positive
    definite; one member is checked against the SVD oracle."""
    import torch

    M, T = 4, 515
    mus, covs = _posterior_covs(M, 5, T, seed=55)
    w = np.array([0.1, 0.2, 0.3, 0.4])
    mu, S, iters, info = backend.barycentre_fullcov(_t(backend, mus[None]), _t(backend, covs[None]), _t(backend, w[None]))
    assert iters[0] == 0 and int(info.abs().sum()) == 0
    roots, _, _, _ = backend.sqrtm_psd(_t(backend, covs))
    want = (roots * _t(backend, w)[:, None, None]).sum(0)
    assert ((S[0] - want).abs().max() / want.abs().max()).item() < 1e-14
    assert torch.equal(S[0], S[0].T)
    _, pd = backend.potrf(S)
    assert int(pd.item()) == 0
    assert rel_err(roots[0].cpu().numpy(), rp.sqrtm_svd(covs[0])) < 1e-9
    assert rel_err(mu[0].cpu().numpy(), (w[:, None] * mus).sum(0)) < 1e-13


# ------------------------------------------------------------------------------------ the integration-level caller
def test_perfect_model_test_end_to_end(backend, tmp_path):
    """PerfectModelTest (ensembles/utils.py:32-228) through the mirrored API: every held-out model's six metrics
    against the oracle's composition of the same steps (DBA mean, fixed-theta posterior, LogLikelihoodWeight,
    time-mean, Barycentre, NLL / RMSE / W2 of barycentre and multi-model mean)."""
    import functools

    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200 import utils
    from bayesian_ensembling_b200.labelled import DataArray
    from oracle import dba

    M, R, Th, Tf = 4, 3, 30, 20
    # members that differ by less than their internal variability, so that the un-shifted exp of Q-EXP stays finite
    rng = np.random.default_rng(77)
    tn = np.linspace(0.0, 1.0, Th + Tf)
    reals = (1.5 * tn + tn ** 2)[None, None, :] + 0.03 * np.arange(M)[:, None, None] + synthetic._ar1(rng, (M, R, Th + Tf))
    hind, fore = reals[:, :, :Th], reals[:, :, Th:]

    def coll(block, t0):
        time = 1900 + t0 + np.arange(block.shape[2])
        return es.ModelCollection([es.ProcessModel(DataArray(block[m], ("realisation", "time"),
                                                             {"realisation": np.arange(R), "time": time}), f"model{m}")
                                   for m in range(M)])

    pmt = utils.PerfectModelTest(coll(hind, 0), coll(fore, Th), functools.partial(es.GPDTW1D, hyperparameters=(0.5, 6.0)),
                                 es.LogLikelihoodWeight, es.Barycentre, ssp="ssp-test", save_dir=str(tmp_path / "pmt"))
    res = pmt.run(n_optim_nits=2)
    assert res["columns"][0] == "model as psuedo obs" and len(res["rows"]) == M
    assert (tmp_path / "pmt" / "csvs" / "prefect_model_test_results_LogLikelihoodWeight_ssp-test.csv").exists()

    def dba_means(block):
        return np.stack([dba.dba_subgradient(b, max_iter=50, tol=1e-3)[0] for b in block])

    for i in range(M):
        others = [m for m in range(M) if m != i]
        oh = rp.cell_pipeline_L1(hind[others], hind[i], 0.5, 6.0, y_means=dba_means(hind[others]))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            w_mean = np.nanmean(oh["weights"], axis=1)                                   # utils.py:112
        of = rp.cell_pipeline_L1(fore[others], fore[i], 0.5, 6.0, y_means=dba_means(fore[others]))
        var_f = np.asarray([np.diag(c) for c in of["cov"]])
        bmu, bsd, _ = rp.barycentre_points(of["mu"], var_f, np.broadcast_to(w_mean[:, None], (M - 1, Tf)))  # :133-135
        Xt, yt, st = rp.gpdtw1d_inputs(fore[i], dba.dba_subgradient(fore[i], max_iter=50, tol=1e-3)[0])
        t_mu, t_cov = rp.gp_posterior_closed_form(Xt, yt, st, 0.5, 6.0)
        want = rp.perfect_model_metrics(bmu, bsd ** 2, t_mu, t_cov, fore[i], np.vstack(list(fore[others])))
        got = res["rows"][i]
        assert got[0] == f"model{i}" and np.isfinite(want).all()
        for name, g, w_ in zip(res["columns"][1:], got[1:], want):
            assert abs(g - w_) <= 1e-6 * max(abs(w_), 1e-12), (i, name, g, w_)


def test_mean_field_approximation(backend):
    """models.py:75-131: the returned Distribution holds the INITIAL moments (the Adam loop is dead code)."""
    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200 import dists
    from bayesian_ensembling_b200.labelled import DataArray

    rng = np.random.default_rng(6)
    data = rng.normal(size=(4, 12, 3, 5))  # realisation, time, lat, lon
    pm = es.ProcessModel(DataArray(data, ("realisation", "time", "latitude", "longitude"),
                                   {"realisation": np.arange(4), "time": np.arange(12), "latitude": np.arange(3.0),
                                    "longitude": np.arange(5.0)}), "m")
    with pytest.warns(UserWarning, match="No optimiser specified"):
        d = es.MeanFieldApproximation().fit(pm, n_optim_nits=3)
    assert isinstance(d._dist, dists.Normal)
    flat = data.reshape(4, -1)
    assert rel_err(d._dist.mean(), flat.mean(axis=0)) < 1e-15
    assert rel_err(d._dist.stddev(), flat.var(axis=0)) < 1e-13       # the variance sits in the SCALE slot (Q-SCALE)
    assert d.mean.shape == (12, 3, 5)
