"""Round-2 behaviour fixes, each against the oracle or a NumPy statement of the reference's semantics:
NaN-skipping normalisers (xarray ``.sum('model')``), per-problem convergence of the batched matrix square
root, the general-vector MVN log-density against the stored factor, and members without constant-vector
statistics in LogLikelihoodWeight.  Needs a B200: ``-m gpu``."""
import warnings

import numpy as np
import pytest
import scipy.linalg as sla

from bayesian_ensembling_b200 import synthetic
from oracle import reference_path as rp
from conftest import rel_err

pytestmark = pytest.mark.gpu


def _t(backend, a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=backend.device)


def _cell(M, R, T, Ro, seed):
    cfg = synthetic.Config("t", 9, 1, M, R, T, Ro, False, "")
    reals, obs = synthetic.make_cells(cfg, seed=seed)
    return reals[0], obs[0]


def _posteriors(M, R, T, seed):
    reals, obs = _cell(M, R, T, 4, seed)
    mus, covs = [], []
    for m in range(M):
        X, y, s = rp.gpdtw1d_inputs(reals[m])
        mu, cov = rp.gp_posterior_closed_form(X, y, s, 0.5, 6.0)
        mus.append(mu)
        covs.append(cov)
    return np.asarray(mus), np.asarray(covs), obs


@pytest.mark.parametrize("M", [5, 40])
def test_one_nan_member_leaves_the_others_normalised(backend, M):
    """weights.py:122: xarray's .sum('model') skips NaN, so a member whose statistics are NaN (a failed
    Cholesky) gets NaN weights while the others still sum to one -- in every normalising kernel."""
    T, Ro = 40, 3
    mus, covs, obs = _posteriors(M, 3, T, seed=12 + M)
    tris = np.stack([np.linalg.cholesky(c) for c in covs])
    stats = []
    for m in range(M):
        a = sla.solve_triangular(tris[m], np.ones(T), lower=True)
        b = sla.solve_triangular(tris[m], mus[m], lower=True)
        stats.append([a @ a, a @ b, b @ b, np.log(np.diag(tris[m])).sum()])
    stats = np.array(stats) * np.array([1e-3, 1e-3, 1e-3, 0.0]) + np.array([0, 0, 0, -0.5 * T * rp.LOG_2PI])  # finite weights
    bad = 2
    stats[bad] = np.nan
    w, le, lm = backend.loglik_weights_mvn(_t(backend, stats), _t(backend, obs[None, :Ro]), M, want_lls=True)
    w = w[0].cpu().numpy()
    assert np.isnan(w[bad]).all()
    ok = np.arange(M) != bad
    assert np.isfinite(w[ok]).all() and np.abs(w[ok].sum(axis=0) - 1.0).max() < 1e-12
    le = le[0].cpu().numpy()
    assert rel_err(w[ok], le[ok] / le[ok].sum(axis=0)) < 1e-12
    # Normal branch, CRPS, KSD: same semantics
    var = np.stack([np.diag(c) for c in covs])
    loc, scale = mus.copy(), np.sqrt(var) + 0.3
    loc[bad] = np.nan
    for fn in (backend.loglik_weights_normal, backend.crps_weights, backend.ksd_weights):
        wn = fn(_t(backend, loc[None]), _t(backend, scale[None]), _t(backend, obs[None, :Ro]))
        wn = (wn[0] if isinstance(wn, tuple) else wn)[0].cpu().numpy()
        assert np.isnan(wn[bad]).all(), fn.__name__
        assert np.abs(wn[ok].sum(axis=0) - 1.0).max() < 1e-12, fn.__name__
    # the oracle agrees (it normalises with nansum, as xarray does)
    wo, _ = rp.crps_weights(loc, scale, obs[:Ro])
    wc = backend.crps_weights(_t(backend, loc[None]), _t(backend, scale[None]), _t(backend, obs[None, :Ro]))[0].cpu().numpy()
    assert (np.isnan(wo) == np.isnan(wc)).all() and rel_err(wc[ok], wo[ok]) < 1e-10
    # member-sharded normaliser: the partial sum skips the NaN member too
    part = backend.barycentre_1d_partial(_t(backend, mus[None]), _t(backend, var[None]), _t(backend, le[None]))
    assert rel_err(part[0, 0].cpu().numpy(), le[ok].sum(axis=0)) < 1e-14


def test_sqrtm_batch_with_one_bad_problem_converges_the_rest(backend):
    """One non-SPD matrix in a batch must not stop the Denman-Beavers loop for the others."""
    from bayesian_ensembling_b200 import _lib  # noqa: F401

    T = 60
    _, covs, _ = _posteriors(3, 4, T, seed=3)
    rng = np.random.default_rng(0)
    G = rng.standard_normal((T, T + 3))
    slow = G @ G.T / T + 1e-6 * np.eye(T)       # ill-conditioned: needs many more iterations than the posteriors
    bad = covs[1].copy()
    bad[5, 5] = -1.0
    A = np.stack([covs[0], bad, slow, covs[2]])
    out, _, iters, info = backend.sqrtm_psd(_t(backend, A))
    info = info.cpu().numpy()
    assert info[1] != 0 and info[0] == 0 and info[2] == 0 and info[3] == 0
    for b in (0, 2, 3):
        got = out[b].cpu().numpy()
        assert rel_err(got, rp.sqrtm_svd(A[b])) < 1e-8, b
        assert rel_err(got @ got, A[b]) < 1e-10, b
    # alone, the slow problem takes the same number of iterations as inside the batch
    _, _, it_alone, _ = backend.sqrtm_psd(_t(backend, A[2:3]))
    assert iters == it_alone
    # an iteration cap that is too small is REPORTED, not silent
    _, _, it2, info2 = backend.sqrtm_psd(_t(backend, A[2:3]), max_iters=2)
    assert it2 == 2 and int(info2[0]) == 0x40000000


def test_w2_batch_reports_bad_pair_and_keeps_the_others(backend):
    T = 30
    mus, covs, _ = _posteriors(3, 4, T, seed=9)
    bad = covs[1].copy()
    bad[3, 3] = -2.0
    s1 = np.stack([covs[0], bad, covs[2]])
    s2 = np.stack([covs[2], covs[0], covs[0]])
    m1, m2 = mus, mus[[2, 0, 0]]
    w2, info = backend.w2_distance(_t(backend, m1), _t(backend, s1), _t(backend, m2), _t(backend, s2))
    info = info.cpu().numpy()
    assert info[1] != 0 and info[0] == 0 and info[2] == 0
    for p in (0, 2):
        want = rp.gaussian_w2_distance(m1[p], s1[p], m2[p], s2[p])
        assert abs(float(w2[p]) - want) < 1e-9 * max(1.0, abs(want))


@pytest.mark.parametrize("T,N", [(1, 3), (33, 5), (165, 4), (300, 2)])
def test_mvn_log_prob_general_vectors(backend, T, N):
    """distrax MultivariateNormalTri.log_prob on general vectors: one forward substitution against the stored factor."""
    mus, covs, _ = _posteriors(1, 4, T, seed=T)
    L = np.linalg.cholesky(covs[0])
    rng = np.random.default_rng(T)
    x = mus[0] + 0.2 * rng.standard_normal((N, T))
    ll = backend.mvn_log_prob(_t(backend, mus[0]), _t(backend, L), _t(backend, x), float(np.log(np.diag(L)).sum()))
    want = rp.mvn_log_prob(mus[0], L, x)
    assert rel_err(ll.cpu().numpy(), want) < 1e-12
    from bayesian_ensembling_b200 import dists

    d = dists.MultivariateNormalFullCovariance(mus[0], covs[0])
    assert rel_err(d.log_prob(x), want) < 1e-11
    if T > 1:
        # and the constant-vector path (weights.py:98-100) is the same density at o * 1
        o = np.array([0.3, -0.1])
        assert rel_err(d.log_prob(o[:, None]), rp.mvn_log_prob(mus[0], L, np.outer(o, np.ones(T)))) < 1e-11


def test_loglik_weight_accepts_members_without_constvec_statistics(backend):
    """A MultivariateNormalDiag member (a Barycentre output, a diag checkpoint) goes through the generic
    log_prob path of weights.py:93-104 instead of raising."""
    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200 import dists
    from bayesian_ensembling_b200.labelled import DataArray

    M, R, T, Ro = 3, 3, 24, 2
    reals, obs = _cell(M, R, T, Ro, seed=5)
    time = np.arange(T)
    pms = [es.ProcessModel(DataArray(reals[m], ("realisation", "time"), {"realisation": np.arange(R), "time": time}),
                           f"model{m}") for m in range(M)]
    obs_pm = es.ProcessModel(DataArray(obs, ("realisation", "time"), {"realisation": np.arange(Ro), "time": time}), "obs")
    mc = es.ModelCollection(pms)
    mc.fit(es.GPDTW1D(), n_optim_nits=2, progress_bar=False)
    full = es.LogLikelihoodWeight()(mc, obs_pm, standardisation_constant=1e-3).values
    # replace member 1 by a diagonal distribution with the same moments
    d = mc[1].distribution
    blank = d.dim_array
    mc[1].distribution = es.Distribution(mu=d._dist.mean(), covariance=np.sqrt(d._dist.variance()), dim_array=blank,
                                         dist_type=dists.MultivariateNormalDiag)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mixed = es.LogLikelihoodWeight()(mc, obs_pm, standardisation_constant=1e-3).values
    assert mixed.shape == full.shape and np.abs(np.nansum(mixed, axis=0) - 1.0).max() < 1e-9
    # oracle: per member log_prob of the constant vectors, mean over realisations, exp, nansum-normalise
    mus = [np.asarray(m.distribution._dist.mean()) for m in mc]
    lls = []
    for k, m in enumerate(mc):
        dist = m.distribution._dist
        if k == 1:
            ll = [rp.mvn_diag_log_prob(mus[k], dist.stddev(), np.outer(o, np.ones(T))) for o in obs]
        else:
            ll = [rp.mvn_log_prob(mus[k], dist.scale_tri, o[:, None]) for o in obs]
        lls.append(np.mean(ll, axis=0))
    le = np.exp(1e-3 * np.asarray(lls))
    assert rel_err(mixed, le / le.sum(axis=0)) < 1e-9


# ------------------------------------------------------------------------------------ small-T member kernels
@pytest.mark.parametrize("T,R,M", [(1, 2, 2), (2, 3, 2), (7, 2, 3), (30, 3, 3), (31, 3, 2), (32, 3, 2), (33, 4, 2),
                                   (62, 3, 2), (86, 5, 3), (128, 4, 2), (165, 10, 4), (222, 3, 2), (223, 3, 2),
                                   (251, 10, 5), (254, 3, 2)])
def test_small_t_member_kernels_vs_oracle_and_blocked_path(backend, T, R, M, monkeypatch):
    """T + 2 <= 256 runs through the one-CTA-per-member kernels (small_posterior.cuh); BE_NO_SMALL_T sends the same
    call through the blocked path.  Both against the oracle (<= 1e-8, north star) and against each other."""
    reals, obs = _cell(M, R, T, 3, seed=4000 + T)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    var, ls = np.full(M, 0.5), np.full(M, 6.0)
    monkeypatch.delenv("BE_NO_SMALL_T", raising=False)
    small = backend.gp_posterior(X, ym, yv, var, ls)
    monkeypatch.setenv("BE_NO_SMALL_T", "1")
    blocked = backend.gp_posterior(X, ym, yv, var, ls)
    monkeypatch.delenv("BE_NO_SMALL_T", raising=False)
    assert int(small.info_fit.abs().sum()) == 0 and int(small.info_dist.abs().sum()) == 0
    worst = 0.0
    for m in range(M):
        Xo, yo, so = rp.gpdtw1d_inputs(reals[m])
        mu_o, cov_o = rp.gp_posterior_closed_form(Xo, yo, so, 0.5, 6.0)
        L_o = np.linalg.cholesky(cov_o)
        a = sla.solve_triangular(L_o, np.ones(T), lower=True)
        b = sla.solve_triangular(L_o, mu_o, lower=True)
        st_o = np.array([a @ a, a @ b, b @ b, np.log(np.diag(L_o)).sum()])
        for name, got, want, tol in (("mu", small.mu[m], mu_o, 1e-8), ("cov", small.cov[m], cov_o, 1e-8),
                                     ("scale_tri", small.scale_tri[m], L_o, 1e-8),
                                     ("var_diag", small.var_diag[m], np.diag(cov_o), 1e-8)):
            e = rel_err(got.cpu().numpy(), want)
            worst = max(worst, e)
            assert e <= tol, (T, m, name, e)
        st = small.mvn_stats[m].cpu().numpy()
        scale = np.array([st_o[0], np.sqrt(st_o[0] * st_o[2]), st_o[2], max(abs(st_o[3]), 1.0)])
        assert (np.abs(st - st_o) / scale).max() <= 1e-10, (T, m, st, st_o)
        assert np.array_equal(small.cov[m].cpu().numpy(), small.cov[m].cpu().numpy().T)
        assert np.abs(np.triu(small.scale_tri[m].cpu().numpy(), 1)).max() == 0.0
    for name in ("mu", "cov", "scale_tri", "var_diag", "mvn_stats"):
        x, y = getattr(small, name).cpu().numpy(), getattr(blocked, name).cpu().numpy()
        assert rel_err(x, y) <= 1e-10, (T, name, rel_err(x, y))
    print(f"small-T kernels T={T}: worst rel err vs oracle {worst:.2e}")


def test_small_t_member_kernels_many_problems_and_non_pd_report(backend):
    """Thousands of CTAs (more than two waves), one of them non-positive-definite: its report is LAPACK's, the
    others are untouched; a repeated call returns bit-identical results."""
    T, M, C = 100, 5, 240
    cfg = synthetic.Config("t", 9, C, M, 3, T, 3, False, "")
    reals, obs = synthetic.make_cells(cfg, seed=99)
    r = _t(backend, reals.reshape(C * M, 3, T))
    X, ym, yv = backend.gpdtw1d_inputs(r)
    yv = yv.clone()
    bad = 777
    yv[bad, 40] = -5.0  # M[40, 40] = K + y_var + jitter < 0: leading minor of order 41 is not positive definite
    var, ls = np.full(C * M, 0.5), np.full(C * M, 6.0)
    a = backend.gp_posterior(X, ym, yv, var, ls, want_cov=False, want_scale_tri=False)
    b = backend.gp_posterior(X, ym, yv, var, ls, want_cov=False, want_scale_tri=False)
    info = a.info_fit.cpu().numpy()
    assert info[bad] == 41 and np.count_nonzero(info) == 1
    ok = np.arange(C * M) != bad
    assert int(a.info_dist.cpu().numpy()[ok].sum()) == 0
    for name in ("mu", "var_diag", "mvn_stats"):
        assert np.array_equal(getattr(a, name).cpu().numpy()[ok], getattr(b, name).cpu().numpy()[ok]), name
    for k in (0, 776, 778, C * M - 1):
        Xo, yo, so = rp.gpdtw1d_inputs(reals.reshape(C * M, 3, T)[k])
        mu_o, cov_o = rp.gp_posterior_closed_form(Xo, yo, so, 0.5, 6.0)
        assert rel_err(a.mu[k].cpu().numpy(), mu_o) <= 1e-8
        assert rel_err(a.var_diag[k].cpu().numpy(), np.diag(cov_o)) <= 1e-8


def test_crps_kernel_table_step_against_exact_values(backend):
    """k_crps_weights evaluates z (2 Phi(z) - 1) + 2 phi(z) by a Taylor step from a 513-entry table
    (weights_next_kernels.cuh).  Swept over |z| from 1e-12 to beyond the table's end -- on the grid points, at the
    half-way points where the nearest grid point changes, and at random -- against a 50-digit evaluation and
    against the oracle's erf / exp form (properscoring.crps_gaussian, weights.py:469-471)."""
    import mpmath as mp

    mp.mp.dps = 50
    rng = np.random.default_rng(12)
    k = np.arange(0, 600)
    z = np.concatenate([k / 64.0, k / 64.0 + 1.0 / 128.0, np.nextafter(k / 64.0 + 1.0 / 128.0, 0.0),
                        rng.uniform(0.0, 8.5, 3000), 10.0 ** rng.uniform(-12, 0, 500), [8.0, np.nextafter(8.0, 0.0), 37.5, 1e6]])
    z = np.concatenate([z, -z])
    N = z.size
    loc = np.zeros((1, 2, N))
    scale = np.ones((1, 2, N))
    _, cm = backend.crps_weights(_t(backend, loc), _t(backend, scale), _t(backend, z[None, None, :]), want_crps=True)
    got = cm[0, 0].cpu().numpy()
    assert rel_err(got, rp.crps_gaussian(z, 0.0, 1.0)) < 2e-15
    worst = 0.0
    for zi, gi in zip(z[::7], got[::7]):
        zz = mp.mpf(float(zi))
        G = zz * mp.erf(zz / mp.sqrt(2)) + mp.sqrt(2 / mp.pi) * mp.exp(-zz * zz / 2)
        worst = max(worst, float(abs(mp.mpf(float(gi)) - (G - 1 / mp.sqrt(mp.pi))) / G))
    assert worst < 5e-16, worst  # G to 2.9e-16 (tools/make_crps_table.py --check) and the rounding of the subtraction
    # infinities and NaN come out as the closed form gives them
    zs = np.array([np.inf, -np.inf, np.nan, 0.0])
    _, cm = backend.crps_weights(_t(backend, np.zeros((1, 2, 4))), _t(backend, np.ones((1, 2, 4))),
                                 _t(backend, zs[None, None, :]), want_crps=True)
    got = cm[0, 0].cpu().numpy()
    assert got[0] == np.inf and got[1] == np.inf and np.isnan(got[2])
    assert abs(got[3] - (np.sqrt(2.0) - 1.0) / np.sqrt(np.pi)) < 1e-16


def test_fit_batch_in_two_halves_equals_one_batch(backend):
    """GPDTW1D.fit_batch fits a group whose covariances exceed 256 MB in two halves (the first half's device-to-host
    copy runs under the second half's kernels).  Every member's host mean / covariance must be bit-identical to the
    one-batch device result, in the order of the collection."""
    import torch

    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200.labelled import DataArray

    M, R, T = 8, 3, 2100  # 8 x 2100^2 x 8 B = 282 MB: the two-halves path
    reals, _ = _cell(M, R, T, 2, seed=77)
    tcoord = np.arange(T)
    pms = [es.ProcessModel(DataArray(reals[m], ("realisation", "time"), {"realisation": np.arange(R), "time": tcoord}),
                           f"model{m}") for m in range(M)]
    mc = es.ModelCollection(pms)
    mc.fit(es.GPDTW1D(hyperparameters=(0.5, 6.0), y_mean="mean"), progress_bar=False)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    var = torch.full((M,), 0.5, dtype=torch.float64, device=backend.device)
    ls = torch.full((M,), 6.0, dtype=torch.float64, device=backend.device)
    ref = backend.gp_posterior(X, ym, yv, var, ls)
    assert int(ref.info_fit.abs().sum()) == 0
    mu_ref, cov_ref = ref.mu.cpu().numpy(), ref.cov.cpu().numpy()
    for m in range(M):
        d = mc[m].distribution
        assert np.array_equal(np.asarray(d.mu), mu_ref[m]), m
        assert np.array_equal(np.asarray(d.covariance), cov_ref[m]), m
        assert np.array_equal(d.mean.values, mu_ref[m])


@pytest.mark.parametrize("T", [61, 165, 251])
def test_l2_loop_small_t_factor_kernel_vs_blocked_path(backend, monkeypatch, T):
    """At T <= 254 the two factor-and-invert steps of an L2 iteration run in k_small_factor_inverse and the loop's
    padded dimension is the small-T kernels' (a multiple of 32); BE_NO_SMALL_T sends the same call through the blocked
    path (padded to a multiple of 16).  Same trained hyper-parameters and posterior to rounding."""
    reals, _ = _cell(3, 4, T, 2, seed=300 + T)
    X, ym, yv = backend.gpdtw1d_inputs(_t(backend, reals))
    post_a, var_a, ls_a = backend.vgp_fit(X, ym, yv, 8)
    monkeypatch.setenv("BE_NO_SMALL_T", "1")
    post_b, var_b, ls_b = backend.vgp_fit(X, ym, yv, 8)
    monkeypatch.delenv("BE_NO_SMALL_T")
    assert int(post_a.info_fit.abs().sum()) == 0 and int(post_b.info_fit.abs().sum()) == 0
    assert rel_err(var_a.cpu().numpy(), var_b.cpu().numpy()) < 1e-10
    assert rel_err(ls_a.cpu().numpy(), ls_b.cpu().numpy()) < 1e-10
    assert rel_err(post_a.mu.cpu().numpy(), post_b.mu.cpu().numpy()) < 1e-9
    assert rel_err(post_a.cov.cpu().numpy(), post_b.cov.cpu().numpy()) < 1e-9
