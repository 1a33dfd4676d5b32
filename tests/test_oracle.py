"""Self-consistency of the CPU oracle (it restates GPflow/distrax arithmetic that cannot run
here, so its pieces are cross-checked against each other and against brute force).  CPU only."""
import numpy as np
import pytest
import scipy.linalg as sla

from bayesian_ensembling_b200 import synthetic
from oracle import reference_path as rp
from conftest import rel_err


def _member(T=40, R=4, seed=3):
    cfg = synthetic.Config("t", 9, 1, 1, R, T, 3, False, "")
    reals, obs = synthetic.make_cells(cfg, seed=seed)
    return reals[0, 0], obs[0]


def test_matern_gram_matches_direct_distances():
    reals, _ = _member()
    X, _, _ = rp.gpdtw1d_inputs(reals)
    K = rp.matern32_gram(X, 0.7, 3.0)
    d = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)) / 3.0
    want = 0.7 * (1 + np.sqrt(3) * d) * np.exp(-np.sqrt(3) * d)
    assert np.abs(K - want).max() < 1e-12
    assert np.allclose(np.diag(K), 0.7, atol=1e-12)


def test_y_var_is_population_variance():
    reals, _ = _member()
    _, y, s = rp.gpdtw1d_inputs(reals)
    assert np.allclose(s, reals.var(axis=0, ddof=0))
    assert np.allclose(y, reals.mean(axis=0))


def test_natgrad_fixed_point_is_closed_form():
    """gamma=1 lands on the optimum in one step; gamma=0.5 converges geometrically to it."""
    reals, _ = _member(T=30)
    X, y, s = rp.gpdtw1d_inputs(reals)
    var, ls = 0.5, 6.0
    mean_cf, cov_cf = rp.gp_posterior_closed_form(X, y, s, var, ls)
    L = np.linalg.cholesky(rp.matern32_gram(X, var, ls) + rp.DEFAULT_JITTER * np.eye(30))
    q_mu, q_sqrt = rp.vgp_natgrad_step(L, y, s, np.zeros(30), np.eye(30), gamma=1.0)
    m, c = rp.vgp_predict_full_cov(X, var, ls, q_mu, q_sqrt)
    # the variational posterior conditions on K + jitter*I as the prior; the closed form uses
    # the same jitter inside the solve => agreement to O(jitter^2)
    assert rel_err(m, mean_cf) < 1e-7
    assert rel_err(c + np.diag(s), cov_cf) < 1e-6
    q_mu, q_sqrt = np.zeros(30), np.eye(30)
    for _ in range(60):
        q_mu, q_sqrt = rp.vgp_natgrad_step(L, y, s, q_mu, q_sqrt, gamma=0.5)
    m2, c2 = rp.vgp_predict_full_cov(X, var, ls, q_mu, q_sqrt)
    assert rel_err(m2, m) < 1e-9 and rel_err(c2, c) < 1e-9


def test_natgrad_k_steps_scale_noise_precision():
    """after k steps at fixed theta the data term carries 1 - 0.5^k."""
    reals, _ = _member(T=25)
    X, y, s = rp.gpdtw1d_inputs(reals)
    L = np.linalg.cholesky(rp.matern32_gram(X, 0.5, 6.0) + rp.DEFAULT_JITTER * np.eye(25))
    q_mu, q_sqrt = np.zeros(25), np.eye(25)
    for _ in range(2):
        q_mu, q_sqrt = rp.vgp_natgrad_step(L, y, s, q_mu, q_sqrt, gamma=0.5)
    q1, s1 = rp.vgp_natgrad_step(L, y, s / 0.75, np.zeros(25), np.eye(25), gamma=1.0)
    assert rel_err(q_mu, q1) < 1e-10
    assert rel_err(q_sqrt @ q_sqrt.T, s1 @ s1.T) < 1e-10


def test_hyper_gradient_matches_finite_differences():
    reals, _ = _member(T=20)
    X, y, s = rp.gpdtw1d_inputs(reals)
    rng = np.random.default_rng(0)
    q_mu = rng.normal(size=20)
    q_sqrt = np.tril(rng.normal(size=(20, 20)) * 0.1) + np.eye(20)
    u = np.array([rp.softplus_inv(0.8), rp.softplus_inv(2.5)])
    g = rp.vgp_hyper_grad(X, y, s, rp.softplus(u[0]), rp.softplus(u[1]), q_mu, q_sqrt)
    for k in range(2):
        h = 1e-6
        up, um = u.copy(), u.copy()
        up[k] += h
        um[k] -= h
        fp = -rp.vgp_elbo(X, y, s, rp.softplus(up[0]), rp.softplus(up[1]), q_mu, q_sqrt)
        fm = -rp.vgp_elbo(X, y, s, rp.softplus(um[0]), rp.softplus(um[1]), q_mu, q_sqrt)
        fd = (fp - fm) / (2 * h)
        assert abs(fd - g[k]) <= 1e-5 * max(1.0, abs(fd)), (k, fd, g[k])


def test_gpdtw1d_fit_improves_elbo_and_is_deterministic():
    reals, _ = _member(T=24, R=3)
    mu_a, cov_a, st_a = rp.gpdtw1d_fit(reals, n_optim_nits=15, return_state=True)
    mu_b, cov_b = rp.gpdtw1d_fit(reals, n_optim_nits=15)
    assert np.array_equal(mu_a, mu_b) and np.array_equal(cov_a, cov_b)
    X, y, s = rp.gpdtw1d_inputs(reals)
    e0 = rp.vgp_elbo(X, y, s, 1.0, 1.0, np.zeros(24), np.eye(24))
    e1 = rp.vgp_elbo(X, y, s, st_a["variance"], st_a["lengthscale"], st_a["q_mu"], st_a["q_sqrt"])
    assert e1 > e0
    assert np.linalg.eigvalsh(cov_a).min() > 0


def test_constant_vector_logprob_quirk():
    """weights.py:98-100: [T,1] input broadcasts to [T,T]; entry i is the density of o_i * 1."""
    rng = np.random.default_rng(1)
    T = 12
    A = rng.normal(size=(T, T))
    cov = A @ A.T + T * np.eye(T)
    mu = rng.normal(size=T)
    L = np.linalg.cholesky(cov)
    o = rng.normal(size=T)
    ll = rp.mvn_log_prob(mu, L, o[:, None])
    assert ll.shape == (T,)
    for i in range(T):
        d = o[i] * np.ones(T) - mu
        want = -0.5 * d @ np.linalg.solve(cov, d) - 0.5 * T * np.log(2 * np.pi) - 0.5 * np.linalg.slogdet(cov)[1]
        assert abs(ll[i] - want) < 1e-10
    # and a genuine [T] vector gives the ordinary joint density (scalar)
    assert np.ndim(rp.mvn_log_prob(mu, L, o)) == 0


@pytest.mark.parametrize("M,Ro", [(2, 1), (5, 2), (10, 5), (10, 10)])
def test_weights_shape_and_normalisation(M, Ro):
    """The reference's own assertion set (tests/test_weights.py:99-101) on its test shapes."""
    cfg = synthetic.Config("t", 9, 1, M, 3, 24, Ro, True, "")
    reals, obs = synthetic.make_cells(cfg, seed=11)
    out = rp.cell_pipeline_L1(reals[0], obs[0], 0.5, 6.0)
    w = out["weights"]
    assert w.shape == (M, 24)
    ok = ~np.isnan(w).any(axis=0)
    assert ok.any()
    assert np.allclose(w[:, ok].sum(axis=0), 1.0, atol=1e-6)


def test_normal_branch_scale_quirk():
    loc, scale, x = 0.3, 0.04, 0.35
    want = -0.5 * ((x - loc) / scale) ** 2 - 0.5 * np.log(2 * np.pi) - np.log(scale)
    assert abs(rp.normal_log_prob(loc, scale, x) - want) < 1e-14
    w, e, m = rp.loglik_weights_normal(np.array([[0.0, 1.0], [0.5, 1.5]]), np.array([[1.0, 1.0], [2.0, 2.0]]),
                                       np.array([[0.1, 0.9]]))
    assert np.allclose(w.sum(axis=0), 1.0)


def test_barycentre_signed_stop_rule():
    """Q-BARY: S < 1 exits at iteration 0 with variance S (not S^2); S > 1 climbs to S^2."""
    mu, sd, it = rp.gaussian_barycentre([1.0, 3.0], [0.1, 0.3], [0.5, 0.5])
    assert it == 0 and abs(sd**2 - 0.2) < 1e-15 and abs(mu - 2.0) < 1e-15
    mu, sd, it = rp.gaussian_barycentre([0.0, 0.0], [2.0, 4.0], [0.5, 0.5])
    assert it > 5 and abs(sd**2 - 9.0) < 1e-4 and sd**2 <= 9.0
    # NaN weights (0/0 columns) propagate: NaN - x < tol is False => runs to the iteration cap
    with pytest.warns(UserWarning):
        mu, sd, it = rp.gaussian_barycentre([0.0, 0.0], [2.0, 4.0], [np.nan, np.nan])
    assert np.isnan(mu) and np.isnan(sd) and it == 201


def test_sqrtm_and_w2():
    rng = np.random.default_rng(2)
    A = rng.normal(size=(9, 9))
    S1 = A @ A.T + np.eye(9)
    B = rng.normal(size=(9, 9))
    S2 = B @ B.T + np.eye(9)
    R = rp.sqrtm_svd(S1)
    assert rel_err(R @ R, S1) < 1e-12 and rel_err(R, sla.sqrtm(S1).real) < 1e-10
    m1, m2 = rng.normal(size=9), rng.normal(size=9)
    assert abs(rp.gaussian_w2_distance(m1, S1, m1, S1)) < 1e-10
    d = rp.gaussian_w2_distance(m1, S1, m2, S2)
    # Q-W2: the location term is the UN-squared norm
    bures = np.trace(S1 + S2 - 2 * sla.sqrtm(R @ S2 @ R).real)
    assert abs(d - (np.linalg.norm(m1 - m2) + bures)) < 1e-9


def test_fullcov_barycentre_reduces_to_1d():
    for stds in ([0.1, 0.3, 0.2], [2.0, 4.0, 3.0]):
        w = [0.2, 0.5, 0.3]
        mus = [1.0, 2.0, 3.0]
        mu1, sd1, it1 = rp.gaussian_barycentre(mus, stds, w)
        mu, S, it = rp.fullcov_barycentre([[m] for m in mus], [np.array([[s * s]]) for s in stds], w)
        assert it == it1
        assert abs(S[0, 0] - sd1**2) < 1e-12 and abs(mu[0] - mu1) < 1e-14


def test_fullcov_barycentre_commuting_case():
    """Diagonal (commuting) covariances: the matrix iteration acts per coordinate."""
    d1, d2 = np.array([4.0, 9.0, 16.0]), np.array([1.0, 25.0, 4.0])
    w = [0.5, 0.5]
    _, S, _ = rp.fullcov_barycentre([np.zeros(3)] * 2, [np.diag(d1), np.diag(d2)], w, tolerance=1e-12, max_iters=500)
    want = (0.5 * np.sqrt(d1) + 0.5 * np.sqrt(d2)) ** 2
    assert np.allclose(np.diag(S), want, rtol=1e-4)
    assert np.abs(S - np.diag(np.diag(S))).max() < 1e-12


def test_crps_gaussian_known_answers():
    """properscoring.crps_gaussian closed form: CRPS of N(0,1) at its mean is (sqrt2 - 1)/sqrt(pi);
    against the defining integral int (F(y) - 1{y >= x})^2 dy elsewhere."""
    import math

    from scipy import integrate, special

    assert abs(rp.crps_gaussian(0.0, 0.0, 1.0) - (math.sqrt(2.0) - 1.0) / math.sqrt(math.pi)) < 1e-15
    for x, mu, s in ((0.7, 0.2, 1.3), (-2.0, 1.0, 0.4)):
        f = lambda y: (special.ndtr((y - mu) / s) - (y >= x)) ** 2  # noqa: E731
        want = integrate.quad(f, -40, x)[0] + integrate.quad(f, x, 40)[0]
        assert abs(rp.crps_gaussian(x, mu, s) - want) < 1e-9


def test_next_row_weights_normalise_over_models():
    rng = np.random.default_rng(11)
    M, T, Ro = 4, 9, 3
    means, var, obs = rng.normal(size=(M, T)), rng.uniform(0.1, 1.0, (M, T)), rng.normal(size=(Ro, T))
    for w in (rp.crps_weights(means, var, obs)[0], rp.model_similarity_weights_temporal(means, var)[0],
              rp.inverse_square_weights(means, obs.mean(axis=0))):
        assert w.shape == (M, T) and np.abs(w.sum(axis=0) - 1.0).max() < 1e-12
    # temporal similarity: identical members are at distance 0 of each other, so a duplicated member pair
    # gets a smaller raw mean distance than the odd one out
    means2 = np.stack([means[0], means[0], means[1]])
    var2 = np.stack([var[0], var[0], var[1]])
    w, w2 = rp.model_similarity_weights_temporal(means2, var2)
    assert np.abs(w2[0, 1]).max() < 1e-15 and (w[2] > w[0]).all()
    # single mode reduces to the pairwise a8 distances
    S = [np.diag(v) for v in var]
    ws, d = rp.model_similarity_weights_single(list(means), S)
    assert abs(ws.sum() - 1.0) < 1e-12 and np.abs(np.diag(d)).max() < 1e-12


def test_ksd_oracle_matches_literal_k0():
    """ksd_imq (vectorised) against a literal transcription of k_0_fun / imq_KSD (weights.py:360-394)."""
    def k0(p1, p2, g1, g2, c=1.0, beta=-0.5):
        diff = p1 - p2
        dim = p1.shape[0]
        q = c ** 2 + np.dot(diff, diff)
        return (np.dot(g1, g2) * q ** beta + -2 * beta * np.dot(g1, diff) * q ** (beta - 1)
                + 2 * beta * np.dot(g2, diff) * q ** (beta - 1) + -2 * dim * beta * q ** (beta - 1)
                + -4 * beta * (beta - 1) * q ** (beta - 2) * np.sum(np.square(diff)))

    rng = np.random.default_rng(2)
    x = rng.normal(size=(7, 1))
    mu, scale = 0.3, 0.4
    g = -(x - mu) / scale ** 2  # d/dx log N(x | mu, scale)
    tot = sum(k0(x[a], x[b], g[a], g[b]) for a in range(7) for b in range(7))
    assert abs(rp.ksd_imq(x, g) - np.sqrt(tot) / 7) < 1e-13 * np.sqrt(tot)
    w, k = rp.ksd_weights(np.full((2, 1), mu), np.array([[scale], [2 * scale]]), x)
    assert abs(k[0, 0] - np.sqrt(tot) / 7) < 1e-13 * np.sqrt(tot)
    assert abs(w.sum() - 1.0) < 1e-15
