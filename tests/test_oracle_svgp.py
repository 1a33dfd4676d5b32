"""The SVGP oracle (oracle/svgp.py, the GPDTW3D stage of ensembles/models.py:357-411) checked against itself on the
CPU: the analytic ELBO gradient (kernel parameters AND inducing inputs) against central finite differences, the
natural-gradient closed form against stationarity of the ELBO, the documented minibatch order, and a short fit."""
import numpy as np

from oracle import svgp


def _problem(N=60, M=9, R=2, seed=3):
    rng = np.random.default_rng(seed)
    lat = rng.uniform(-60, 60, N)
    lon = rng.uniform(0, 360, N)
    X = np.column_stack([np.cos(np.radians(lat)) * np.cos(np.radians(lon)), np.cos(np.radians(lat)) * np.sin(np.radians(lon)),
                         np.sin(np.radians(lat)), rng.uniform(-1, 1, N), rng.normal(0.5, 0.4, (N, R))])
    y = 0.3 * X[:, 3] + 0.2 * X[:, 2] + 0.05 * rng.standard_normal(N)
    s = rng.uniform(0.01, 0.05, N)
    Z = svgp.inducing_points(X, M) + 0.01 * rng.standard_normal((M, X.shape[1]))
    var = rng.uniform(0.4, 1.5, 4)
    ls = rng.uniform(0.5, 2.0, 4)
    q_mu = 0.3 * rng.standard_normal(M)
    q_sqrt = np.tril(0.1 * rng.standard_normal((M, M))) + np.eye(M)
    return X, y, s, Z, var, ls, q_mu, q_sqrt, svgp.component_dims(R)


def test_elbo_gradient_matches_finite_differences():
    X, y, s, Z, var, ls, q_mu, q_sqrt, dims = _problem()
    g_var, g_ls, g_Z = svgp.elbo_grads(Z, X, y, s, var, ls, dims, q_mu, q_sqrt)
    f = lambda Z_, v_, l_: svgp.svgp_elbo(Z_, X, y, s, v_, l_, dims, q_mu, q_sqrt)  # noqa: E731
    h = 1e-6
    for c in range(4):
        e = np.zeros(4)
        e[c] = h
        fd_v = (f(Z, var + e, ls) - f(Z, var - e, ls)) / (2 * h)
        fd_l = (f(Z, var, ls + e) - f(Z, var, ls - e)) / (2 * h)
        assert abs(fd_v - g_var[c]) <= 1e-6 * max(1.0, abs(g_var[c])), (c, fd_v, g_var[c])
        assert abs(fd_l - g_ls[c]) <= 1e-6 * max(1.0, abs(g_ls[c])), (c, fd_l, g_ls[c])
    rng = np.random.default_rng(0)
    for _ in range(25):
        m, d = rng.integers(Z.shape[0]), rng.integers(Z.shape[1])
        E = np.zeros_like(Z)
        E[m, d] = h
        fd = (f(Z + E, var, ls) - f(Z - E, var, ls)) / (2 * h)
        assert abs(fd - g_Z[m, d]) <= 2e-6 * max(1.0, abs(g_Z[m, d])), (m, d, fd, g_Z[m, d])


def test_natgrad_step_with_gamma_one_is_the_minibatch_optimum():
    """For a Gaussian likelihood one natural-gradient step with gamma = 1 lands on the optimal q of that minibatch: the
    ELBO is stationary in q_mu there, and gamma = 0.5 moves exactly half way in natural parameters."""
    X, y, s, Z, var, ls, q_mu, q_sqrt, dims = _problem(seed=5)
    _, _, A, _ = svgp.conditional(Z, X, var, ls, dims, q_mu, q_sqrt)
    mu1, sq1 = svgp.natgrad_step(A, y, s, q_mu, q_sqrt, gamma=1.0)
    h = 1e-6
    for m in range(Z.shape[0]):
        e = np.zeros(Z.shape[0])
        e[m] = h
        fd = (svgp.svgp_elbo(Z, X, y, s, var, ls, dims, mu1 + e, sq1) - svgp.svgp_elbo(Z, X, y, s, var, ls, dims, mu1 - e, sq1)) / (2 * h)
        assert abs(fd) < 1e-5, (m, fd)
    assert svgp.svgp_elbo(Z, X, y, s, var, ls, dims, mu1, sq1) >= svgp.svgp_elbo(Z, X, y, s, var, ls, dims, q_mu, q_sqrt)
    muh, sqh = svgp.natgrad_step(A, y, s, q_mu, q_sqrt, gamma=0.5)
    P0, P1, Ph = (np.linalg.inv(q @ q.T) for q in (np.tril(q_sqrt), sq1, sqh))
    assert np.abs(Ph - 0.5 * (P0 + P1)).max() < 1e-9 * np.abs(P1).max()
    assert np.abs(Ph @ muh - 0.5 * (P0 @ q_mu + P1 @ mu1)).max() < 1e-9 * np.abs(P1 @ mu1).max()


def test_minibatch_order_is_reproducible_and_covers_epochs():
    a = svgp.batch_indices(103, 10, 25, seed=7)
    b = svgp.batch_indices(103, 10, 25, seed=7)
    assert np.array_equal(a, b) and a.shape == (25, 10)
    assert sorted(a.ravel()[:103].tolist()) == list(range(103))  # the first epoch is a permutation
    assert not np.array_equal(a, svgp.batch_indices(103, 10, 25, seed=8))


def test_short_fit_improves_the_elbo_and_returns_the_reference_shapes():
    X, y, s, *_ = _problem(N=80, R=2, seed=9)
    Y = np.column_stack([y, s])
    mu0, v0, st0 = svgp.svgp_fit(X, Y, 0, n_inducing=8, minibatch_size=20, seed=1, return_state=True)
    mu, v, st = svgp.svgp_fit(X, Y, 30, n_inducing=8, minibatch_size=20, seed=1, return_state=True)
    assert mu.shape == (80,) and v.shape == (80,) and (v > s).all()  # var + Y[:, 1], models.py:411
    dims = svgp.component_dims(2)
    e0 = svgp.svgp_elbo(st0["Z"], X, y, s, st0["variances"], st0["lengthscales"], dims, st0["q_mu"], st0["q_sqrt"])
    e1 = svgp.svgp_elbo(st["Z"], X, y, s, st["variances"], st["lengthscales"], dims, st["q_mu"], st["q_sqrt"])
    assert e1 > e0
    assert np.mean((mu - y) ** 2) < np.mean((mu0 - y) ** 2)
