"""The SVGP stage of GPDTW3D.fit (ensembles/models.py:357-411) on the device (be_svgp_fit, through
the C ABI) against the oracle (oracle/svgp.py) on the SAME seeded minibatch order.  Needs a B200: ``-m gpu``.

Tolerance: what limits agreement is the conditioning of Kuu, not the step count.  The reference places its inducing
inputs on ONE line through input space (linspace(min X, max X, M), models.py:370), so neighbouring inducing points are
nearly identical and cond(Kuu + 1e-6 I) is 4e4 at M = 24, 3e7 at M = 130 and 6e8 at M = 400 (the reference's own
default): two correct fp64 Cholesky factorisations of such a matrix differ by ~cond x 1e-16.  Observed against the
oracle: 1e-13 (M = 24), 4e-8 (M = 130), 9e-7 (M = 400).  Bars: 1e-9 / 1e-6 / 1e-5 by M."""
import numpy as np
import pytest

from oracle import svgp
from conftest import rel_err

pytestmark = pytest.mark.gpu


def _t(backend, a, dtype=None):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype or torch.float64, device=backend.device)


def _points(N, R, seed):
    rng = np.random.default_rng(seed)
    lat = rng.uniform(-75, 75, N)
    lon = rng.uniform(0, 360, N)
    t = rng.uniform(-1, 1, N)
    X = np.column_stack([np.cos(np.radians(lat)) * np.cos(np.radians(lon)), np.cos(np.radians(lat)) * np.sin(np.radians(lon)),
                         np.sin(np.radians(lat)), t, 0.4 * t[:, None] + 0.1 * rng.standard_normal((N, R))])
    y = 0.4 * t + 0.2 * np.sin(np.radians(lat)) + 0.05 * rng.standard_normal(N)
    s = rng.uniform(0.005, 0.03, N)
    return X, np.column_stack([y, s])


@pytest.mark.parametrize("N,R,M,batch,n_steps,train", [(300, 3, 24, 50, 6, False), (300, 3, 24, 50, 12, True),
                                                       (900, 5, 130, 200, 4, True), (1500, 10, 400, 500, 2, True)])
def test_svgp_fit_vs_oracle(backend, N, R, M, batch, n_steps, train):
    X, Y = _points(N, R, seed=N + M)
    Z0 = svgp.inducing_points(X, M)
    idx = svgp.batch_indices(N, batch, 2 * n_steps, seed=11)
    out = backend.svgp_fit(_t(backend, X), _t(backend, Y), _t(backend, Z0), idx, n_steps, train_hypers=train, predict_chunk=256)
    assert int(out["info"].item()) == 0
    mu_o, var_o, st = svgp.svgp_fit(X, Y, n_steps, n_inducing=M, minibatch_size=batch, seed=11, train_hypers=train,
                                    return_state=True)
    tol = 1e-9 if M <= 24 else (1e-6 if M <= 130 else 1e-5)
    errs = dict(
        mu=rel_err(out["mu"].cpu().numpy(), mu_o), var=rel_err(out["var"].cpu().numpy(), var_o),
        q_mu=rel_err(out["q_mu"].cpu().numpy(), st["q_mu"]), q_sqrt=rel_err(out["q_sqrt"].cpu().numpy(), np.tril(st["q_sqrt"])),
        variances=rel_err(out["variances"].cpu().numpy(), st["variances"]),
        lengthscales=rel_err(out["lengthscales"].cpu().numpy(), st["lengthscales"]), Z=rel_err(out["Z"].cpu().numpy(), st["Z"]))
    print(f"svgp N={N} M={M} steps={n_steps} train={train}:", {k: f"{v:.1e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v <= tol, (k, v)
    if not train:
        assert np.array_equal(out["Z"].cpu().numpy(), Z0)


def test_gpdtw3d_fit_end_to_end(backend):
    """GPDTW3D.fit through the reference-shaped API: a Distribution of dx.Normal over (time, latitude, longitude) whose
    loc / scale are the oracle's SVGP prediction on the same DBA means and the same minibatch order."""
    import warnings

    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200 import dists
    from bayesian_ensembling_b200.labelled import DataArray

    rng = np.random.default_rng(5)
    R, T, n_lat, n_lon = 3, 12, 4, 5
    t = np.linspace(0, 1, T)
    data = (0.8 * t[None, :, None, None] + 0.3 * np.sin(np.linspace(0, 3, n_lat))[None, None, :, None]
            + 0.08 * rng.standard_normal((R, T, n_lat, n_lon)))
    lat, lon = np.linspace(-60.0, 60.0, n_lat), np.linspace(0.0, 288.0, n_lon)
    pm = es.ProcessModel(DataArray(data, ("realisation", "time", "latitude", "longitude"),
                                   {"realisation": np.arange(R), "time": 1990 + np.arange(T), "latitude": lat,
                                    "longitude": lon}, name="tas"), "m")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g3 = es.GPDTW3D()
        dist = g3.fit(pm, n_optim_nits=2, n_inducing=20, minibatch_size=60, seed=3)
    assert dist.dist_type is dists.Normal
    assert dist.mean.shape == (T, n_lat, n_lon) and dist.mean.dims == ("time", "latitude", "longitude")
    mean_array, var_array = g3._dtw_to_xarray(pm)
    X, Y = g3._prep_data(pm.model_data, mean_array, var_array)
    N = X.shape[0]
    n_steps = 2 * (N // 60)
    assert np.array_equal(es.GPDTW3D.minibatch_order(N, 60, 2 * n_steps, 3), svgp.batch_indices(N, 60, 2 * n_steps, 3))
    mu_o, var_o = svgp.svgp_fit(X, Y, n_steps, n_inducing=20, minibatch_size=60, seed=3)
    assert rel_err(dist._dist.mean().ravel(), mu_o) <= 1e-6
    assert rel_err(dist._dist.stddev().ravel(), var_o) <= 1e-6  # dx.Normal(mu, cov): the variance is the SCALE (Q-SCALE)
    assert (dist._dist.stddev().ravel() > Y[:, 1]).all()
