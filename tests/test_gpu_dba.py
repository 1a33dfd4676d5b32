"""Parity of the device DTW-barycentre-averaging path through the C ABI:
against the reference's own ensembles/dtwa.py outputs (tests/golden/dba_reference.npz), against the
C oracle (oracle/dba.c) on seeded inputs for every (columns-per-thread, warps-per-pair) kernel
shape, and through size-independent properties at BASELINE sizes.  Needs a B200: ``-m gpu``.

The DTW table is integer-like work in fp64: every cell is one subtract, one multiply, two
compares and one add, rounded identically on both sides, so squared distances, paths, iteration
counts and the barycentres are compared BIT-EXACTLY unless a test says otherwise.
"""
import os

import numpy as np
import pytest

from bayesian_ensembling_b200 import synthetic
from oracle import dba
from oracle import reference_path as rp
from conftest import rel_err

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dba_reference.npz")
# one T per kernel shape of be_dtw_api.cuh:dtw_shape, plus ragged sizes around the boundaries
SHAPE_TS = [1, 2, 7, 32, 33, 64, 100, 128, 251, 257, 512, 513, 700, 1025, 1980, 2049, 3012, 3136, 3137]


def _t(backend, a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=backend.device)


def _series(rng, R, T, kind="gmst"):
    t = np.linspace(0.0, 1.0, max(T, 2))[:T]
    if kind == "gmst":
        g = rng.uniform(0.5, 4.0) * t + rng.uniform(0.0, 2.0) * t * t
        e = np.zeros((R, T))
        z = rng.normal(0.0, 0.12, (R, T))
        for i in range(1, T):
            e[:, i] = 0.6 * e[:, i - 1] + z[:, i]
        return g[None] + e
    if kind == "shifted":
        c = 0.3 + 0.4 * rng.random(R)
        return np.exp(-0.5 * ((t[None] - c[:, None]) / 0.08) ** 2) + 0.02 * rng.normal(size=(R, T))
    return rng.integers(0, 3, size=(R, T)).astype(np.float64)  # "ties"


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_dtw_squared_golden_bit_exact(backend, golden):
    for c in range(int(golden["n_cases"])):
        X = golden[f"c{c}_X"]
        R = X.shape[0]
        a = np.repeat(X, R, axis=0)
        x = np.tile(X, (R, 1))
        sq = backend.dtw_squared(_t(backend, a), _t(backend, x)).cpu().numpy().reshape(R, R)
        assert np.array_equal(sq, golden[f"c{c}_sqdtw"]), c


@pytest.mark.parametrize("T", SHAPE_TS)
def test_dtw_squared_vs_oracle_every_kernel_shape(backend, T):
    rng = np.random.default_rng(1000 + T)
    P = 5 if T > 1000 else 9
    a = _series(rng, P, T, "gmst")
    x = _series(rng, P, T, "shifted" if T % 2 else "gmst")
    got = backend.dtw_squared(_t(backend, a), _t(backend, x)).cpu().numpy()
    want = np.array([dba.squared_dtw(a[p], x[p]) for p in range(P)])
    assert np.array_equal(got, want)


def test_perform_dba_golden(backend, golden):
    """ensembles/dtwa.py:6-20 executed by the reference itself: medoid exact, centres <= 1e-15, and
    bit-exact on the integer-valued cases (which exercise the tie rule of dtwa.py:113-129)."""
    for c in range(int(golden["n_cases"])):
        X = golden[f"c{c}_X"]
        for n in (1, 3, 10):
            cen, med = backend.perform_dba(_t(backend, X[None]), n_iterations=n, want_medoid=True)
            ref = golden[f"c{c}_center{n}"]
            assert int(med[0]) == int(golden[f"c{c}_medoid"])
            cen = cen[0].cpu().numpy()
            assert np.abs(cen - ref).max() <= 1e-15 * np.abs(ref).max(), (c, n)
            if str(golden[f"c{c}_kind"]) == "ties":
                assert np.array_equal(cen, ref)


@pytest.mark.parametrize("R,T,kind", [(3, 24, "ties"), (4, 100, "shifted"), (5, 251, "gmst"), (6, 300, "ties"),
                                      (3, 700, "shifted"), (5, 1100, "gmst")])
def test_perform_dba_vs_oracle_bit_exact(backend, R, T, kind):
    rng = np.random.default_rng(T * 7 + R)
    B = 3
    X = np.stack([_series(rng, R, T, kind) for _ in range(B)])
    cen, med = backend.perform_dba(_t(backend, X), n_iterations=4, want_medoid=True)
    for b in range(B):
        want, wmed = dba.perform_dba(X[b], 4)
        assert int(med[b]) == wmed
        assert np.array_equal(cen[b].cpu().numpy(), want), (b, np.abs(cen[b].cpu().numpy() - want).max())


@pytest.mark.parametrize("R,T,kind,max_iter,tol", [
    (1, 1, "gmst", 5, 1e-5), (2, 2, "gmst", 5, 1e-5), (3, 24, "ties", 50, 1e-3), (3, 33, "shifted", 50, 1e-6),
    (5, 165, "gmst", 50, 1e-3), (10, 251, "gmst", 50, 1e-3), (4, 300, "shifted", 30, 1e-5), (5, 600, "shifted", 50, 1e-3),
    (3, 1500, "gmst", 20, 1e-3), (5, 1980, "gmst", 12, 1e-3), (5, 3012, "gmst", 6, 1e-3),
])
def test_dba_subgradient_vs_oracle_bit_exact(backend, R, T, kind, max_iter, tol):
    """tslearn's dtw_barycenter_averaging_subgradient as models.py:176-178 calls it (max_iter=50,
    tol=1e-3) and at other settings: barycentre, iteration count and last cost equal the oracle's."""
    rng = np.random.default_rng(T * 13 + R)
    B = 3 if T <= 700 else 2
    X = np.stack([_series(rng, R, T, kind) for _ in range(B)])
    bary, n_iter, cost = backend.dtw_barycenter_averaging_subgradient(_t(backend, X), max_iter=max_iter, tol=tol,
                                                                      want_info=True)
    for b in range(B):
        want, n, c = dba.dba_subgradient(X[b], max_iter=max_iter, tol=tol)
        got = bary[b].cpu().numpy()
        assert int(n_iter[b]) == n, (b, int(n_iter[b]), n)
        assert np.array_equal(got, want), (b, np.abs(got - want).max())
        assert float(cost[b]) == c


def test_dba_problems_stop_independently(backend):
    """Problems of one batch converge at different iterations; a finished problem must not move."""
    rng = np.random.default_rng(77)
    T, R = 120, 4
    x = np.cumsum(rng.normal(size=T))
    X = np.stack([np.stack([x] * R), _series(rng, R, T, "shifted"), _series(rng, R, T, "gmst"),
                  _series(rng, R, T, "ties")])
    bary, n_iter, _ = backend.dtw_barycenter_averaging_subgradient(_t(backend, X), max_iter=50, tol=1e-3, want_info=True)
    its = [int(v) for v in n_iter]
    assert its[0] == 2 and len(set(its)) > 1
    for b in range(4):
        want, n, _ = dba.dba_subgradient(X[b], max_iter=50, tol=1e-3)
        assert its[b] == n and np.array_equal(bary[b].cpu().numpy(), want)
    # the same problem alone and inside a batch: bit-identical
    alone = backend.dtw_barycenter_averaging_subgradient(_t(backend, X[2:3]), max_iter=50, tol=1e-3)
    assert np.array_equal(alone[0].cpu().numpy(), bary[2].cpu().numpy())


def test_dba_init_barycenter_and_zero_iterations(backend):
    rng = np.random.default_rng(3)
    X = np.stack([_series(rng, 4, 90, "shifted") for _ in range(2)])
    b0 = backend.dtw_barycenter_averaging_subgradient(_t(backend, X), max_iter=0).cpu().numpy()
    assert rel_err(b0, X.mean(axis=1)) < 1e-15
    init = X[:, 1, :]
    bi = backend.dtw_barycenter_averaging_subgradient(_t(backend, X), max_iter=7, tol=0.0,
                                                      init_barycenter=_t(backend, init)).cpu().numpy()
    for b in range(2):
        want, n, _ = dba.dba_subgradient(X[b], max_iter=7, tol=0.0, init_barycenter=init[b])
        assert n == 7 and np.array_equal(bi[b], want)


def test_dba_properties_at_cfg2_size(backend):
    """BASELINE configs[1] shape (R = 5, T = 3012), sizes the C oracle would need minutes for:
    (i) identical realisations are a fixed point; (ii) adding a constant to every realisation adds
    it to the barycentre; (iii) the DBA cost of the result is below the cost of the arithmetic mean."""
    import torch

    cfg = synthetic.Config("t", 9, 1, 4, 5, 3012, 2, True, "")
    reals, _ = synthetic.make_cells(cfg, seed=99)
    X = _t(backend, reals[0])  # [4, 5, 3012]
    same = X[:, :1, :].expand(-1, 2, -1).contiguous()
    b_same, n_same, _ = backend.dtw_barycenter_averaging_subgradient(same, max_iter=50, tol=1e-3, want_info=True)
    assert torch.equal(b_same, same[:, 0, :]) and [int(v) for v in n_same] == [2] * 4
    bary, n_iter, cost = backend.dtw_barycenter_averaging_subgradient(X, max_iter=50, tol=1e-3, want_info=True)
    assert all(1 <= int(v) <= 50 for v in n_iter)
    shifted = backend.dtw_barycenter_averaging_subgradient(X + 3.0, max_iter=50, tol=1e-3)
    assert float((shifted - 3.0 - bary).abs().max()) < 1e-9
    mean = X.mean(dim=1)
    R = X.shape[1]
    sq_mean = backend.dtw_squared(mean.repeat_interleave(R, dim=0), X.reshape(-1, 3012)).view(4, R).mean(dim=1)
    sq_bary = backend.dtw_squared(bary.repeat_interleave(R, dim=0), X.reshape(-1, 3012)).view(4, R).mean(dim=1)
    assert bool((sq_bary < sq_mean).all())


def test_python_mirror_signatures(backend):
    """tslearn's call as the reference makes it (models.py:176-178) and dtwa.performDBA."""
    import bayesian_ensembling_b200 as es

    rng = np.random.default_rng(8)
    X = _series(rng, 5, 60, "gmst")
    y = es.dtw_barycenter_averaging_subgradient(X, max_iter=50, tol=1e-3)
    assert y.shape == (60, 1)
    assert np.array_equal(y[:, 0], dba.dba_subgradient(X, max_iter=50, tol=1e-3)[0])
    assert np.array_equal(es.dtw_barycenter_averaging_subgradient(X[:, :, None], max_iter=3)[:, 0],
                          dba.dba_subgradient(X, max_iter=3)[0])
    c = es.performDBA(list(X), n_iterations=3)
    assert np.array_equal(c, dba.perform_dba(X, 3)[0])
    with pytest.raises(NotImplementedError):
        es.dtw_barycenter_averaging_subgradient(X, barycenter_size=10)
    with pytest.raises(NotImplementedError):
        es.dtw_barycenter_averaging_subgradient(X, weights=np.ones(5))
    with pytest.raises(ValueError):
        backend.dtw_barycenter_averaging_subgradient(_t(backend, np.zeros((1, 2, 5000))))


def test_gpdtw1d_default_uses_dba_mean(backend):
    """GPDTW1D().fit follows models.py:176-182: y_mean is the DBA mean, not the arithmetic mean."""
    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200.labelled import DataArray

    M, R, T = 3, 4, 40
    rng = np.random.default_rng(12)
    reals = np.stack([_series(rng, R, T, "shifted") for _ in range(M)])
    pms = [es.ProcessModel(DataArray(reals[m], ("realisation", "time")), f"model{m}") for m in range(M)]
    mc = es.ModelCollection(pms)
    mc.fit(es.GPDTW1D(hyperparameters=(0.5, 6.0)), progress_bar=False)
    mc_mean = es.ModelCollection([es.ProcessModel(DataArray(reals[m], ("realisation", "time")), f"model{m}")
                                  for m in range(M)])
    mc_mean.fit(es.GPDTW1D(hyperparameters=(0.5, 6.0), y_mean="mean"), progress_bar=False)
    for m in range(M):
        y_dba = dba.dba_subgradient(reals[m], max_iter=50, tol=1e-3)[0]
        X, y, s = rp.gpdtw1d_inputs(reals[m], y_dba)
        mu, cov = rp.gp_posterior_closed_form(X, y, s, 0.5, 6.0)
        assert rel_err(mc[m].distribution.mean.values, mu) <= 1e-8
        assert rel_err(mc[m].distribution._dist.covariance(), cov) <= 1e-8
        X, y, s = rp.gpdtw1d_inputs(reals[m])
        mu_m, _ = rp.gp_posterior_closed_form(X, y, s, 0.5, 6.0)
        assert rel_err(mc_mean[m].distribution.mean.values, mu_m) <= 1e-8
        assert rel_err(mu, mu_m) > 1e-4  # the two means really differ on warped inputs


def test_dba_nan_inputs_do_not_leave_the_table(backend):
    """NaN realisations poison their own problem only, and the path walk stays inside the table
    (a NaN table records "diag" on the borders): same outputs as the oracle, NaN pattern included."""
    rng = np.random.default_rng(21)
    X = np.stack([_series(rng, 3, 70, "gmst") for _ in range(3)])
    X[1, 0, 5] = np.nan
    bary, n_iter, _ = backend.dtw_barycenter_averaging_subgradient(_t(backend, X), max_iter=4, tol=1e-3, want_info=True)
    for b in range(3):
        want, n, _ = dba.dba_subgradient(X[b], max_iter=4, tol=1e-3)
        got = bary[b].cpu().numpy()
        assert (np.isnan(got) == np.isnan(want)).all()
        ok = ~np.isnan(want)
        assert np.array_equal(got[ok], want[ok]) and int(n_iter[b]) == n
    assert not np.isnan(bary[0].cpu().numpy()).any() and np.isnan(bary[1].cpu().numpy()).any()


def test_gpdtw3d_dtw_to_xarray_and_prep_data(backend):
    """GPDTW3D._dtw_to_xarray (models.py:238-268): the per-cell double loop of DBA calls as ONE batched device
    call, bit-identical to the oracle cell by cell; _prep_data (:270-322): X / Y against a literal restatement of
    the reference's coordinate arithmetic in to_dataframe row order."""
    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200.labelled import DataArray

    rng = np.random.default_rng(31)
    R, T, n_lat, n_lon = 3, 25, 4, 5
    data = np.stack([[_series(rng, R, T, "shifted") for _ in range(n_lon)] for _ in range(n_lat)])  # [lat,lon,R,T]
    data = np.ascontiguousarray(data.transpose(2, 3, 0, 1))                                          # [R,T,lat,lon]
    lat, lon = np.linspace(-60.0, 60.0, n_lat), np.linspace(0.0, 288.0, n_lon)
    pm = es.ProcessModel(DataArray(data, ("realisation", "time", "latitude", "longitude"),
                                   {"realisation": np.arange(R), "time": 1990 + np.arange(T), "latitude": lat,
                                    "longitude": lon}, name="tas"), "m")
    with pytest.warns(UserWarning, match="experimental"):
        g3 = es.GPDTW3D()
    mean_array, var_array = g3._dtw_to_xarray(pm)
    assert mean_array.dims == ("time", "latitude", "longitude") and mean_array.shape == (T, n_lat, n_lon)
    for i in range(n_lat):
        for j in range(n_lon):
            want = dba.dba_subgradient(data[:, :, i, j], max_iter=50, tol=1e-3)[0]
            assert np.array_equal(mean_array.values[:, i, j], want), (i, j)
            assert rel_err(var_array.values[:, i, j], np.var(data[:, :, i, j], axis=0)) < 1e-13
    X, Y = g3._prep_data(pm.model_data, mean_array, var_array)
    N = T * n_lat * n_lon
    assert X.shape == (N, 4 + R) and Y.shape == (N, 2)
    t_idx, la_idx, lo_idx = np.unravel_index(np.arange(N), (T, n_lat, n_lon))
    assert np.allclose(X[:, 0], np.cos(lat[la_idx] * np.pi / 180) * np.cos(lon[lo_idx] * np.pi / 180), atol=0, rtol=1e-15)
    assert np.allclose(X[:, 1], np.cos(lat[la_idx] * np.pi / 180) * np.sin(lon[lo_idx] * np.pi / 180), atol=0, rtol=1e-15)
    assert np.array_equal(X[:, 2], np.sin(lat[la_idx] * np.pi / 180))
    assert np.array_equal(X[:, 3], 2 * t_idx / (T - 1) - 1)
    assert np.array_equal(X[:, 4:], data[:, t_idx, la_idx, lo_idx].T)
    assert np.array_equal(Y[:, 0], mean_array.values[t_idx, la_idx, lo_idx])
    assert np.array_equal(Y[:, 1], var_array.values[t_idx, la_idx, lo_idx])
    bad = es.ProcessModel(DataArray(data[:, :, 0, 0], ("realisation", "time")), "b")
    with pytest.raises(NotImplementedError, match="4 dimensions"):
        g3.fit(bad)


@pytest.mark.parametrize("R,T,kind", [(12, 40, "shifted"), (9, 130, "gmst"), (10, 251, "ties")])
def test_dba_many_pairs_uses_the_thread_per_pair_walk(backend, R, T, kind):
    """From 8192 (problem, realisation) pairs on, the path walk runs one THREAD per pair: same barycentres,
    iteration counts and costs, bit for bit (a sample of problems against the oracle, all against a
    small-batch run of the warp-per-pair kernel)."""
    rng = np.random.default_rng(R * 1000 + T)
    B = 8192 // R + 8
    base = np.stack([_series(rng, R, T, kind) for _ in range(16)])
    X = np.tile(base, (B // 16 + 1, 1, 1))[:B] + 0.0
    X[:, 0, :] += 1e-3 * np.arange(B)[:, None]  # every problem differs
    Xd = _t(backend, X)
    bary, n_iter, cost = backend.dtw_barycenter_averaging_subgradient(Xd, max_iter=6, tol=1e-3, want_info=True)
    small, n_small, _ = backend.dtw_barycenter_averaging_subgradient(Xd[:64], max_iter=6, tol=1e-3, want_info=True)
    import torch

    assert torch.equal(bary[:64], small) and torch.equal(n_iter[:64], n_small)
    for b in (0, 17, B - 1):
        want, n, c = dba.dba_subgradient(X[b], max_iter=6, tol=1e-3)
        assert int(n_iter[b]) == n and float(cost[b]) == c
        assert np.array_equal(bary[b].cpu().numpy(), want)
