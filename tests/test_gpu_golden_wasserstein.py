"""The CUDA path (through the C ABI) against outputs of the REFERENCE'S OWN function bodies for rows a5-a8
and the Stein kernel of f-3 (tests/golden/wasserstein_reference.npz, made by executing the source text of
ensembles/wasserstein.py:10-100, ensemble_scheme.py:43-81 and weights.py:360-393 under jnp = numpy).
Needs a B200: ``-m gpu``."""
import os
import warnings

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wasserstein_reference.npz")


@pytest.fixture(scope="module")
def ref():
    return np.load(GOLDEN)


def _t(backend, a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=backend.device)


def test_a7_sqrtm_against_reference_outputs(backend, ref):
    """The device computes the principal root by scaled Denman-Beavers (stated deviation in METHOD); the
    reference's SVD form is the same matrix for SPD input: <= 1e-9 relative (north star: 1e-6)."""
    for i in range(int(ref["sqrtm_n"])):
        A, want = ref[f"sqrtm{i}_A"], ref[f"sqrtm{i}_root"]
        root, _, _, info = backend.sqrtm_psd(_t(backend, A[None]))
        if str(ref[f"sqrtm{i}_kind"]) == "singular":
            assert int(info[0]) != 0  # rank-deficient input is reported, never silently wrong
            continue
        assert int(info[0]) == 0
        assert rel_err(root[0].cpu().numpy(), want) < 1e-9, i


def test_a8_w2_against_reference_outputs(backend, ref):
    from bayesian_ensembling_b200 import wasserstein as ws

    class D:
        def __init__(self, m, S):
            self.m, self.S = m, S

        def mean(self):
            return self.m

        def covariance(self):
            return self.S

        def variance(self):
            return np.diag(self.S).copy()

    for i in range(int(ref["w2_n"])):
        m1, S1, m2, S2 = (ref[f"w2_{i}_{k}"] for k in ("mu1", "S1", "mu2", "S2"))
        a, b = D(m1, S1), D(m2, S2)
        full, diag = float(ref[f"w2_{i}_full"]), float(ref[f"w2_{i}_diag"])
        assert abs(ws.gaussian_w2_distance_distrax(a, b, full_cov=True) - full) <= 1e-9 * max(1.0, abs(full)), i
        assert abs(ws.gaussian_w2_distance_distrax(a, b, full_cov=False) - diag) <= 1e-9 * max(1.0, abs(diag)), i
        assert abs(ws.wasserstien_distance(S1, S2) - float(ref[f"w2_{i}_covonly"])) <= 1e-9 * max(1.0, abs(full)), i
        # W2(a, a): the reference's own value is rounding noise of the two SVD roots (|.| < 1e-12 * tr S)
        assert abs(ws.gaussian_w2_distance_distrax(a, a) - float(ref[f"w2_{i}_self"])) <= 1e-9 * np.trace(S1), i


def test_a5_gaussian_barycentre_against_reference_outputs(backend, ref):
    from bayesian_ensembling_b200 import wasserstein as ws

    for i in range(int(ref["bary_n"])):
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            mu, sigma = ws.gaussian_barycentre(ref[f"bary{i}_means"], ref[f"bary{i}_sd"], ref[f"bary{i}_w"],
                                               tolerance=float(ref[f"bary{i}_tol"]), init_var=float(ref[f"bary{i}_init"]))
        want_mu, want_sigma = float(ref[f"bary{i}_mu"]), float(ref[f"bary{i}_sigma"])
        assert (len(caught) > 0) == bool(ref[f"bary{i}_warned"]), str(ref[f"bary{i}_regime"])
        if np.isnan(want_mu):
            assert np.isnan(mu) and np.isnan(sigma)
            continue
        # the kernel takes variances (std**2 -> sqrt): one rounding away from the reference's std
        assert abs(mu - want_mu) <= 1e-14 * max(1.0, abs(want_mu)), (i, mu - want_mu)
        assert abs(sigma - want_sigma) <= 1e-12 * want_sigma, (i, sigma - want_sigma)


def test_a6_barycentre_scheme_against_reference_outputs(backend, ref):
    mu, sd, _ = backend.barycentre_1d(_t(backend, ref["scheme_mus"][None]), _t(backend, ref["scheme_var"][None]),
                                      _t(backend, ref["scheme_w"][None]))
    assert rel_err(mu[0].cpu().numpy(), ref["scheme_mu_out"]) <= 1e-14
    assert rel_err(sd[0].cpu().numpy() ** 2, ref["scheme_covariance_out"]) <= 1e-12


def test_f3_stein_discrepancy_against_reference_outputs(backend, ref):
    for i in range(int(ref["ksd_n"])):
        x = ref[f"ksd{i}_samples"]  # [N,1]
        loc = np.full((1, 1, 1), float(ref[f"ksd{i}_mean"]))
        scale = np.full((1, 1, 1), float(ref[f"ksd{i}_scale"]))
        _, ksd = backend.ksd_weights(_t(backend, loc), _t(backend, scale), _t(backend, x.reshape(1, -1, 1)), want_ksd=True)
        want = float(ref[f"ksd{i}_value"])
        assert abs(float(ksd.item()) - want) <= 1e-10 * want, (i, float(ksd.item()), want)
