"""The algebraic identities the per-point weight kernels rest on, checked on the CPU against the oracle's direct
restatement of the reference arithmetic (the device runs of the same identities are in tests/test_gpu_parity.py):

* KSD (weights_next_kernels.cuh::k_ksd_weights): the IMQ pair kernel's powers do not depend on the model, so
  sum_ab k0 is a quadratic in (loc - mean(x)) over five per-point moments;
* Normal-branch log-likelihood (be_kernels.cuh::k_loglik_weights_normal): the mean log-density from the
  observations' mean and centred second moment, with the residual of the rounded pivot carried;
* pairwise temporal similarity (k_similarity_pointwise): M square roots per clean point instead of M^2.
"""
import numpy as np

from oracle import reference_path as rp


def _ksd_factored(x, loc, scale):
    ro = x.size
    xbar = x.sum() / ro
    u = x - xbar
    sa, sua, suua, sub, sc = float(ro), u.sum(), float((u * u).sum()), 0.0, float(ro)
    for a in range(ro):
        for b in range(a + 1, ro):
            d = x[a] - x[b]
            d2 = d * d
            q = 1.0 + d2
            p05 = 1.0 / np.sqrt(q)
            p15 = p05 / q
            p25 = p15 / q
            sa += 2.0 * p05
            sua += (u[a] + u[b]) * p05
            suua += 2.0 * (u[a] * u[b]) * p05
            sub += d2 * p15
            sc += 2.0 * (p15 - 3.0 * p25 * d2)
    i2 = 1.0 / (scale * scale)
    dl = loc - xbar
    return np.sqrt((((dl * dl) * sa - 2.0 * dl * sua + suua) * i2 - 2.0 * sub) * i2 + sc) / ro


def test_ksd_factored_form_equals_direct_double_sum():
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(600):
        ro = int(rng.integers(1, 12))
        spread = 10 ** rng.uniform(-3, 0.5)
        x = 0.8 + spread * rng.normal(size=ro)
        loc = 0.8 + rng.normal() * 10 ** rng.uniform(-4, 0.5)
        scale = rng.uniform(0.01, 0.6)
        want = rp.ksd_imq(x, -(x - loc) / (scale * scale))
        worst = max(worst, abs(_ksd_factored(x, loc, scale) / want - 1.0))
    assert worst < 5e-14, worst


def test_normal_branch_moment_form_equals_mean_of_log_densities():
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(600):
        ro = int(rng.integers(1, 12))
        centre = rng.normal() * 10 ** rng.uniform(-1, 2.5)
        spread = 10 ** rng.uniform(-4, 0.5)
        o = centre + spread * rng.normal(size=ro)
        loc = centre + spread * rng.normal() * 10 ** rng.uniform(-3, 1)
        scale = spread * 10 ** rng.uniform(-0.5, 1)
        want = np.mean([rp.normal_log_prob(loc, scale, np.asarray([v]))[0] for v in o])
        mean_o = o.sum() / ro
        d = o - mean_o
        var_o, mean_d = (d * d).sum() / ro, d.sum() / ro
        dl = mean_o - loc
        got = (-0.5 * ((var_o + dl * (2.0 * mean_d + dl)) / (scale * scale)) - 0.5 * np.log(2.0 * np.pi)) - np.log(scale)
        worst = max(worst, abs(got - want) / max(1.0, abs(want)))
    assert worst < 1e-13, worst


def test_similarity_row_sum_with_m_square_roots():
    rng = np.random.default_rng(2)
    m_models, n = 24, 50
    mean, var = rng.normal(size=(m_models, n)), rng.uniform(0.05, 0.5, size=(m_models, n))
    want_w, want_d = rp.model_similarity_weights_temporal(mean, var)
    v = var * var  # the reference's dx.Normal(mean, variance).variance()
    r = np.sqrt(v)
    sabs = np.abs(mean[:, None] - mean[None]).sum(axis=1)
    row = (sabs + ((m_models * v + v.sum(axis=0)[None]) - 2.0 * r * r.sum(axis=0)[None])) / m_models
    got = row / row.sum(axis=0)
    assert np.abs(got - want_w).max() < 1e-13
    assert np.abs(row - np.asarray(want_d).mean(axis=1)).max() < 1e-13
