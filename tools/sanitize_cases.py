"""A small tour of the hot path for compute-sanitizer (tools/gpu_sanitize.sh): every kernel family once, at sizes
that cover the one-block (T=17), two-block (T=129, 251) and ragged-edge regimes of the padded layout.  Results are
checked against the oracle so that a sanitizer run is also a correctness run.  Usage:
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_cases.py [--quick]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bayesian_ensembling_b200 import grid, synthetic  # noqa: E402
from bayesian_ensembling_b200.backend import Backend  # noqa: E402
from oracle import reference_path as rp  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    ok = ~np.isnan(b)
    assert (np.isnan(a) == np.isnan(b)).all()
    return float(np.abs(a[ok] - b[ok]).max() / max(np.abs(b[ok]).max(), 1e-300)) if ok.any() else 0.0


def repeat_check(n):
    """Run-to-run determinism of the posterior pipeline: a shared-memory race or an ordering bug between launches
    shows up as results that differ between identical calls (every kernel here has a fixed summation order)."""
    for T, C in ((17, 3), (129, 2), (251, 40)):
        cfg = synthetic.Config("s", 7, C, 4, 3, T, 3, False, "sanitize")
        reals, obs = synthetic.make_cells(cfg)
        first = None
        for _ in range(n):
            res = grid.fit_weight_barycentre(reals, obs, 0.5, 6.0, keep_posteriors=True)
            got = [getattr(res, k).cpu().numpy() for k in ("mu", "cov", "scale_tri", "weights", "bary_mu", "bary_std")]
            if first is None:
                first = got
            else:
                for a, b in zip(first, got):
                    assert np.array_equal(a, b, equal_nan=True), f"T={T}: results differ between identical calls"
        print(f"T={T} x {C} cells: {n} identical calls, bit-identical results", flush=True)


def main():
    quick = "--quick" in sys.argv
    be = Backend.get()
    if "--repeat" in sys.argv:
        repeat_check(int(sys.argv[sys.argv.index("--repeat") + 1]))
        print("repeat check ok")
        return
    worst = 0.0
    for T in ((17, 129) if quick else (17, 129, 251)):
        cfg = synthetic.Config("s", 7, 2, 3, 3, T, 3, False, "sanitize")
        reals, obs = synthetic.make_cells(cfg)
        res = grid.fit_weight_barycentre(reals, obs, 0.5, 6.0, keep_posteriors=True)
        be.sync()
        o = rp.cell_pipeline_L1(reals[0], obs[0], 0.5, 6.0)
        for name, got, want, tol in (("mu", res.mu[0], o["mu"], 1e-8), ("cov", res.cov[0], o["cov"], 1e-8),
                                     ("scale_tri", res.scale_tri[0], o["scale_tri"], 1e-8),
                                     ("weights", res.weights[0], o["weights"], 1e-6),
                                     ("bary_mu", res.bary_mu[0], o["bary_mu"], 1e-6)):
            e = rel(got.cpu().numpy(), want)
            assert e <= tol, (T, name, e)
            worst = max(worst, e)
        rf = grid.fit_weight_barycentre(reals, obs, 0.5, 6.0, posterior="factored")
        assert rel(rf.mu.cpu().numpy(), res.mu.cpu().numpy()) < 1e-10
        print(f"T={T}: pipeline ok", flush=True)
    # many problems: the two-CTA diagonal-block form and the small-T member kernel regimes
    cfg = synthetic.Config("s", 7, 80, 4, 3, 40, 3, False, "sanitize")
    reals, obs = synthetic.make_cells(cfg)
    res = grid.fit_weight_barycentre(reals, obs, 0.5, 6.0)
    o = rp.cell_pipeline_L1(reals[5], obs[5], 0.5, 6.0)
    assert rel(res.mu[5].cpu().numpy(), o["mu"]) < 1e-8 and rel(res.weights[5].cpu().numpy(), o["weights"]) < 1e-6
    print("320 problems: ok", flush=True)
    # the training loop (CUDA graph), two iterations
    T = 40
    cfg = synthetic.Config("s", 7, 1, 2, 3, T, 3, False, "sanitize")
    reals, obs = synthetic.make_cells(cfg)
    X, ym, yv = be.gpdtw1d_inputs(torch.as_tensor(reals[0], device=be.device))
    post, var, ls = be.vgp_fit(X, ym, yv, 2)
    mu_o, cov_o, st = rp.gpdtw1d_fit(reals[0, 0], n_optim_nits=2, return_state=True)
    assert rel(post.mu[0].cpu().numpy(), mu_o) < 1e-7 and rel(post.cov[0].cpu().numpy(), cov_o) < 1e-7
    print("vgp_fit: ok", flush=True)
    # sqrtm / W2 / full-covariance barycentre
    covs = np.stack([rp.gp_posterior_closed_form(*rp.gpdtw1d_inputs(reals[0, m]), 0.5, 6.0)[1] for m in range(2)])
    mus = np.stack([rp.gp_posterior_closed_form(*rp.gpdtw1d_inputs(reals[0, m]), 0.5, 6.0)[0] for m in range(2)])
    root, _, _, info = be.sqrtm_psd(torch.as_tensor(covs, device=be.device))
    assert int(info.abs().sum()) == 0 and rel(root[0].cpu().numpy(), rp.sqrtm_svd(covs[0])) < 1e-8
    w2, _ = be.w2_distance(torch.as_tensor(mus[:1], device=be.device), torch.as_tensor(covs[:1], device=be.device),
                           torch.as_tensor(mus[1:], device=be.device), torch.as_tensor(covs[1:], device=be.device))
    assert abs(float(w2[0]) - rp.gaussian_w2_distance(mus[0], covs[0], mus[1], covs[1])) < 1e-8
    w = np.array([[0.4, 0.6]])
    mu_b, S_b, iters, _ = be.barycentre_fullcov(torch.as_tensor(mus[None], device=be.device),
                                                 torch.as_tensor(covs[None], device=be.device), torch.as_tensor(w, device=be.device))
    mo, So, _ = rp.fullcov_barycentre(mus, covs, w[0])
    assert rel(S_b[0].cpu().numpy(), So) < 1e-6
    print("sqrtm / w2 / fullcov barycentre: ok", flush=True)
    # next-row weights and the DBA step
    var = np.stack([np.diag(c) for c in covs])
    for fn, ofn in ((be.crps_weights, rp.crps_weights), (be.ksd_weights, rp.ksd_weights)):
        got = fn(torch.as_tensor(mus[None], device=be.device), torch.as_tensor(var[None], device=be.device),
                 torch.as_tensor(obs[:1], device=be.device))
        assert rel(got[0].cpu().numpy(), ofn(mus, var, obs[0])[0]) < 1e-9
    from oracle import dba as oracle_dba

    bary = be.dtw_barycenter_averaging_subgradient(torch.as_tensor(reals[0], device=be.device), max_iter=5, tol=1e-3)
    want, _, _ = oracle_dba.dba_subgradient(reals[0, 0], max_iter=5, tol=1e-3)
    assert np.array_equal(bary[0].cpu().numpy(), want)
    print("crps / ksd / dba: ok", flush=True)
    print(f"sanitize tour ok, worst relative error {worst:.2e}")


if __name__ == "__main__":
    main()
