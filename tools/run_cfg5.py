"""BASELINE config 5 at full size: full-covariance Gaussian W2 barycentre over the 24 member posteriors
(3012-dim) of one cfg2 cell.  Prints one JSON line: timings and the size-independent checks
(the reference has no code for this config).

    python tools/run_cfg5.py [--members 24] [--steps 3012]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_ensembling_b200 import synthetic  # noqa: E402
from bayesian_ensembling_b200.backend import Backend  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=24)
    ap.add_argument("--steps", type=int, default=3012)
    ap.add_argument("--scale", type=float, default=1.0, help="multiply the covariances (>~500: the fixed point iterates)")
    args = ap.parse_args()
    be = Backend.get()
    cfg = synthetic.Config("cfg5", 5, 1, args.members, 5, args.steps, 10, True, "cfg5")
    reals, _ = synthetic.make_cells(cfg)
    M, T = args.members, args.steps
    X, ym, yv = be.gpdtw1d_inputs(torch.as_tensor(reals[0], device=be.device))
    var = torch.full((M,), synthetic.L1_VARIANCE, dtype=torch.float64, device=be.device)
    ls = torch.full((M,), synthetic.L1_LENGTHSCALE, dtype=torch.float64, device=be.device)
    post = be.gp_posterior(X, ym, yv, var, ls, want_scale_tri=False)
    cov = post.cov * args.scale
    w = torch.full((1, M), 1.0 / M, dtype=torch.float64, device=be.device)
    torch.cuda.synchronize()
    out = {}
    for rep in range(2):
        l0 = be.launch_count
        t0 = time.perf_counter()
        mu, S, iters, info = be.barycentre_fullcov(post.mu[None], cov[None], w)
        torch.cuda.synchronize()
        out = {"seconds": time.perf_counter() - t0, "launches": be.launch_count - l0, "outer_iterations": iters[0]}
    assert int(info.abs().sum()) == 0
    S = S[0]
    checks = {"symmetric": bool(torch.equal(S, S.T)),
              "mean_err": float((mu[0] - post.mu.mean(0)).abs().max())}
    # positive definite: the device Cholesky succeeds
    _, pd_info = be.potrf(S[None])
    checks["cholesky_info"] = int(pd_info.item())
    if iters[0] == 0:
        # signed stop rule exits at iteration 0 (tr S < init_var): S = sum_m w_m sqrtm(Sigma_m) exactly
        t0 = time.perf_counter()
        roots, _, sq_iters, _ = be.sqrtm_psd(cov)
        torch.cuda.synchronize()
        out["sqrtm_batch_seconds"] = time.perf_counter() - t0
        out["sqrtm_iterations"] = sq_iters
        want = (roots * w[0][:, None, None]).sum(0)
        checks["iteration0_identity_err"] = float((S - want).abs().max() / want.abs().max())
        r0 = roots[0]
        checks["sqrtm_residual"] = float(((r0 @ r0 - cov[0]).abs().max() / cov[0].abs().max()))
        flops = sq_iters * 2.0 * M * float(T) ** 3
        out["sqrtm_tflops"] = flops / out["sqrtm_batch_seconds"] / 1e12
    print(json.dumps({"config": f"cfg5: {M} members x {T}-dim posteriors, uniform weights, scale {args.scale}", **out, **checks}))


if __name__ == "__main__":
    main()
