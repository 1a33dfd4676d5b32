"""BASELINE config 3 at FULL size: synthetic HadCRUT5-shaped 5 x 5 degree grid (36 x 72 = 2592 cells) x 165 years
monthly (T = 1980), 24 models x 5 realisations, per-cell GP posterior + LogLikelihoodWeight + Barycentre, cells
sharded across the ranks of one box with no data-path collective.  One JSON line from rank 0.

    python tools/run_cfg3.py [--workload cfg3|cfg4] [--cells 2592] [--wave 24] [--posterior dense|factored] [--y-mean mean|dba]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_cfg3.py

Inputs are generated on the host per wave (seeded per cell, so any sharding reproduces the same numbers) and
copied to the device outside the timed region; the timed region is the device pipeline of every wave
(CUDA events, summed per rank, max over ranks).  Size-independent checks: finite weights sum to 1 over
models, NaN pattern of the barycentre equals that of the weights, every factorisation reports info == 0,
and a sample of cells is re-run alone and must reproduce the wave's numbers bit for bit.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_ensembling_b200 import grid, synthetic  # noqa: E402
from bayesian_ensembling_b200.backend import Backend  # noqa: E402


def make_cells_device(cfg, n_cells, cell_offset, device):
    """The construction of synthetic.make_cells (trend + cell offset + seasonal term + AR(1) noise per realisation,
    SURVEY 8d) generated ON THE DEVICE from a per-wave seed: the host generator takes minutes for the 52 GB of the
    full cfg4 grid.  Different random numbers than the host generator (Philox instead of PCG64), same distribution;
    used for the full-size throughput / property runs only -- parity tests keep the host generator."""
    M, R, T, Ro = cfg.members, cfg.realisations, cfg.steps, cfg.obs_realisations
    g = torch.Generator(device=device).manual_seed(20240 + cfg.index + 7919 * cell_offset)
    gm = torch.Generator(device=device).manual_seed(20240 + cfg.index)  # member coefficients: shared by all cells
    f64 = dict(dtype=torch.float64, device=device)
    a = 0.5 + 3.5 * torch.rand(M + 1, generator=gm, **f64)
    b = 2.0 * torch.rand(M + 1, generator=gm, **f64)
    tn = torch.linspace(0.0, 1.0, T, **f64)
    season = 0.3 * torch.sin(2.0 * np.pi * (torch.arange(T, **f64) % 12) / 12.0) if cfg.monthly else torch.zeros(T, **f64)
    off = 0.5 * torch.randn(n_cells, generator=g, **f64)

    def ar1(shape, phi=0.6, sigma=0.12):
        eps = torch.randn(shape, generator=g, **f64) * (sigma * np.sqrt(1.0 - phi * phi))
        out = torch.empty(shape, **f64)
        out[..., 0] = torch.randn(shape[:-1], generator=g, **f64) * sigma
        for t in range(1, shape[-1]):
            out[..., t] = phi * out[..., t - 1] + eps[..., t]
        return out

    trend = a[:M, None] * tn[None, :] + b[:M, None] * tn[None, :] ** 2
    reals = (off[:, None, None] + trend[None] + season[None, None, :])[:, :, None, :] + ar1((n_cells, M, R, T))
    otrend = a[M] * tn + b[M] * tn**2
    obs = (off[:, None] + otrend[None] + season[None])[:, None, :] + ar1((n_cells, Ro, T))
    return reals.contiguous(), obs.contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4"],
                    help="cfg4: 1 x 1 degree grid (64 800 cells) x 251 years annual, 40 models x 10 realisations")
    ap.add_argument("--cells", type=int, default=0, help="0: the whole grid (2592 / 64 800)")
    ap.add_argument("--wave", type=int, default=0, help="cells per device wave (0: 24 / 128)")
    ap.add_argument("--posterior", default="dense", choices=["dense", "factored"])
    ap.add_argument("--y-mean", default="mean", choices=["mean", "dba"])
    ap.add_argument("--device-inputs", action="store_true", help="generate the synthetic inputs on the device (SURVEY 8d)")
    args = ap.parse_args()
    real = os.dup(1)
    os.dup2(2, 1)
    out_stream = os.fdopen(real, "w", buffering=1)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = Backend.get()
    cfg = synthetic.CONFIGS[args.workload]
    args.cells = args.cells or (2592 if args.workload == "cfg3" else 64800)
    args.wave = args.wave or (24 if args.workload == "cfg3" else 256)
    lo, hi = grid.shard_range(args.cells, rank, world)
    var, ls = synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE
    dev_ms, n_nan_cols, n_cols, worst_sum, bad_info = 0.0, 0, 0, 0.0, 0
    recheck = []
    t_wall = time.perf_counter()
    first = True
    for c0 in range(lo, hi, args.wave):
        n = min(args.wave, hi - c0)
        if args.device_inputs:
            r, o = make_cells_device(cfg, n, c0, be.device)
        else:
            reals, obs = synthetic.make_cells(cfg, n_cells=n, cell_offset=c0)
            r, o = torch.as_tensor(reals, device=be.device), torch.as_tensor(obs, device=be.device)
        if first:  # warm-up (workspace allocation, kernel attributes) outside the timed region
            grid.fit_weight_barycentre(r, o, var, ls, cells_per_wave=n, posterior=args.posterior, y_mean=args.y_mean)
            first = False
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = grid.fit_weight_barycentre(r, o, var, ls, cells_per_wave=n, posterior=args.posterior, y_mean=args.y_mean)
        e1.record()
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e1)
        w = res.weights
        nan_col = torch.isnan(w).any(dim=1)
        n_nan_cols += int(nan_col.sum())
        n_cols += nan_col.numel()
        if (~nan_col).any():
            worst_sum = max(worst_sum, float((w.sum(dim=1)[~nan_col] - 1.0).abs().max()))
        assert bool((torch.isnan(res.bary_mu) == nan_col).all()), "barycentre NaN pattern differs from the weights'"
        bad_info += int(res.info_fit.abs().sum()) + int(res.info_dist.abs().sum())
        if len(recheck) < 2:  # a cell of this wave alone: bit-identical
            k = n // 2
            alone = grid.fit_weight_barycentre(r[k:k + 1], o[k:k + 1], var, ls, cells_per_wave=1,
                                               posterior=args.posterior, y_mean=args.y_mean)
            recheck.append(bool(torch.equal(alone.mu[0], res.mu[k]) and torch.equal(alone.var_diag[0], res.var_diag[k])
                                and torch.equal(torch.nan_to_num(alone.weights[0]), torch.nan_to_num(res.weights[k]))))
    wall = time.perf_counter() - t_wall
    stats = torch.tensor([dev_ms, float(n_nan_cols), float(n_cols), worst_sum, float(bad_info), float(all(recheck)), wall],
                         dtype=torch.float64, device=be.device)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        mn = stats.clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    else:
        mx = sm = mn = stats
    if rank == 0:
        line = {
            "config": "%s%s: %d cells x %d members x %d realisations x %d time steps" % (
                args.workload, " full size" if args.cells in (2592, 64800) else " (part of the grid)", args.cells,
                cfg.members, cfg.realisations, cfg.steps),
            "n_gpus": world, "cells": args.cells, "cells_per_wave": args.wave, "posterior": args.posterior,
            "y_mean": args.y_mean, "inputs": "generated on the device" if args.device_inputs else "generated on the host",
            "device_seconds_max_over_ranks": float(mx[0]) / 1e3,
            "cells_per_sec": args.cells / (float(mx[0]) / 1e3),
            "member_posteriors_per_sec": args.cells * cfg.members / (float(mx[0]) / 1e3),
            "wall_seconds_incl_host_generation_max_over_ranks": float(mx[6]),
            "weights_nan_fraction_of_points": float(sm[1]) / float(sm[2]),
            "max_abs_weight_sum_minus_one": float(mx[3]),
            "nonzero_info": int(sm[4]),
            "single_cell_rerun_bit_identical": bool(float(mn[5]) == 1.0),
            "sharding": "cells across ranks, no collective on the data path" if world > 1 else "single GPU",
        }
        print(json.dumps(line), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
