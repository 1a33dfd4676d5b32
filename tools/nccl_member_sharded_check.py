"""Member-sharded fit -> weight -> barycentre under torchrun with NCCL (N = 2, 4, 8 GPUs):
every rank owns M / N members of every cell; one packed all-reduce joins them.
Checks the result against the single-GPU cell-resident path computed on rank 0, prints one JSON line.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_member_sharded_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesian_ensembling_b200 import grid, synthetic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    report = {}
    for T in (12, 24, int(os.environ.get("CHECK_T", "251"))):
        report[f"T={T}"] = run(T, rank, world)
    if rank == 0:
        assert report["T=12"]["time_mean=False"]["nan_fraction"] < 0.5, "the small case must exercise finite weights"
        print(json.dumps({"check": "member_sharded_nccl", "world": world, "members": 8, "cells": 2, **report}), flush=True)
    dist.destroy_process_group()


def run(T, rank, world):
    cfg = synthetic.Config("t", 9, 2, 8, 3, T, 10, False, "")
    reals, obs = synthetic.make_cells(cfg, seed=4242)
    var = np.linspace(0.4, 0.9, cfg.members)
    ls = np.linspace(5.0, 7.0, cfg.members)
    out = {}
    for time_mean in (False, True):
        lo, hi = grid.shard_range(cfg.members, rank, world)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        res = grid.fit_weight_barycentre_member_sharded(reals[:, lo:hi], obs, var[lo:hi], ls[lo:hi],
                                                        time_mean_weights=time_mean)
        dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        res = grid.fit_weight_barycentre_member_sharded(reals[:, lo:hi], obs, var[lo:hi], ls[lo:hi],
                                                        time_mean_weights=time_mean)
        ev1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # all ranks must hold the same barycentre
        b = torch.stack([res.bary_mu, res.bary_std])
        ref = b.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(torch.nan_to_num(b, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)))
        flags = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            full = grid.fit_weight_barycentre(reals, obs, var, ls, time_mean_weights=time_mean)
            def err(a, b_):
                a, b_ = a.cpu().numpy(), b_.cpu().numpy()
                assert (np.isnan(a) == np.isnan(b_)).all()
                ok = ~np.isnan(b_)
                return float(np.abs(a[ok] - b_[ok]).max() / max(np.abs(b_[ok]).max(), 1e-300)) if ok.any() else 0.0
            out[f"time_mean={time_mean}"] = {
                "ms": float(ms.item()), "ranks_bitwise_equal": bool(flags.item()),
                "weights_err": err(res.weights, full.weights[:, lo:hi]),
                "bary_mu_err": err(res.bary_mu, full.bary_mu), "bary_std_err": err(res.bary_std, full.bary_std),
                "nan_fraction": float(torch.isnan(full.weights).double().mean().item())}
            for k in ("weights_err", "bary_mu_err", "bary_std_err"):
                assert out[f"time_mean={time_mean}"][k] < 1e-10, out
            assert out[f"time_mean={time_mean}"]["ranks_bitwise_equal"]
    return out


if __name__ == "__main__":
    main()
