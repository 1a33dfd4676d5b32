"""N > 1 host logic on CPU: two gloo ranks run the member-sharded pipeline (each owns half of the
members of every cell) and must reproduce the single-process oracle.  The arithmetic operators are
an oracle-backed stand-in (tests/_oracle_ops.py) -- what is under test is the partitioning, the
packed all-reduce and its ordering, which is identical for NCCL."""
import os
import warnings

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesian_ensembling_b200 import grid, synthetic
from oracle import reference_path as rp


def _worker(rank, world, port, time_mean, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _oracle_ops import OracleOps

    cfg = synthetic.Config("t", 9, 2, 6, 3, 8, 4, False, "")
    reals, obs = synthetic.make_cells(cfg, seed=99)
    lo, hi = grid.shard_range(cfg.members, rank, world)
    var = np.linspace(0.4, 0.9, cfg.members)
    ls = np.linspace(5.0, 7.0, cfg.members)
    res = grid.fit_weight_barycentre_member_sharded(reals[:, lo:hi], obs, var[lo:hi], ls[lo:hi], ops=OracleOps(),
                                                    time_mean_weights=time_mean)
    q.put((rank, lo, hi, res.weights.numpy(), res.bary_mu.numpy(), res.bary_std.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("time_mean", [False, True])
def test_member_sharded_two_ranks_equal_single_process(time_mean):
    world, port = 2, 29500 + (os.getpid() % 500) + (7 if time_mean else 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, time_mean, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = synthetic.Config("t", 9, 2, 6, 3, 8, 4, False, "")
    reals, obs = synthetic.make_cells(cfg, seed=99)
    var = np.linspace(0.4, 0.9, cfg.members)
    ls = np.linspace(5.0, 7.0, cfg.members)
    for c in range(cfg.cells):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o = rp.cell_pipeline_L1(reals[c], obs[c], var, ls, time_mean_weights=time_mean)
        for rank, lo, hi, w, bmu, bsd in outs:
            ok = ~np.isnan(o["weights"][lo:hi])
            assert ok.any()
            assert (np.isnan(w[c]) == ~ok).all()
            assert np.abs(w[c][ok] - o["weights"][lo:hi][ok]).max() < 1e-10
            okb = ~np.isnan(o["bary_mu"])
            assert (np.isnan(bmu[c]) == ~okb).all()
            assert np.abs(bmu[c][okb] - o["bary_mu"][okb]).max() < 1e-10
            assert np.abs(bsd[c][okb] - o["bary_std"][okb]).max() < 1e-10


def test_shard_range_partitions_exactly():
    for n in (1, 7, 24, 2592, 64800):
        for world in (1, 2, 4, 8):
            spans = [grid.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
