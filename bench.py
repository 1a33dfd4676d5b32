#!/usr/bin/env python
"""Benchmark of the fit -> weight -> barycentre hot path (BASELINE.json metric: grid-cells/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the CPU arm (oracle port, host cores)

Workload (N=1 and every N, weak scaling): BASELINE.json configs[1] -- single-location monthly
series, 24 CMIP6-shaped members x 5 realisations x 3012 months, full-covariance GP posterior per
member, 10 observation realisations -- batched ``--cells-per-step`` cells per GPU per step.  A
"step" is one pass of gpdtw1d_inputs -> gp_posterior (gram, Cholesky, inverse, covariance,
distribution Cholesky) -> LogLikelihoodWeight -> Barycentre over that batch.  Cells are sharded
across ranks (they never interact), so there is no collective on the data path.

One JSON line on stdout (rank 0).  ``value`` = cells/s with inputs resident in HBM; ``e2e`` = the
same through the public batched API with pinned HOST buffers (H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid_cells_per_sec_fit_weight_barycentre"
UNIT = "cells/s"
FP64_PEAK_TFLOPS = 37.15  # DMMA issue-rate peak measured on this pool: profiles/r01_ubench_fp64.txt


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def workload_config(cfg, cells_per_step, n_gpus):
    return {
        "workload": f"{cfg.name}: {cfg.description}",
        "cells_per_step_per_gpu": cells_per_step,
        "members": cfg.members, "realisations": cfg.realisations, "time_steps": cfg.steps,
        "obs_realisations": cfg.obs_realisations,
        "fit_level": "L1 fixed kernel hyper-parameters (variance 0.5, lengthscale 6.0): posterior + "
                     "distribution Cholesky + LogLikelihoodWeight + Barycentre",
        "sharding": "cells across ranks, no collective" if n_gpus > 1 else "single GPU",
        "l2": f"no flush needed: work matrices per step = {cells_per_step * cfg.members * 2 * (cfg.steps + 2) ** 2 * 8 / 1e6:.0f} MB "
              f"(2 padded T x T fp64 matrices per member) vs 126 MB L2",
    }


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 200 ms while the timed region runs."""

    REASONS = {
        0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
        0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference arithmetic, all host threads
# ------------------------------------------------------------------------------------------------
def _blas_threads():
    try:
        from threadpoolctl import threadpool_info

        infos = [i for i in threadpool_info() if i.get("user_api") == "blas"]
        return max([i.get("num_threads", 1) for i in infos] or [1]), (infos[0].get("internal_api") if infos else "?")
    except Exception:
        return os.cpu_count() or 1, "?"


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all host cores."""
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count() or 1, user_api="blas")
    except Exception:
        pass


def cpu_sample_seconds(cfg, members, repeats=1):
    """Oracle fit -> weight -> barycentre on ``members`` members of cell 0 (bounded sample)."""
    from bayesian_ensembling_b200 import synthetic
    from oracle import reference_path as rp

    _use_all_host_threads()

    reals, obs = synthetic.make_cells(cfg, n_cells=1)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        rp.cell_pipeline_L1(reals[0, :members], obs[0], synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE)
        best = min(best, time.perf_counter() - t0)
    return best


def _one_cell_single_thread(cfg_name, cell):
    """Worker of cpu_small_t_cell_seconds: one whole cell through the oracle with one BLAS thread."""
    from threadpoolctl import threadpool_limits

    from bayesian_ensembling_b200 import synthetic
    from oracle import reference_path as rp

    cfg = synthetic.CONFIGS[cfg_name]
    reals, obs = synthetic.make_cells(cfg, n_cells=1, cell_offset=cell)
    with threadpool_limits(limits=1, user_api="blas"):
        t0 = time.perf_counter()
        rp.cell_pipeline_L1(reals[0], obs[0], synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE)
        return time.perf_counter() - t0


def cpu_small_t_cell_seconds(cfg, rounds=2):
    """Small T (cfg1 / cfg4 shapes): threaded BLAS does not pay on 251 x 251 matrices, so the host cores are used
    the way SURVEY 8d asks -- one process per core, one BLAS thread each, every process a whole cell (all its
    members).  Returns (seconds per cell at full occupancy of the host, processes, description)."""
    from joblib import Parallel, delayed

    procs = os.cpu_count() or 1
    with Parallel(n_jobs=procs, backend="loky") as par:
        par(delayed(_one_cell_single_thread)(cfg.name, c) for c in range(procs))  # worker start-up and imports
        t0 = time.perf_counter()
        par(delayed(_one_cell_single_thread)(cfg.name, c) for c in range(procs * rounds))
        wall = time.perf_counter() - t0
    return wall / (procs * rounds), procs, (
        f"{procs * rounds} whole {cfg.name} cells (T={cfg.steps}, {cfg.members} members, Ro={cfg.obs_realisations}) on "
        f"{procs} processes with one BLAS thread each: {wall:.1f} s")


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _use_all_host_threads()
    threads, blas = _blas_threads()
    if cfg.steps < 1000 and not args.cpu_members:
        cell_s, threads, sample = cpu_small_t_cell_seconds(cfg, rounds=max(1, args.steps))
        dt, blas = cell_s, blas + ", 1 thread per process"
    else:
        m = min(cfg.members, args.cpu_members or 6)
        for _ in range(args.warmup):
            cpu_sample_seconds(cfg, m)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_sample_seconds(cfg, m)
        dt = (time.perf_counter() - t0) / args.steps
        cell_s = dt * cfg.members / m
        sample = (f"{m} of {cfg.members} members of one {cfg.name} cell per step (T={cfg.steps}, "
                  f"Ro={cfg.obs_realisations}), scaled x{cfg.members / m:.0f} to a cell")
    value = 1.0 / cell_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, args.cells_per_step, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "blas": blas, "host_cpus": os.cpu_count(),
                         "note": "NumPy/SciPy oracle restating the reference arithmetic; the reference itself "
                                 "(GPflow/TF/JAX) cannot be installed here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# memory-bound stages measured alone
# ------------------------------------------------------------------------------------------------
def measure_hbm_stages(be, cfg, n_points, hbm_peak, reps=5):
    """LogLikelihoodWeight, weight time-mean and Barycentre kernels on C x T = n_points points with M
    members (cfg3/cfg4-shaped: many cells), finite weights, buffers >> 126 MB L2.  Algorithmic bytes
    per point are DESIGN.md section 4's figures; time = best of ``reps`` CUDA-event timings."""
    import torch

    M, Ro, T = cfg.members, cfg.obs_realisations, 1980
    C = max(1, n_points // T)
    N = C * T
    g = torch.Generator(device=be.device).manual_seed(1)
    rnd = lambda *s: torch.rand(*s, dtype=torch.float64, device=be.device, generator=g)  # noqa: E731
    # constant-vector log-prob statistics (|a|^2, a.b, |b|^2, sum log diag L) giving ll in about [-6, 0]
    a2 = 0.5 + rnd(C * M)
    stats = torch.stack([a2, a2 * (0.9 + 0.2 * rnd(C * M)), a2 * (1.0 + 0.2 * rnd(C * M)),
                         -0.5 * T * 1.8378770664093453 + rnd(C * M)], dim=1).contiguous()
    obs = 0.8 + 0.4 * rnd(C, Ro, T)
    means, variances = rnd(C, M, T), 0.01 + 0.05 * rnd(C, M, T)

    def timed(fn):
        best = float("inf")
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, out

    res = {}
    ms, w = timed(lambda: be.loglik_weights_mvn(stats, obs, M))
    assert bool(torch.isfinite(w).all())
    res["k_loglik_weights"] = (ms, N * 8.0 * (Ro + M) + C * M * 32.0)
    ms, wb = timed(lambda: be.weights_time_mean(w))
    res["k_weights_time_mean"] = (ms, N * M * 16.0)
    ms, out = timed(lambda: be.barycentre_1d(means, variances, w))
    assert int(out[2].max()) == 0  # degC-anomaly scale: the signed stop rule exits at iteration 0
    res["k_barycentre"] = (ms, N * (24.0 * M + 16.0 + 4.0))
    # SURVEY 8f rows 2-3 (CRPSWeight is the weight the published experiment uses): same [C, M, N] layout.  Their
    # HBM bytes are as small as the log-likelihood kernel's, but the reference arithmetic asks for M * Ro erfc + exp
    # (CRPS), M * Ro^2 kernel evaluations (KSD; the kernel's factored form does Ro (Ro - 1) / 2 per point, the
    # figure below counts the reference's M * Ro^2) and M^2 square roots (similarity) per point: FP64-pipe work,
    # so the fraction of the HBM peak is reported for what it is, next to the evaluations per second.
    sd = variances.sqrt()
    ms, wc = timed(lambda: be.crps_weights(means, sd, obs))
    assert bool(torch.isfinite(wc).all())
    res["k_crps_weights"] = (ms, N * 8.0 * (Ro + 3.0 * M))
    ms, _ = timed(lambda: be.ksd_weights(means, sd, obs))
    res["k_ksd_weights"] = (ms, N * 8.0 * (Ro + 3.0 * M))
    loc_n, scale_n = (obs.mean(dim=1, keepdim=True) + 0.1 * (means - 0.5)).contiguous(), 0.3 + sd
    ms, wn = timed(lambda: be.loglik_weights_normal(loc_n, scale_n, obs))
    assert bool(torch.isfinite(wn).all())
    res["k_loglik_weights_normal"] = (ms, N * 8.0 * (Ro + 3.0 * M))
    ms, _ = timed(lambda: be.similarity_weights_pointwise(means, variances))
    res["k_similarity_pointwise"] = (ms, N * 8.0 * 3.0 * M)
    evals = {"k_crps_weights": float(N) * M * Ro, "k_ksd_weights": float(N) * M * Ro * Ro,
             "k_similarity_pointwise": float(N) * M * M}
    out = {"points": N, "cells": C, "time_steps": T, "members": M, "peak_gbs": hbm_peak,
           "kernels": {k: {"ms": ms, "algorithmic_bytes": b, "gbs": b / ms / 1e6, "frac_of_hbm_peak": b / ms / 1e6 / hbm_peak}
                       for k, (ms, b) in res.items()}}
    for k, n in evals.items():
        out["kernels"][k]["bound"] = "fp64 pipe (per-point transcendental evaluations), not HBM"
        out["kernels"][k]["evaluations_per_sec"] = n / out["kernels"][k]["ms"] * 1e3
    return out


# ------------------------------------------------------------------------------------------------
# the DTW-barycentre-averaging step that produces y_mean (models.py:176-178)
# ------------------------------------------------------------------------------------------------
def measure_dba(be, r_dev, cfg, step_ms, max_iter, with_cpu):
    """be_dtw_barycenter_averaging_subgradient (max_iter=50, tol=1e-3 as the reference calls it) on the
    step's (cell, member) problems; CUDA events; per-kernel figures from the C ABI profiler.  The CPU
    figure is the C oracle (oracle/dba.c, one thread per member, all host cores)."""
    import torch

    C, M, R, T = r_dev.shape
    X = r_dev.reshape(C * M, R, T)
    be.dtw_barycenter_averaging_subgradient(X, max_iter=2, tol=1e-3)  # warm-up (workspace)
    torch.cuda.synchronize()
    be.profile(True)
    be.profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, n_iter, _ = be.dtw_barycenter_averaging_subgradient(X, max_iter=max_iter, tol=1e-3, want_info=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    prof = be.profile_read()
    be.profile(False)
    dp = prof.get("k_dtw_dp")
    out = {
        "ms_per_step": ms, "problems": C * M, "series_per_problem": R, "time_steps": T, "max_iter": max_iter, "tol": 1e-3,
        "iterations_mean": float(n_iter.float().mean()), "iterations_max": int(n_iter.max()),
        "cells_per_sec_dba_only": C / ms * 1e3,
        "cells_per_sec_fit_weight_barycentre_with_dba": C / (ms + step_ms) * 1e3,
        "kernels": {k: {"ms": v["ms"], "launches": v["launches"]} for k, v in prof.items()
                    if k in ("k_dtw_dp", "k_dtw_backtrack", "k_dba_update")},
        "k_dtw_dp": None if not dp else {
            "dtw_table_cells_per_sec": dp["flops"] / 5.0 / dp["ms"] * 1e3, "fp64_ops_per_cell": 5,
            "tflops": dp["flops"] / dp["ms"] / 1e9,
            "bound": "instruction issue: ~14 instructions per table cell (5 on the FP64 pipe); no tensor-core form "
                     "exists for a min-plus recurrence, and the 2-bit-per-cell path record is 0.25 B of HBM per cell"},
        "note": "optional stage in front of the timed step (y_mean = 'dba' in grid.fit_weight_barycentre / GPDTW1D); "
                "the headline value keeps y_mean as an input, as SURVEY 8 scopes it",
    }
    if with_cpu:
        from concurrent.futures import ThreadPoolExecutor

        from oracle import dba as oracle_dba

        host = r_dev[0].cpu().numpy()
        n = min(os.cpu_count() or 1, M)
        oracle_dba.dba_subgradient(host[0][:, :64], max_iter=1)  # loads the library
        t0 = time.perf_counter()
        with ThreadPoolExecutor(n) as ex:
            its = list(ex.map(lambda m: oracle_dba.dba_subgradient(host[m], max_iter=max_iter, tol=1e-3)[1], range(n)))
        dt = time.perf_counter() - t0
        out["cpu"] = {"value": 1.0 / (dt * M / n), "unit": UNIT, "cores": n, "kind": "port",
                      "sample": f"{n} of {M} members of one cell, one thread each (C oracle): {dt:.1f} s, "
                                f"{float(np.mean(its)):.0f} iterations"}
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
TENSOR_FAMILIES = {"k_chol_update", "k_trtri_accum", "k_lauum_cov", "k_panel_scale", "k_diag_block",
                   "k_small_factor_inverse", "k_small_cov_factor"}

# ------------------------------------------------------------------------------------------------
# one workload, timed: device-resident steps (CUDA events, per-kernel profile) and end-to-end steps
# ------------------------------------------------------------------------------------------------
def time_workload(be, cfg, cps, steps, warmup, e2e_steps, rank, world, barrier, max_over_ranks, sample_clocks=None):
    """Returns a dict with ms (max over ranks, for ``steps`` device-resident steps), launches, the per-kernel
    profile, the last result, and the end-to-end seconds for ``e2e_steps`` steps from pinned host buffers."""
    import torch

    from bayesian_ensembling_b200 import grid, synthetic

    dev = be.device
    reals, obs = synthetic.make_cells(cfg, n_cells=cps, cell_offset=rank * cps)
    var, ls = synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE
    r_dev = torch.as_tensor(reals, device=dev)
    o_dev = torch.as_tensor(obs, device=dev)

    def step_device():
        return grid.fit_weight_barycentre(r_dev, o_dev, var, ls, cells_per_wave=cps)

    for _ in range(warmup):
        res = step_device()
    barrier()
    assert int(res.info_fit.abs().sum()) == 0 and int(res.info_dist.abs().sum()) == 0, "non-PD matrix in the bench"
    be.profile(True)
    be.profile_reset()
    sampler = sample_clocks() if sample_clocks else None
    l0 = be.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        res = step_device()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = be.launch_count - l0
    clocks = sampler.summary() if sampler else None
    prof = be.profile_read()
    be.profile(False)

    # end-to-end: pinned host buffers through the public batched API, H2D + D2H inside the timed region
    r_pin = torch.as_tensor(reals).pin_memory()
    o_pin = torch.as_tensor(obs).pin_memory()
    out_w = torch.empty((cps, cfg.members, cfg.steps), dtype=torch.float64).pin_memory()
    out_mu = torch.empty((cps, cfg.steps), dtype=torch.float64).pin_memory()
    out_sd = torch.empty((cps, cfg.steps), dtype=torch.float64).pin_memory()

    def step_e2e():
        # cells_per_wave left to the library: with host inputs it pipelines the host-to-device copies of the
        # next wave under the kernels of the current one
        r = grid.fit_weight_barycentre(r_pin, o_pin, var, ls)
        out_w.copy_(r.weights, non_blocking=True)
        out_mu.copy_(r.bary_mu, non_blocking=True)
        out_sd.copy_(r.bary_std, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    return {
        "ms": ms, "launches": launches, "prof": prof, "res": res, "clocks": clocks, "e2e_s": e2e_s,
        "h2d": (r_pin.numel() + o_pin.numel()) * 8, "d2h": (out_w.numel() + out_mu.numel() + out_sd.numel()) * 8,
        "nan_frac": float(np.isnan(out_w.numpy()).mean()), "reals": reals, "obs": obs, "r_dev": r_dev, "o_dev": o_dev,
    }


def stage_table(prof, steps, hbm_peak):
    total = max(sum(q["ms"] for q in prof.values()), 1e-30)
    stages = {}
    for name, p in prof.items():
        tensor = name in TENSOR_FAMILIES
        stages[name] = {
            "ms_per_step": p["ms"] / steps, "launches_per_step": p["launches"] / steps, "share": p["ms"] / total,
            "tflops": p["flops"] / p["ms"] / 1e9 if p["ms"] > 0 else None,
            "gbs": p["bytes"] / p["ms"] / 1e6 if p["ms"] > 0 else None,
            "bound": "tensor" if tensor else "hbm", "avg_launch_ms": p["ms"] / p["launches"],
        }
        if not tensor and p["ms"] > 0:
            stages[name]["frac_of_hbm_peak"] = p["bytes"] / p["ms"] / 1e6 / hbm_peak
    return stages


def tensor_stage(prof, cfg, cps, steps, ms):
    """All factorisation kernels together, and the whole step, against the FP64 tensor peak; the algorithmic
    flops are 4/3 T^3 per member on the REAL T."""
    tensor_ms = sum(p["ms"] for n, p in prof.items() if n in TENSOR_FAMILIES)
    tensor_flops = sum(p["flops"] for n, p in prof.items() if n in TENSOR_FAMILIES)
    algo = 4.0 / 3.0 * cfg.steps ** 3 * cfg.members * cps * steps
    return {
        "tflops": tensor_flops / tensor_ms / 1e9 if tensor_ms else None,
        "frac_of_peak": tensor_flops / tensor_ms / 1e9 / FP64_PEAK_TFLOPS if tensor_ms else None,
        "share_of_step": tensor_ms / max(sum(q["ms"] for q in prof.values()), 1e-30),
        "pipeline_tflops": algo / (ms * 1e-3) / 1e12,
        "pipeline_frac": algo / (ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
        "note": "tflops / frac_of_peak: all factorisation kernels (Cholesky x2, triangular inverse, lauum) over their own "
                "time, flops booked per launch on the real T; pipeline_*: 4/3 T^3 per member x members x cells over the "
                "WHOLE step time (gram, memory-bound stages and launch gaps included)",
    }



def measure_member_sharded(be, cfg, rank, world, barrier, max_over_ranks, reps=3):
    """SURVEY 8e for C < #GPUs: ONE cfg cell, its members split over the ranks, joined by the two small NCCL
    all-reduces of grid.fit_weight_barycentre_member_sharded (the normaliser sum_m w~, weights.py:122-123, then
    (sum_m w mu, sum_m w sigma), wasserstein.py:85-86,98).  Strong scaling: the same cell on one rank is the
    baseline.  CUDA events bracket each collective on the stream it is enqueued on."""
    import torch
    import torch.distributed as dist

    from bayesian_ensembling_b200 import grid, synthetic

    dev = be.device
    reals, obs = synthetic.make_cells(cfg, n_cells=1, cell_offset=0)  # the SAME cell on every rank
    M = cfg.members
    lo, hi = grid.shard_range(M, rank, world)
    var, ls = synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE
    r_loc = torch.as_tensor(reals[:, lo:hi], device=dev)
    o_dev = torch.as_tensor(obs, device=dev)
    coll = []

    def all_reduce(t):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        b.record()
        coll.append((a, b, t.numel() * 8))
        return t

    def once():
        return grid.fit_weight_barycentre_member_sharded(r_loc, o_dev, var, ls, all_reduce=all_reduce)

    for _ in range(2):
        res = once()
    barrier()
    coll.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(reps):
        res = once()
    e1.record()
    barrier()
    ms_cell = max_over_ranks(e0.elapsed_time(e1)) / reps
    ar_ms = [a.elapsed_time(b) for a, b, _ in coll]
    ar_bytes = [n for _, _, n in coll]
    ar1 = max_over_ranks(float(np.mean(ar_ms[0::2])))
    ar2 = max_over_ranks(float(np.mean(ar_ms[1::2])))
    # every rank must hold bit-identical barycentres (the all-reduce result is the same buffer everywhere)
    pack = torch.stack([res.bary_mu[0], res.bary_std[0]])
    gathered = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(gathered, pack)
    bitwise = all(torch.equal(torch.nan_to_num(g, nan=-7.0), torch.nan_to_num(gathered[0], nan=-7.0)) for g in gathered)
    # the one-rank baseline: all members of the cell on rank 0 (the other ranks wait at the barrier)
    single_ms, max_err = None, None
    if rank == 0:
        r_all = torch.as_tensor(reals, device=dev)
        for _ in range(2):
            ref = grid.fit_weight_barycentre(r_all, o_dev, var, ls)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            ref = grid.fit_weight_barycentre(r_all, o_dev, var, ls)
        s1.record()
        torch.cuda.synchronize()
        single_ms = s0.elapsed_time(s1) / reps

        def err(a, b):
            a, b = a.cpu().numpy(), b.cpu().numpy()
            if not (np.isnan(a) == np.isnan(b)).all():
                return float("inf")
            ok = ~np.isnan(b)
            return float(np.abs(a[ok] - b[ok]).max() / max(np.abs(b[ok]).max(), 1e-300)) if ok.any() else 0.0

        max_err = max(err(res.bary_mu, ref.bary_mu), err(res.bary_std, ref.bary_std),
                      err(res.mu[0], ref.mu[0, lo:hi]), err(res.weights[0], ref.weights[0, lo:hi]))
    barrier()
    if rank != 0:
        return None
    return {
        "workload": f"one {cfg.name} cell ({M} members x {cfg.realisations} realisations x {cfg.steps} steps), "
                    f"members sharded {world} ways ({hi - lo} on rank 0)",
        "ms_per_cell": ms_cell, "cells_per_sec": 1e3 / ms_cell, "ms_per_cell_one_gpu": single_ms,
        "strong_scaling_speedup": single_ms / ms_cell, "strong_scaling_efficiency": single_ms / ms_cell / world,
        "all_reduce_1_normaliser": {"ms": ar1, "bytes": ar_bytes[0], "what": "sum_m exp(c * mean ll) [C,T] (weights.py:122-123)"},
        "all_reduce_2_barycentre": {"ms": ar2, "bytes": ar_bytes[1], "what": "[3,C,T] (1, sum_m w mu, sum_m w sigma) (wasserstein.py:85-86,98)"},
        "collective_share_of_cell": (ar1 + ar2) / ms_cell,
        "ranks_bitwise_equal": bool(bitwise), "max_rel_err_vs_one_gpu": max_err,
        "backend": "NCCL all-reduce (fp64 sum) over NVLink; both messages are latency-sized, so the cell time is bounded "
                   "by SM under-fill (members per rank) plus two collective latencies, not by link bandwidth",
    }


def measure_reference_api(be, cfg, reals, obs):
    """The same cell through the REFERENCE-SHAPED API, host arrays in and host arrays out:
    ModelCollection.fit(GPDTW1D) -> LogLikelihoodWeight() -> Barycentre() (utils.py:102-135).  ModelCollection.fit
    leaves a Distribution with host mu / covariance per member (data.py:392-395), i.e. members x T^2 x 8 bytes of
    device-to-host traffic, inside the timed region."""
    import bayesian_ensembling_b200 as es
    from bayesian_ensembling_b200 import synthetic
    from bayesian_ensembling_b200.labelled import DataArray

    M, R, T, Ro = cfg.members, cfg.realisations, cfg.steps, cfg.obs_realisations
    tcoord = np.arange(T)

    def once():
        pms = [es.ProcessModel(DataArray(reals[0, m], ("realisation", "time"), {"realisation": np.arange(R), "time": tcoord}),
                               f"model{m}") for m in range(M)]
        obs_pm = es.ProcessModel(DataArray(obs[0], ("realisation", "time"), {"realisation": np.arange(Ro), "time": tcoord}), "obs")
        mc = es.ModelCollection(pms)
        mc.fit(es.GPDTW1D(hyperparameters=(synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE), y_mean="mean"),
               progress_bar=False)
        w = es.LogLikelihoodWeight()(mc, obs_pm)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            bary = es.Barycentre()(mc, w)
        return np.asarray(bary.mean.values), mc

    once()  # warm-up: pinned staging buffers, workspace
    t0 = time.perf_counter()
    _, mc = once()
    dt = time.perf_counter() - t0
    d2h = M * (T * T + T) * 8 + (M + 2) * T * 8
    return {"value": 1.0 / dt, "unit": UNIT, "seconds_per_cell": dt, "h2d_bytes_per_cell": (M * R + Ro) * T * 8,
            "d2h_bytes_per_cell": d2h,
            "note": "one cell; fixed hyper-parameters (GPDTW1D(hyperparameters=...), y_mean='mean') so that the work equals "
                    "the headline step's; the covariance of every member crosses PCIe as the reference's Distribution "
                    "holds it on the host"}


def measure_svgp(be, with_cpu, n_steps=10):
    """the SVGP stage of GPDTW3D.fit (models.py:357-411) at the reference's own sizes -- 400 inducing
    inputs, minibatches of 500, 10 realisation columns -- on N = 20 000 (t, lat, lon) points: ms per optimisation step
    (natural-gradient half + Adam half) and the prediction over all points; the oracle (NumPy / SciPy, all host BLAS
    threads) beside it."""
    import torch

    from bayesian_ensembling_b200.models import GPDTW3D

    rng = np.random.default_rng(20240 + 357)
    N, R, M, batch = 20000, 10, 400, 500
    lat, lon, t = rng.uniform(-85, 85, N), rng.uniform(0, 360, N), rng.uniform(-1, 1, N)
    X = np.column_stack([np.cos(np.radians(lat)) * np.cos(np.radians(lon)), np.cos(np.radians(lat)) * np.sin(np.radians(lon)),
                         np.sin(np.radians(lat)), t, 0.4 * t[:, None] + 0.1 * rng.standard_normal((N, R))])
    Y = np.column_stack([0.4 * t + 0.2 * np.sin(np.radians(lat)) + 0.05 * rng.standard_normal(N), rng.uniform(0.005, 0.03, N)])
    Z0 = np.linspace(X.min(axis=0), X.max(axis=0), M)
    idx = GPDTW3D.minibatch_order(N, batch, 2 * (n_steps + 1), 1)
    Xd, Yd, Zd = (torch.as_tensor(a, device=be.device) for a in (X, Y, Z0))
    be.svgp_fit(Xd, Yd, Zd, idx, 1)  # warm-up (workspace)
    torch.cuda.synchronize()
    times = []
    for k in (1, 1 + n_steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = be.svgp_fit(Xd, Yd, Zd, idx, k)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms_step = (times[1] - times[0]) / n_steps
    res = {"points": N, "inducing": M, "minibatch": batch, "realisation_columns": R, "ms_per_step": ms_step,
           "ms_predict_all_points_plus_one_step": times[0], "info": int(out["info"].item()),
           "steps_per_sec": 1e3 / ms_step,
           "note": "be_svgp_fit: natural-gradient step on one minibatch + Adam step (8 kernel parameters and the 400 x 14 "
                   "inducing inputs) on the next; M x M factorisations on the blocked DMMA path, rectangular products on a "
                   "warp-per-tile DMMA GEMM reading its fragments from global memory; one step is replayed from a CUDA graph and is bound by the latency of ONE-problem factorisation kernels (thirteen 128-wide diagonal blocks per step), not by launches"}
    if with_cpu:
        from oracle import svgp as osvgp  # the CPU leg: the oracle is what is timed here, never the product

        _use_all_host_threads()
        threads, _ = _blas_threads()
        t0 = time.perf_counter()
        osvgp.svgp_fit(X, Y, 0, n_inducing=M, minibatch_size=batch, seed=1)
        t1 = time.perf_counter()
        osvgp.svgp_fit(X, Y, 2, n_inducing=M, minibatch_size=batch, seed=1)
        t2 = time.perf_counter()
        cpu_step = max((t2 - t1) - (t1 - t0), 1e-9) / 2
        res["cpu"] = {"ms_per_step": cpu_step * 1e3, "cores": threads, "kind": "port", "gpu_over_cpu": cpu_step * 1e3 / ms_step,
                      "sample": "two steps of the oracle minus its prediction-only run"}
    return res


SIDE_CONFIGS = {"cfg1": 512, "cfg3": 6, "cfg4": 256}  # cells per step per GPU of the short side runs


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    from bayesian_ensembling_b200 import grid, synthetic
    from bayesian_ensembling_b200.backend import Backend

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = Backend.get()
    dev = be.device
    cps = args.cells_per_step
    var, ls = synthetic.L1_VARIANCE, synthetic.L1_LENGTHSCALE

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def clock_sampler():
        c = ClockSampler(local)
        c.start()
        return c

    # ---- the headline workload: device-resident value and end-to-end -----------------------------------
    hw = time_workload(be, cfg, cps, args.steps, args.warmup, args.steps, rank, world, barrier, max_over_ranks,
                       sample_clocks=clock_sampler)
    ms, prof, res, clocks = hw["ms"], hw["prof"], hw["res"], hw["clocks"]
    reals, obs, r_dev, o_dev = hw["reals"], hw["obs"], hw["r_dev"], hw["o_dev"]
    value = world * cps * args.steps / (ms * 1e-3)
    e2e_value = world * cps * args.steps / hw["e2e_s"]
    h2d, d2h, nan_frac = hw["h2d"], hw["d2h"], hw["nan_frac"]

    # ---- roofline of the dominant kernel -------------------------------------------------------
    hbm_peak, hbm_src = _hbm_peak()
    stages = stage_table(prof, args.steps, hbm_peak)
    top = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
    roofline = None
    tstage = tensor_stage(prof, cfg, cps, args.steps, ms)
    if top:
        p = prof[top]
        traffic = None
        try:
            # measured once with ncu (dram__bytes_read.sum + dram__bytes_write.sum, averaged per launch of this
            # kernel over one step of this workload): tools/gpu_traffic.sh -> profiles/traffic.json
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["bytes_per_launch"].get(top)
        except Exception:
            pass
        if top in TENSOR_FAMILIES:
            achieved = p["flops"] / p["ms"] / 1e9
            roofline = {"kernel": top, "bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS,
                        "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS, "traffic": traffic,
                        "peak_source": "FP64 DMMA issue-rate measured on this pool by tools/ubench_fp64.cu "
                                       "(profiles/r01_ubench_fp64.txt); MEASURED_PEAKS.json carries no fp64 figure; "
                                       "cuBLAS DGEMM 8192^3 measured 35.5 (profiles/r01_peak_fp64.json)",
                        "flops_per_launch": p["flops"] / p["launches"], "avg_launch_ms": p["ms"] / p["launches"],
                        "flops_booked_on": "the real T (not the padded Tp)",
                        "share_of_step": stages[top]["share"], "pipeline_frac": tstage["pipeline_frac"],
                        "pipeline_tflops": tstage["pipeline_tflops"]}
        else:
            achieved = p["bytes"] / p["ms"] / 1e6
            roofline = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json ({hbm_src})",
                        "bytes_per_launch": p["bytes"] / p["launches"], "avg_launch_ms": p["ms"] / p["launches"],
                        "share_of_step": stages[top]["share"], "pipeline_frac": tstage["pipeline_frac"],
                        "pipeline_tflops": tstage["pipeline_tflops"]}
    total_launches = sum_over_ranks(float(hw["launches"]))

    # ---- the member-sharded cell: the one place the path has a collective (N > 1 only) ------------------------
    member_sharded = None
    if world > 1 and not args.no_member_sharded:
        member_sharded = measure_member_sharded(be, cfg, rank, world, barrier, max_over_ranks)

    # ---- the other BASELINE shapes, short steps (north-star target: cfg3 on 8 GPUs) ----------------------------
    side = {}
    if not args.no_side_configs:
        for name, side_cps in SIDE_CONFIGS.items():
            if name == cfg.name:
                continue
            scfg = synthetic.CONFIGS[name]
            sw = time_workload(be, scfg, side_cps, args.side_steps, 2, 2, rank, world, barrier, max_over_ranks)
            if rank != 0:
                continue
            sstages = stage_table(sw["prof"], args.side_steps, hbm_peak)
            entry = {
                "workload": f"{scfg.name}: {scfg.description}", "cells_per_step_per_gpu": side_cps, "steps": args.side_steps,
                "value": world * side_cps * args.side_steps / (sw["ms"] * 1e-3), "unit": UNIT,
                "ms_per_step": sw["ms"] / args.side_steps,
                "e2e": {"value": world * side_cps * 2 / sw["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": sw["h2d"],
                        "d2h_bytes_per_step": sw["d2h"]},
                "weights_nan_fraction": sw["nan_frac"],
                "fp64_tensor_stage": tensor_stage(sw["prof"], scfg, side_cps, args.side_steps, sw["ms"]),
                "stages": {k: {kk: v[kk] for kk in ("ms_per_step", "share", "tflops", "gbs", "frac_of_hbm_peak") if kk in v}
                           for k, v in sstages.items()},
            }
            if world == 1 and not args.no_cpu_baseline:
                _use_all_host_threads()
                threads, blas = _blas_threads()
                if scfg.steps < 1000:
                    cell_s, procs, sample = cpu_small_t_cell_seconds(scfg, rounds=1)
                    entry["cpu_baseline"] = {"value": 1.0 / cell_s, "unit": UNIT, "cores": procs, "kind": "port",
                                             "sample": sample, "blas": blas + ", 1 thread per process"}
                else:
                    m = 3
                    dt = cpu_sample_seconds(scfg, m)
                    entry["cpu_baseline"] = {"value": 1.0 / (dt * scfg.members / m), "unit": UNIT, "cores": threads,
                                             "kind": "port", "blas": blas,
                                             "sample": f"{m} of {scfg.members} members of one {scfg.name} cell: {dt:.1f} s, "
                                                       f"scaled x{scfg.members / m:.0f}"}
            side[name] = entry
            del sw
            torch.cuda.empty_cache()

    # ---- L2: the natgrad + Adam training loop GPDTW1D.fit actually runs (models.py:208-215) ---------
    l2 = None
    if args.l2_iters > 0:
        Bm = cfg.members
        X, ym, yv = be.gpdtw1d_inputs(r_dev[0])
        be.vgp_fit(X, ym, yv, 1, want_scale_tri=False)  # warm-up (graph instantiation, workspace)
        torch.cuda.synchronize()
        t_it = []
        for n_it in (1, 1 + args.l2_iters):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            be.vgp_fit(X, ym, yv, n_it, want_scale_tri=False)
            f1.record()
            torch.cuda.synchronize()
            t_it.append(f0.elapsed_time(f1))
        ms_iter = (t_it[1] - t_it[0]) / args.l2_iters
        l2 = {"ms_per_iteration_per_cell": ms_iter, "iterations_timed": args.l2_iters, "members": Bm,
              "fixed_cost_ms_per_cell": t_it[0] - ms_iter,
              "cells_per_sec_at_2000_iterations": 1e3 / (t_it[0] - ms_iter + 2000 * ms_iter),
              "note": "be_vgp_fit on one cell (24 members batched): natural-gradient step + Adam step per iteration, "
                      "CUDA-graph replay; 2000 iterations is what experiments/full_experiment_script.py:87-113 uses"}
        # algorithmic flops of one iteration per member, on the real T: chol(K) 1/3, G = L^T D^-1 L 1/3, S = P^-1 (potrf,
        # trtri, lauum) 1, tril(G S) 1, L^-1 1/3, L^-T Phi^T 1/3, Kbar (triangular x triangular, full) 2/3  = 4 T^3
        l2_flops = 4.0 * float(cfg.steps) ** 3 * Bm
        l2["algorithmic_flops_per_iteration"] = l2_flops
        l2["tflops"] = l2_flops / (ms_iter * 1e-3) / 1e12
        l2["frac_of_fp64_tensor_peak"] = l2["tflops"] / FP64_PEAK_TFLOPS
        # the same loop at the small-T shape (cfg4 members, T = 251: the size of the reference's own per-member fits)
        cfg_s = synthetic.CONFIGS["cfg4"]
        rs, _ = synthetic.make_cells(cfg_s, n_cells=16)
        rs_dev = torch.as_tensor(rs, device=be.device)
        Xs, yms, yvs = be.gpdtw1d_inputs(rs_dev.reshape(-1, cfg_s.realisations, cfg_s.steps))
        be.vgp_fit(Xs, yms, yvs, 1, want_scale_tri=False)
        torch.cuda.synchronize()
        ts_it, n_small = [], 10
        for n_it in (1, 1 + n_small):
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            be.vgp_fit(Xs, yms, yvs, n_it, want_scale_tri=False)
            f1.record()
            torch.cuda.synchronize()
            ts_it.append(f0.elapsed_time(f1))
        ms_small = (ts_it[1] - ts_it[0]) / n_small
        nm = int(Xs.shape[0])
        l2["small_t"] = {"members": nm, "time_steps": cfg_s.steps, "ms_per_iteration": ms_small,
                         "member_iterations_per_sec": nm / ms_small * 1e3,
                         "tflops": 4.0 * float(cfg_s.steps) ** 3 * nm / (ms_small * 1e-3) / 1e12,
                         "note": "640 cfg4 members batched; the loop runs on the blocked path at every T (DESIGN 10 item 4)"}
        del rs_dev, Xs, yms, yvs
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            # the same iteration in the oracle (NumPy/SciPy, all host BLAS threads): ONE member, (fit with one
            # iteration) - (fit with none), scaled to the cell's members
            from oracle import reference_path as rp

            _use_all_host_threads()
            threads, _ = _blas_threads()
            t0 = time.perf_counter()
            rp.gpdtw1d_fit(reals[0, 0], n_optim_nits=0)
            t1 = time.perf_counter()
            rp.gpdtw1d_fit(reals[0, 0], n_optim_nits=1)
            t2 = time.perf_counter()
            cpu_it = max((t2 - t1) - (t1 - t0), 1e-9)
            l2["cpu"] = {"seconds_per_iteration_per_cell": cpu_it * Bm, "cores": threads, "kind": "port",
                         "sample": f"1 of {Bm} members, one iteration: {cpu_it:.1f} s, scaled x{Bm}",
                         "gpu_over_cpu": cpu_it * Bm / (ms_iter * 1e-3)}

    # ---- memory-bound stages on their own, at a size >> L2 and with FINITE weights ----------------
    # (inside the cfg2 step they see 6 x 3012 points -- launch-latency sized -- and, at T=3012, NaN weights;
    #  the cfg4 side run above reports the same kernels INSIDE a pipeline step on data with mostly finite weights)
    hbm_stages = None
    if args.hbm_points > 0 and rank == 0:
        hbm_stages = measure_hbm_stages(be, cfg, args.hbm_points, hbm_peak)

    # ---- the same step with the posterior covariance kept in factored (Woodbury) form ------------------
    # (not the headline: BASELINE's config names a full-covariance posterior per member, which this mode never
    # forms; weights, barycentre, posterior mean and variance agree with the dense path to ~1e-14)
    factored = None
    if args.factored_steps > 0:
        def step_factored():
            return grid.fit_weight_barycentre(r_dev, o_dev, var, ls, cells_per_wave=cps, posterior="factored")

        for _ in range(2):
            rf = step_factored()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.factored_steps):
            rf = step_factored()
        g1.record()
        barrier()
        ms_f = max_over_ranks(g0.elapsed_time(g1))
        dmu = float((rf.mu - res.mu).abs().max() / res.mu.abs().max())
        dvar = float((rf.var_diag - res.var_diag).abs().max() / res.var_diag.abs().max())
        factored = {"value": world * cps * args.factored_steps / (ms_f * 1e-3), "unit": UNIT,
                    "ms_per_step": ms_f / args.factored_steps, "steps": args.factored_steps,
                    "tensor_flops_per_member": "T^3 (chol M, triangular inverse, chol N) instead of 4/3 T^3",
                    "max_rel_diff_vs_dense": {"posterior_mean": dmu, "posterior_variance": dvar},
                    "note": "grid.fit_weight_barycentre(..., posterior='factored') -> be_gp_posterior_factored; "
                            "optional mode, NOT the headline value"}

    dba = None
    if args.dba_iters > 0 and rank == 0:
        dba = measure_dba(be, r_dev, cfg, ms / args.steps, args.dba_iters, world == 1 and not args.no_cpu_baseline)

    svgp_stage = None
    if rank == 0 and not args.no_svgp:
        svgp_stage = measure_svgp(be, world == 1 and not args.no_cpu_baseline)

    ref_api = None
    if rank == 0 and not args.no_reference_api:
        del r_dev, o_dev, res
        torch.cuda.empty_cache()
        ref_api = measure_reference_api(be, cfg, reals, obs)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _use_all_host_threads()
        threads, blas = _blas_threads()
        if cfg.steps < 1000 and not args.cpu_members:
            cell_s, procs, sample = cpu_small_t_cell_seconds(cfg, rounds=4)
            cpu_baseline = {"value": 1.0 / cell_s, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample,
                            "blas": blas + ", 1 thread per process", "host_cpus": os.cpu_count()}
        else:
            m = min(cfg.members, args.cpu_members or 6)
            dt = cpu_sample_seconds(cfg, m)
            cpu_baseline = {
                "value": 1.0 / (dt * cfg.members / m), "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{m} of {cfg.members} members of one {cfg.name} cell (T={cfg.steps}, "
                          f"Ro={cfg.obs_realisations}): {dt:.1f} s, scaled x{cfg.members / m:.0f} to a cell",
                "blas": blas, "host_cpus": os.cpu_count(),
                "algorithm_note": "about 80 % of the oracle's time per member is the reference's own R_o x T right-hand-side "
                                  "triangular solve for the constant-vector log-density (weights.py:97-100, R_o T^3 flops); "
                                  "the CUDA path carries L^-1 1 and L^-1 mu through the factorisation (O(T^2)) and forms "
                                  "the posterior in 4/3 T^3 instead of 8/3 T^3 flops: roughly one order of magnitude of "
                                  "the GPU/CPU ratio is algorithm, not hardware"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, cps, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "weights_nan_fraction": nan_frac,
                    "note": "at T=3012 every member's constant-vector log-likelihood is < -745, so exp() underflows "
                            "and the reference's un-guarded normalisation gives 0/0 = NaN (quirk Q-EXP, "
                            "weights.py:107,122-123) -- reproduced, not repaired"},
            "e2e_reference_api": ref_api,
            "gpu_launches": int(total_launches),
            "roofline": roofline,
            "fp64_tensor_stage": tstage,
            "stages": stages,
            "member_sharded": member_sharded,
            "configs": side or None,
            "l2_training_loop": l2,
            "hbm_stages": hbm_stages,
            "factored_posterior": factored,
            "dtw_barycentre_averaging": dba,
            "svgp_stage": svgp_stage,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _stdout_json_only():
    """C libraries (NCCL prints its version banner with printf) share fd 1 with us; the contract is ONE JSON
    line on stdout.  fd 1 is pointed at stderr and Python's sys.stdout keeps the real stdout."""
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    _stdout_json_only()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--cells-per-step", type=int, default=6)
    ap.add_argument("--cpu-members", type=int, default=0,
                    help="members of one cell the CPU arm times per step with threaded BLAS (0: 6 for T >= 1000 -- about "
                         "12 s of CPU work at cfg2 on 16 threads; for T < 1000 whole cells on one process per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hbm-points", type=int, default=4_000_000,
                    help="(cell, time) points of the stand-alone memory-bound stage measurements (0: skip)")
    ap.add_argument("--factored-steps", type=int, default=3,
                    help="steps timed for the factored_posterior line (0: skip)")
    ap.add_argument("--dba-iters", type=int, default=50,
                    help="max_iter of the stand-alone DTW-barycentre-averaging measurement (0: skip)")
    ap.add_argument("--l2-iters", type=int, default=3, help="training-loop iterations timed for the l2_training_loop line (0: skip)")
    ap.add_argument("--side-steps", type=int, default=3, help="timed steps of each side config (cfg1 / cfg3 / cfg4 short runs)")
    ap.add_argument("--no-side-configs", action="store_true", help="skip the cfg1 / cfg3 / cfg4 short runs")
    ap.add_argument("--no-member-sharded", action="store_true", help="skip the member-sharded cell (N > 1)")
    ap.add_argument("--no-reference-api", action="store_true", help="skip the e2e_reference_api measurement")
    ap.add_argument("--no-svgp", action="store_true", help="skip the svgp_stage measurement (GPDTW3D's SVGP step)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    from bayesian_ensembling_b200 import synthetic

    cfg = synthetic.CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
