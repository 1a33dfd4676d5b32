"""Batched entry point: fit -> weight -> barycentre for many grid cells at once.

The reference fits one member at a time in a Python loop (ensembles/data.py:391-395) and has
no per-cell driver at all; a B200 needs thousands of independent (cell, member) problems in
flight, so this module adds the batched call SURVEY 8b asks for.  Cells never interact, so
work is sharded across GPUs BY CELL with no collective; when there are fewer cells than GPUs
(BASELINE configs 1, 2, 5) members are sharded instead and ONE all-reduce of three partial
sums per (cell, time) joins them.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .backend import Backend, DEFAULT_JITTER


@dataclass
class CellBatchResult:
    weights: torch.Tensor  # [C,M,T]  LogLikelihoodWeight
    bary_mu: torch.Tensor  # [C,T]    Barycentre mean
    bary_std: torch.Tensor  # [C,T]   Barycentre std (the reference stores std**2 as "covariance")
    bary_iters: torch.Tensor  # [C,T] int32
    mu: torch.Tensor  # [C,M,T] posterior means
    var_diag: torch.Tensor  # [C,M,T] posterior variances
    info_fit: torch.Tensor  # [C,M] int32
    info_dist: torch.Tensor  # [C,M] int32
    cov: torch.Tensor | None = None  # [C,M,T,T] if keep_posteriors
    scale_tri: torch.Tensor | None = None


def shard_range(n: int, rank: int, world: int):
    """Static block partition of ``n`` uniform work items (cells or members)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _per_problem(x, C, M, device):
    """scalar | [M] | [C,M]  ->  [C*M] device tensor"""
    t = torch.as_tensor(x, dtype=torch.float64, device=device)
    if t.ndim == 0:
        t = t.expand(C, M)
    elif t.ndim == 1:
        t = t[None, :].expand(C, M)
    return t.reshape(C * M).contiguous()


HOST_PIPELINE_MIN_CELLS = 16  # below this a wave is too small to be worth splitting for copy / compute overlap


def _is_host(x) -> bool:
    return not (isinstance(x, torch.Tensor) and x.is_cuda)


def _as_f64(x) -> torch.Tensor:
    """host array / tensor -> contiguous fp64 host tensor (no copy when it already is one, e.g. a pinned buffer)"""
    if not isinstance(x, torch.Tensor):
        if hasattr(x, "flags") and not x.flags.writeable:
            x = x.copy()
        x = torch.as_tensor(x)
    return x.to(torch.float64).contiguous()


class _HostPrefetcher:
    """Wave-by-wave host-to-device staging of the realisations on a side stream, one wave ahead of the compute
    stream (double buffered): the copy of wave k+1 overlaps the kernels of wave k.  Pinned host memory makes the
    copies truly asynchronous; pageable memory still works (the copy then blocks the host, not the device)."""

    def __init__(self, be: Backend, r_host: torch.Tensor, waves):
        self.be, self.r, self.waves = be, r_host, waves
        # one copy stream per backend, kept: a fresh stream per call would give every staging buffer its own
        # allocator pool (a cudaMalloc per call)
        if getattr(be, "_copy_stream", None) is None:
            be._copy_stream = torch.cuda.Stream(device=be.device)
        self.stream = be._copy_stream
        self.pending = {}
        self._issue(0)

    def _issue(self, wi):
        if wi >= len(self.waves) or wi in self.pending:
            return
        c0, c1 = self.waves[wi]
        cur = torch.cuda.current_stream(self.be.device)
        # the staging buffer comes from the compute stream's pool (reused from step to step); the copy stream waits
        # for whatever the compute stream last did with that block, then fills it
        buf = torch.empty((c1 - c0,) + tuple(self.r.shape[1:]), dtype=torch.float64, device=self.be.device)
        free_ev = torch.cuda.Event()
        free_ev.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(free_ev)
            buf.copy_(self.r[c0:c1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        buf.record_stream(self.stream)
        self.pending[wi] = (buf, ev)

    def get(self, wi):
        self._issue(wi)
        buf, ev = self.pending.pop(wi)
        torch.cuda.current_stream(self.be.device).wait_event(ev)
        self._issue(wi + 1)  # the next wave's copy runs under this wave's kernels
        return buf


def wave_size(be: Backend, C: int, M: int, R: int, T: int, keep_posteriors: bool, budget_bytes: int | None = None):
    """Cells per wave so that workspace + outputs stay inside the memory budget (180 GB HBM3e
    holds ~1400 T=1980 problems with both work matrices resident).  The driver's free-memory query costs
    ~20 ms on a B200 with a large pool (measured: profiles/r02k_bench.json e2e vs value), so its answer is kept per
    problem shape for the life of the process; a call that runs out of memory drops the cache and asks again."""
    key = (C, M, R, T, keep_posteriors)
    cache = be.__dict__.setdefault("_wave_cache", {})
    if budget_bytes is None and key in cache:
        return cache[key]
    explicit = budget_bytes is not None
    if budget_bytes is None:
        free, _total = torch.cuda.mem_get_info(be.device)
        # memory torch's caching allocator holds but has not handed out is available to this call too (the
        # workspace of the previous wave / step lives there)
        cached = torch.cuda.memory_reserved(be.device) - torch.cuda.memory_allocated(be.device)
        budget_bytes = int((free + max(cached, 0)) * 0.8)
    per_cell = be.posterior_workspace_bytes(M, T, R) + (2 * M * T * T * 8 if keep_posteriors else 0) + 64 * M * T
    n = max(1, min(C, budget_bytes // max(per_cell, 1)))
    if not explicit:
        cache[key] = n
    return n


def fit_weight_barycentre(realisations, observations, variance, lengthscale, *, cells_per_wave=None, **kwargs):
    """See ``_fit_weight_barycentre``.  With ``cells_per_wave`` left to the library, a wave that runs out of device
    memory (the cached wave size is from an earlier, roomier moment) is retried once with a fresh memory query."""
    try:
        return _fit_weight_barycentre(realisations, observations, variance, lengthscale, cells_per_wave=cells_per_wave,
                                      **kwargs)
    except torch.cuda.OutOfMemoryError:
        if cells_per_wave is not None:
            raise
        be = Backend.get()
        be.__dict__.pop("_wave_cache", None)
        be._workspace = None
        torch.cuda.empty_cache()
        return _fit_weight_barycentre(realisations, observations, variance, lengthscale, cells_per_wave=None, **kwargs)


def _fit_weight_barycentre(realisations, observations, variance, lengthscale, *, jitter=DEFAULT_JITTER,
                           standardisation_constant=1.0, time_mean_weights=False, keep_posteriors=False,
                           cells_per_wave=None, tolerance=1e-6, init_var=1.0, y_mean="mean",
                           posterior="dense") -> CellBatchResult:
    """realisations [C,M,R,T], observations [C,Ro,T] (host arrays or device tensors);
    variance / lengthscale: scalar, [M] or [C,M] kernel hyper-parameters (fixed-theta posterior,
    the fixed point of models.py:208-215).  Order of operations follows
    ``PerfectModelTest._run_single_test`` (utils.py:102-135); ``time_mean_weights`` reproduces
    utils.py:111,133 (NaN-skipping mean over time, broadcast back).  ``y_mean``: "mean" = arithmetic
    mean over realisations (the y_mean-given path BASELINE's metric is quoted on), "dba" = the DTW
    barycentre average of models.py:176-178 computed on the device, or a [C,M,T] array.
    ``posterior``: "dense" forms every member's T x T covariance and its Cholesky factor as the
    reference does (data.py:38-39); "factored" keeps the covariance in Woodbury form and returns the
    same weights / barycentre / posterior mean and variance at 3/4 of the flops
    (``be_gp_posterior_factored``; not available with ``keep_posteriors``)."""
    if posterior not in ("dense", "factored"):
        raise ValueError(f"posterior must be 'dense' or 'factored', got {posterior!r}")
    if posterior == "factored" and keep_posteriors:
        raise ValueError("posterior='factored' does not form the dense covariance keep_posteriors asks for")
    be = Backend.get()
    host_in = _is_host(realisations)
    r = _as_f64(realisations) if host_in else be._in(realisations)
    o = be._in(observations)
    C, M, R, T = r.shape
    Ro = o.shape[1]
    var = _per_problem(variance, C, M, be.device)
    ls = _per_problem(lengthscale, C, M, be.device)
    if cells_per_wave is None:
        cells_per_wave = wave_size(be, C, M, R, T, keep_posteriors)
        if host_in and C >= 2 * HOST_PIPELINE_MIN_CELLS:
            # host inputs: at least two waves, so that wave k+1's host-to-device copy (side stream) runs under
            # wave k's kernels instead of in front of them
            cells_per_wave = min(cells_per_wave, max(HOST_PIPELINE_MIN_CELLS, (C + 3) // 4))
    waves = [(c0, min(C, c0 + cells_per_wave)) for c0 in range(0, C, cells_per_wave)]
    fetch = _HostPrefetcher(be, r, waves) if host_in else None
    outs = []
    for wi, (c0, c1) in enumerate(waves):
        Cw = c1 - c0
        r_w = fetch.get(wi) if host_in else r[c0:c1]
        X, ym, yv = be.gpdtw1d_inputs(r_w.reshape(Cw * M, R, T))
        if isinstance(y_mean, str):
            if y_mean == "dba":
                ym = be.dtw_barycenter_averaging_subgradient(r_w.reshape(Cw * M, R, T), max_iter=50, tol=1e-3)
            elif y_mean != "mean":
                raise ValueError(f"y_mean must be 'mean', 'dba' or an array, got {y_mean!r}")
        else:
            ym = be._in(y_mean, (C, M, T), "y_mean")[c0:c1].reshape(Cw * M, T)
        if posterior == "factored":
            post = be.gp_posterior_factored(X, ym, yv, var[c0 * M:c1 * M], ls[c0 * M:c1 * M], jitter)
        else:
            post = be.gp_posterior(X, ym, yv, var[c0 * M:c1 * M], ls[c0 * M:c1 * M], jitter,
                                   want_cov=keep_posteriors, want_scale_tri=keep_posteriors)
        w = be.loglik_weights_mvn(post.mvn_stats, o[c0:c1], M, standardisation_constant)
        w_used = be.weights_time_mean(w) if time_mean_weights else w
        mu3, var3 = post.mu.view(Cw, M, T), post.var_diag.view(Cw, M, T)
        bmu, bsd, bit = be.barycentre_1d(mu3, var3, w_used, tolerance, init_var, 200)
        outs.append(CellBatchResult(
            weights=w, bary_mu=bmu, bary_std=bsd, bary_iters=bit, mu=mu3, var_diag=var3,
            info_fit=post.info_fit.view(Cw, M), info_dist=post.info_dist.view(Cw, M),
            cov=post.cov.view(Cw, M, T, T) if keep_posteriors else None,
            scale_tri=post.scale_tri.view(Cw, M, T, T) if keep_posteriors else None))
    if len(outs) == 1:
        return outs[0]
    cat = lambda name: (None if getattr(outs[0], name) is None else torch.cat([getattr(x, name) for x in outs]))  # noqa: E731
    return CellBatchResult(**{k: cat(k) for k in CellBatchResult.__dataclass_fields__})


def fit_weight_barycentre_member_sharded(realisations_local, observations, variance_local, lengthscale_local, *,
                                         group=None, jitter=DEFAULT_JITTER, standardisation_constant=1.0,
                                         time_mean_weights=False, tolerance=1e-6, init_var=1.0,
                                         all_reduce=None, ops=None) -> CellBatchResult:
    """Member-sharded form for few-cell configs: every rank holds ``M_local`` members of ALL C
    cells.  Two small all-reduces (sum, fp64) join them -- first the normaliser ``sum_m w~ [C,T]``,
    then the packed ``[3,C,T] = (1, sum w mu, sum w sigma)`` of the NORMALISED weights -- NCCL over
    NVLink when launched under torchrun; ``all_reduce`` can be injected (tests use gloo on CPU
    tensors).  The time-mean of the weights (utils.py:111,133) sits between the two.  ``ops`` is the operator object (default: the CUDA
    ``Backend``); the CPU test-suite injects a stand-in to exercise this host logic over gloo."""
    import torch.distributed as dist

    be = Backend.get() if ops is None else ops
    r = be._in(realisations_local)
    o = be._in(observations)
    C, Ml, R, T = r.shape
    if all_reduce is None:
        def all_reduce(t):
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return t
    var = _per_problem(variance_local, C, Ml, be.device)
    ls = _per_problem(lengthscale_local, C, Ml, be.device)
    X, ym, yv = be.gpdtw1d_inputs(r.reshape(C * Ml, R, T))
    post = be.gp_posterior(X, ym, yv, var, ls, jitter, want_cov=False, want_scale_tri=False)
    _, lls_exp, _ = be.loglik_weights_mvn(post.mvn_stats, o, Ml, standardisation_constant, want_lls=True)
    mu3, var3 = post.mu.view(C, Ml, T), post.var_diag.view(C, Ml, T)
    # 1st exchange: the normaliser sum_m w~ (weights.py:122-123).  The weights are then formed exactly as the
    # reference forms them, w = w~ / total, BEFORE they multiply anything: with Q-EXP's un-shifted exp the
    # w~ are routinely denormal (1e-310), and sum(w~ mu) / sum(w~) would lose the bits that w~ / total keeps
    # (measured 3e-6 relative on the barycentre mean at T=12 when one packed all-reduce of un-normalised
    # partial sums was used).
    total = all_reduce(be.barycentre_1d_partial(mu3, var3, lls_exp)[0].clone())
    w = be.weights_normalise(lls_exp, total)
    w_used = be.weights_time_mean(w) if time_mean_weights else w  # utils.py:111,133
    # 2nd exchange: (sum_m w mu, sum_m w sigma) over the local members, packed in one buffer
    partial = all_reduce(be.barycentre_1d_partial(mu3, var3, w_used))
    partial[0].fill_(1.0)  # the weights are already normalised (time-mean weights are used as they are, utils.py:119)
    bmu, bsd, bit = be.barycentre_1d_finish(partial, tolerance, init_var, 200)
    return CellBatchResult(weights=w, bary_mu=bmu, bary_std=bsd, bary_iters=bit, mu=mu3, var_diag=var3,
                           info_fit=post.info_fit.view(C, Ml), info_dist=post.info_dist.view(C, Ml))
