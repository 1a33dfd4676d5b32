"""Device-side operator layer: torch tensors in, torch tensors out, arithmetic in the C ABI.

PyTorch is plumbing only here (device memory, the current CUDA stream, dtype/contiguity
checks).  Every numerical operation is a call into ``libbe_b200.so``.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import torch

from . import _lib

DEFAULT_JITTER = 1e-6  # gpflow.config.default_jitter()
_POISON = bool(os.environ.get("BE_B200_POISON_WORKSPACE"))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


@dataclass
class PosteriorBatch:
    """Result of ``Backend.gp_posterior`` for B = cells * members problems (all on device)."""

    mu: torch.Tensor  # [B,T]
    var_diag: torch.Tensor  # [B,T]
    mvn_stats: torch.Tensor  # [B,4]
    info_fit: torch.Tensor  # [B] int32
    info_dist: torch.Tensor  # [B] int32
    cov: torch.Tensor | None = None  # [B,T,T]
    scale_tri: torch.Tensor | None = None  # [B,T,T]


class Backend:
    """One per (process, device).  Calls are ordered on torch's current stream."""

    _instances: dict = {}

    @classmethod
    def get(cls, device=None) -> "Backend":
        if not torch.cuda.is_available():
            raise _lib.BackendError(
                "bayesian_ensembling_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback."
            )
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if dev.index not in cls._instances:
            cls._instances[dev.index] = cls(dev)
        return cls._instances[dev.index]

    def __init__(self, device: torch.device):
        self.lib = _lib.load_library()
        self.device = device
        self._workspace = None
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            rc = self.lib.be_ctx_create(device.index, ctypes.c_void_p(stream), ctypes.byref(handle))
        if rc != 0:
            raise _lib.BackendError(f"be_ctx_create failed ({rc})")
        self.ctx = handle

    # ------------------------------------------------------------------ plumbing
    def _sync_stream(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.lib.be_ctx_set_stream(self.ctx, ctypes.c_void_p(stream))

    def _ws(self, nbytes: int) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = None
            self._workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        if _POISON:  # test hook: every byte 0xFF = NaN doubles, so a read of unwritten workspace shows up
            self._workspace.fill_(255)
        return self._workspace

    def _in(self, t, shape=None, name="tensor"):
        if not isinstance(t, torch.Tensor):
            if hasattr(t, "flags") and not t.flags.writeable:  # broadcast views etc.: torch wants a writable buffer
                t = t.copy()
            t = torch.as_tensor(t, dtype=torch.float64)
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        if t.device != self.device:
            t = t.to(self.device)
        t = t.contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def _new(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def sync(self):
        _lib.check(self.ctx, self.lib.be_ctx_sync(self.ctx), "be_ctx_sync")

    @property
    def launch_count(self) -> int:
        return int(self.lib.be_ctx_launch_count(self.ctx))

    # ------------------------------------------------------------------ per-kernel timing
    def profile(self, on: bool = True):
        _lib.check(self.ctx, self.lib.be_ctx_profile_enable(self.ctx, int(on)), "be_ctx_profile_enable")

    def profile_reset(self):
        _lib.check(self.ctx, self.lib.be_ctx_profile_reset(self.ctx), "be_ctx_profile_reset")

    def profile_read(self) -> dict:
        """{kernel family: dict(ms, launches, flops, bytes)} since the last reset (synchronises)."""
        out = {}
        name = ctypes.create_string_buffer(64)
        ms, fl, by = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        n = ctypes.c_longlong()
        for f in range(self.lib.be_ctx_profile_families()):
            rc = self.lib.be_ctx_profile_get(self.ctx, f, name, 64, ctypes.byref(ms), ctypes.byref(n),
                                             ctypes.byref(fl), ctypes.byref(by))
            _lib.check(self.ctx, rc, "be_ctx_profile_get")
            if n.value:
                out[name.value.decode()] = dict(ms=ms.value, launches=n.value, flops=fl.value, bytes=by.value)
        return out

    def posterior_workspace_bytes(self, B, T, R) -> int:
        return int(self.lib.be_gp_posterior_workspace_bytes(B, T, R))

    # ------------------------------------------------------------------ a1
    def gpdtw1d_inputs(self, realisations, want_X=True):
        """[B,R,T] -> X [B,T,R] (None unless want_X), y_mean [B,T], y_var [B,T]   (models.py:175-182)"""
        r = self._in(realisations)
        B, R, T = r.shape
        X, ym, yv = (self._new(B, T, R) if want_X else None), self._new(B, T), self._new(B, T)
        self._sync_stream()
        rc = self.lib.be_gpdtw1d_inputs(self.ctx, _ptr(r), B, R, T, _ptr(X), _ptr(ym), _ptr(yv))
        _lib.check(self.ctx, rc, "be_gpdtw1d_inputs")
        return X, ym, yv

    def matern32_gram(self, X, variance, lengthscale):
        X = self._in(X)
        B, T, R = X.shape
        var = self._in(variance, (B,), "variance")
        ls = self._in(lengthscale, (B,), "lengthscale")
        K = self._new(B, T, T)
        self._sync_stream()
        rc = self.lib.be_matern32_gram(self.ctx, _ptr(X), B, T, R, _ptr(var), _ptr(ls), _ptr(K))
        _lib.check(self.ctx, rc, "be_matern32_gram")
        return K

    def potrf(self, A):
        A = self._in(A)
        B, T, _ = A.shape
        L = self._new(B, T, T)
        info = self._new(B, dtype=torch.int32)
        nbytes = int(self.lib.be_potrf_workspace_bytes(B, T))
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_potrf_batched(self.ctx, _ptr(A), B, T, _ptr(L), _ptr(info), _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_potrf_batched")
        return L, info

    def gp_posterior(self, X, y_mean, y_var, variance, lengthscale, jitter=DEFAULT_JITTER,
                     want_cov=True, want_scale_tri=True) -> PosteriorBatch:
        X = self._in(X)
        B, T, R = X.shape
        ym = self._in(y_mean, (B, T), "y_mean")
        yv = self._in(y_var, (B, T), "y_var")
        var = self._in(variance, (B,), "variance")
        ls = self._in(lengthscale, (B,), "lengthscale")
        out = PosteriorBatch(
            mu=self._new(B, T), var_diag=self._new(B, T), mvn_stats=self._new(B, 4),
            info_fit=self._new(B, dtype=torch.int32), info_dist=self._new(B, dtype=torch.int32),
            cov=self._new(B, T, T) if want_cov else None,
            scale_tri=self._new(B, T, T) if want_scale_tri else None,
        )
        nbytes = self.posterior_workspace_bytes(B, T, R)
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_gp_posterior(
            self.ctx, _ptr(X), _ptr(ym), _ptr(yv), _ptr(var), _ptr(ls), float(jitter), B, T, R,
            _ptr(out.mu), _ptr(out.var_diag), _ptr(out.cov), _ptr(out.scale_tri), _ptr(out.mvn_stats),
            _ptr(out.info_fit), _ptr(out.info_dist), _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_gp_posterior")
        return out

    def gp_posterior_factored(self, X, y_mean, y_var, variance, lengthscale, jitter=DEFAULT_JITTER) -> PosteriorBatch:
        """mu, var_diag, mvn_stats of the same posterior with the covariance kept in factored (Woodbury) form:
        what LogLikelihoodWeight and Barycentre consume, at T^3 instead of 4/3 T^3 tensor flops."""
        X = self._in(X)
        B, T, R = X.shape
        ym = self._in(y_mean, (B, T), "y_mean")
        yv = self._in(y_var, (B, T), "y_var")
        var = self._in(variance, (B,), "variance")
        ls = self._in(lengthscale, (B,), "lengthscale")
        out = PosteriorBatch(mu=self._new(B, T), var_diag=self._new(B, T), mvn_stats=self._new(B, 4),
                             info_fit=self._new(B, dtype=torch.int32), info_dist=self._new(B, dtype=torch.int32))
        nbytes = int(self.lib.be_gp_posterior_factored_workspace_bytes(B, T, R))
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_gp_posterior_factored(
            self.ctx, _ptr(X), _ptr(ym), _ptr(yv), _ptr(var), _ptr(ls), float(jitter), B, T, R,
            _ptr(out.mu), _ptr(out.var_diag), _ptr(out.mvn_stats), _ptr(out.info_fit), _ptr(out.info_dist),
            _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_gp_posterior_factored")
        return out

    def vgp_fit(self, X, y_mean, y_var, n_iters, gamma=0.5, lr=0.01, train_hypers=True, init_variance=1.0,
                init_lengthscale=1.0, jitter=DEFAULT_JITTER, want_scale_tri=True):
        """The natgrad + Adam loop of models.py:185-220 on the device.  Returns (PosteriorBatch,
        variance [B], lengthscale [B]) with the trained kernel hyper-parameters."""
        X = self._in(X)
        B, T, R = X.shape
        ym = self._in(y_mean, (B, T), "y_mean")
        yv = self._in(y_var, (B, T), "y_var")
        var = torch.full((B,), float(init_variance), dtype=torch.float64, device=self.device)
        ls = torch.full((B,), float(init_lengthscale), dtype=torch.float64, device=self.device)
        out = PosteriorBatch(
            mu=self._new(B, T), var_diag=self._new(B, T), mvn_stats=self._new(B, 4),
            info_fit=self._new(B, dtype=torch.int32), info_dist=self._new(B, dtype=torch.int32),
            cov=self._new(B, T, T), scale_tri=self._new(B, T, T) if want_scale_tri else None)
        nbytes = int(self.lib.be_vgp_fit_workspace_bytes(B, T, R))
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_vgp_fit(
            self.ctx, _ptr(X), _ptr(ym), _ptr(yv), B, T, R, int(n_iters), float(gamma), float(lr), int(bool(train_hypers)),
            float(jitter), _ptr(var), _ptr(ls), _ptr(out.mu), _ptr(out.var_diag), _ptr(out.cov), _ptr(out.scale_tri),
            _ptr(out.mvn_stats), _ptr(out.info_fit), _ptr(out.info_dist), _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_vgp_fit")
        return out, var, ls

    # ------------------------------------------------------------------ a3
    def mvn_from_cov(self, mu, cov, want_scale_tri=True):
        cov = self._in(cov)
        B, T, _ = cov.shape
        mu = self._in(mu, (B, T), "mu")
        tri = self._new(B, T, T) if want_scale_tri else None
        var_diag, stats = self._new(B, T), self._new(B, 4)
        info = self._new(B, dtype=torch.int32)
        nbytes = int(self.lib.be_mvn_from_cov_workspace_bytes(B, T))
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_mvn_from_cov(self.ctx, _ptr(mu), _ptr(cov), B, T, _ptr(tri), _ptr(var_diag), _ptr(stats),
                                      _ptr(info), _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_mvn_from_cov")
        return tri, var_diag, stats, info

    # ------------------------------------------------------------------ a4
    def loglik_weights_mvn(self, mvn_stats, obs, M, standardisation_constant=1.0, want_lls=False):
        """mvn_stats [C*M,4], obs [C,Ro,T] -> weights [C,M,T] (+ lls_exp, lls_mean)"""
        obs = self._in(obs)
        C, Ro, T = obs.shape
        st = self._in(mvn_stats, (C * M, 4), "mvn_stats")
        w = self._new(C, M, T)
        le = self._new(C, M, T) if want_lls else None
        lm = self._new(C, M, T) if want_lls else None
        self._sync_stream()
        rc = self.lib.be_loglik_weights_mvn(self.ctx, _ptr(st), _ptr(obs), C, M, Ro, T,
                                            float(standardisation_constant), _ptr(w), _ptr(le), _ptr(lm))
        _lib.check(self.ctx, rc, "be_loglik_weights_mvn")
        return (w, le, lm) if want_lls else w

    def mvn_constvec_logprob(self, mvn_stats, obs, M):
        obs = self._in(obs)
        C, Ro, T = obs.shape
        st = self._in(mvn_stats, (C * M, 4), "mvn_stats")
        ll = self._new(C, M, Ro, T)
        self._sync_stream()
        rc = self.lib.be_mvn_constvec_logprob(self.ctx, _ptr(st), _ptr(obs), C, M, Ro, T, _ptr(ll))
        _lib.check(self.ctx, rc, "be_mvn_constvec_logprob")
        return ll

    def mvn_log_prob(self, mu, scale_tri, x, sum_log_diag):
        """mu [T], scale_tri [T,T] (dense lower factor), x [N,T] -> log N(x_n | mu, L L^T) [N]: one forward
        substitution per vector against the stored factor (distrax MultivariateNormalTri.log_prob)."""
        mu = self._in(mu)
        T = mu.shape[0]
        L = self._in(scale_tri, (T, T), "scale_tri")
        x = self._in(x)
        N = x.shape[0]
        x = self._in(x, (N, T), "x")
        ll = self._new(N)
        self._sync_stream()
        rc = self.lib.be_mvn_log_prob(self.ctx, _ptr(mu), _ptr(L), _ptr(x), T, N, float(sum_log_diag), _ptr(ll))
        _lib.check(self.ctx, rc, "be_mvn_log_prob")
        return ll

    def normal_logprob(self, loc, scale, x):
        x = self._in(x)
        loc = self._in(loc).expand_as(x).contiguous()
        scale = self._in(scale).expand_as(x).contiguous()
        ll = torch.empty_like(x)
        self._sync_stream()
        rc = self.lib.be_normal_logprob(self.ctx, _ptr(loc), _ptr(scale), _ptr(x), x.numel(), _ptr(ll))
        _lib.check(self.ctx, rc, "be_normal_logprob")
        return ll

    def loglik_weights_normal(self, loc, scale, obs, standardisation_constant=1.0, want_lls=False):
        """loc/scale [C,M,N], obs [C,Ro,N] -> weights [C,M,N]"""
        loc = self._in(loc)
        C, M, N = loc.shape
        scale = self._in(scale, (C, M, N), "scale")
        obs = self._in(obs)
        Ro = obs.shape[1]
        w = self._new(C, M, N)
        le = self._new(C, M, N) if want_lls else None
        lm = self._new(C, M, N) if want_lls else None
        self._sync_stream()
        rc = self.lib.be_loglik_weights_normal(self.ctx, _ptr(loc), _ptr(scale), _ptr(obs), C, M, Ro, N,
                                               float(standardisation_constant), _ptr(w), _ptr(le), _ptr(lm))
        _lib.check(self.ctx, rc, "be_loglik_weights_normal")
        return (w, le, lm) if want_lls else w

    def weights_time_mean(self, weights):
        w = self._in(weights)
        C, M, T = w.shape
        out = torch.empty_like(w)
        self._sync_stream()
        rc = self.lib.be_weights_time_mean(self.ctx, _ptr(w), C, M, T, _ptr(out))
        _lib.check(self.ctx, rc, "be_weights_time_mean")
        return out

    def weights_normalise(self, lls_exp, total):
        le = self._in(lls_exp)
        C, M, T = le.shape
        total = self._in(total, (C, T), "total")
        w = torch.empty_like(le)
        self._sync_stream()
        rc = self.lib.be_weights_normalise(self.ctx, _ptr(le), _ptr(total), C, M, T, _ptr(w))
        _lib.check(self.ctx, rc, "be_weights_normalise")
        return w

    # ------------------------------------------------------------------ a5 / a6
    def barycentre_1d(self, means, variances, weights, tolerance=1e-6, init_var=1.0, max_iters=200):
        """[C,M,N] x3 -> mu [C,N], sigma [C,N], iters [C,N] (int32)"""
        means = self._in(means)
        C, M, N = means.shape
        variances = self._in(variances, (C, M, N), "variances")
        weights = self._in(weights, (C, M, N), "weights")
        mu, sigma = self._new(C, N), self._new(C, N)
        iters = self._new(C, N, dtype=torch.int32)
        self._sync_stream()
        rc = self.lib.be_barycentre_1d(self.ctx, _ptr(means), _ptr(variances), _ptr(weights), C, M, N,
                                       float(tolerance), float(init_var), int(max_iters), _ptr(mu), _ptr(sigma),
                                       _ptr(iters))
        _lib.check(self.ctx, rc, "be_barycentre_1d")
        return mu, sigma, iters

    def barycentre_1d_partial(self, means, variances, lls_exp):
        means = self._in(means)
        C, M, N = means.shape
        variances = self._in(variances, (C, M, N), "variances")
        lls_exp = self._in(lls_exp, (C, M, N), "lls_exp")
        partial = self._new(3, C, N)
        self._sync_stream()
        rc = self.lib.be_barycentre_1d_partial(self.ctx, _ptr(means), _ptr(variances), _ptr(lls_exp), C, M, N,
                                               _ptr(partial))
        _lib.check(self.ctx, rc, "be_barycentre_1d_partial")
        return partial

    def barycentre_1d_finish(self, partial, tolerance=1e-6, init_var=1.0, max_iters=200):
        partial = self._in(partial)
        _, C, N = partial.shape
        mu, sigma = self._new(C, N), self._new(C, N)
        iters = self._new(C, N, dtype=torch.int32)
        self._sync_stream()
        rc = self.lib.be_barycentre_1d_finish(self.ctx, _ptr(partial), C, N, float(tolerance), float(init_var),
                                              int(max_iters), _ptr(mu), _ptr(sigma), _ptr(iters))
        _lib.check(self.ctx, rc, "be_barycentre_1d_finish")
        return mu, sigma, iters

    # ------------------------------------------------------------------ a7 / a8 / a9
    SQRTM_TOL = 1e-10
    SQRTM_MAX_ITERS = 40

    def sqrtm_psd(self, A, want_inverse=False, tol=None, max_iters=None):
        """[B,T,T] symmetric positive definite -> principal square root (wasserstein.py:10-13).
        Returns (sqrt [B,T,T], inv_sqrt or None, iterations, info [B] int32)."""
        A = self._in(A)
        B, T, _ = A.shape
        out = self._new(B, T, T)
        inv = self._new(B, T, T) if want_inverse else None
        info = self._new(B, dtype=torch.int32)
        iters = ctypes.c_int(0)
        nbytes = int(self.lib.be_sqrtm_psd_workspace_bytes(B, T))
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_sqrtm_psd(self.ctx, _ptr(A), B, T, float(tol or self.SQRTM_TOL),
                                   int(max_iters or self.SQRTM_MAX_ITERS), _ptr(out), _ptr(inv),
                                   ctypes.cast(ctypes.byref(iters), ctypes.c_void_p), _ptr(info), _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_sqrtm_psd")
        return out, inv, int(iters.value), info

    def w2_distance(self, mu1, sigma1, mu2, sigma2):
        """P pairs of full-covariance Gaussians -> w2 [P] (wasserstein.py:21-47, quirk Q-W2)."""
        sigma1 = self._in(sigma1)
        P, T, _ = sigma1.shape
        sigma2 = self._in(sigma2, (P, T, T), "sigma2")
        mu1 = self._in(mu1, (P, T), "mu1")
        mu2 = self._in(mu2, (P, T), "mu2")
        w2 = self._new(P)
        info = self._new(P, dtype=torch.int32)
        nbytes = int(self.lib.be_w2_distance_workspace_bytes(P, T))
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_w2_distance(self.ctx, _ptr(mu1), _ptr(sigma1), _ptr(mu2), _ptr(sigma2), P, T,
                                     self.SQRTM_TOL, self.SQRTM_MAX_ITERS, _ptr(w2), _ptr(info), _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_w2_distance")
        return w2, info

    def w2_distance_diag(self, mu1, var1, mu2, var2):
        """full_cov=False branch (wasserstein.py:36-39): variances on a diagonal."""
        mu1 = self._in(mu1)
        P, T = mu1.shape
        var1 = self._in(var1, (P, T), "var1")
        mu2 = self._in(mu2, (P, T), "mu2")
        var2 = self._in(var2, (P, T), "var2")
        w2 = self._new(P)
        self._sync_stream()
        rc = self.lib.be_w2_distance_diag(self.ctx, _ptr(mu1), _ptr(var1), _ptr(mu2), _ptr(var2), P, T, _ptr(w2))
        _lib.check(self.ctx, rc, "be_w2_distance_diag")
        return w2

    def barycentre_fullcov(self, mus, sigmas, weights, tolerance=1e-6, init_var=1.0, max_iters=200):
        """mus [C,M,T], sigmas [C,M,T,T], weights [C,M] -> mu [C,T], S [C,T,T], iters [C] (host),
        info [C*M] (BASELINE config 5; the definition is in DESIGN.md 3.4)."""
        sigmas = self._in(sigmas)
        C, M, T, _ = sigmas.shape
        mus = self._in(mus, (C, M, T), "mus")
        weights = self._in(weights, (C, M), "weights")
        mu, S = self._new(C, T), self._new(C, T, T)
        info = self._new(C * M, dtype=torch.int32)
        iters = (ctypes.c_int * C)()
        nbytes = int(self.lib.be_barycentre_fullcov_workspace_bytes(C, M, T))
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_barycentre_fullcov(self.ctx, _ptr(mus), _ptr(sigmas), _ptr(weights), C, M, T,
                                            float(tolerance), float(init_var), int(max_iters), self.SQRTM_TOL,
                                            self.SQRTM_MAX_ITERS, _ptr(mu), _ptr(S),
                                            ctypes.cast(iters, ctypes.c_void_p), _ptr(info), _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_barycentre_fullcov")
        return mu, S, list(iters), info

    # ------------------------------------------------------------------ SURVEY 8f "next" row 4: the SVGP stage of GPDTW3D
    def svgp_fit(self, X, Y, Z0, batch_idx, n_steps, gamma=0.5, lr=0.01, train_hypers=True, jitter=DEFAULT_JITTER,
                 init_variance=1.0, init_lengthscale=1.0, predict_chunk=4096):
        """X [N, 4+R], Y [N, 2], Z0 [M, 4+R], batch_idx [2*n_steps, batch] int64 -> dict(mu [N], var [N] (+ Y[:,1]),
        Z, variances [4], lengthscales [4], q_mu [M], q_sqrt [M,M], info)   (ensembles/models.py:357-411)"""
        X = self._in(X)
        N, D = X.shape
        Y = self._in(Y, (N, 2), "Y")
        Z = self._in(Z0).clone()
        M = Z.shape[0]
        if Z.shape[1] != D:
            raise ValueError(f"Z0: expected {D} columns, got {Z.shape[1]}")
        idx = torch.as_tensor(batch_idx, dtype=torch.int64, device=self.device).contiguous()
        if n_steps > 0 and (idx.ndim != 2 or idx.shape[0] < 2 * n_steps):
            raise ValueError("batch_idx: need [2 * n_steps, minibatch_size] (one minibatch per half-step)")
        batch = int(idx.shape[1]) if idx.ndim == 2 else 1
        if n_steps > 0 and (int(idx.min()) < 0 or int(idx.max()) >= N):
            raise ValueError("batch_idx: index out of range")
        var = torch.full((4,), float(init_variance), dtype=torch.float64, device=self.device)
        ls = torch.full((4,), float(init_lengthscale), dtype=torch.float64, device=self.device)
        q_mu, q_sqrt = self._new(M), self._new(M, M)
        mu, v = self._new(N), self._new(N)
        info = self._new(1, dtype=torch.int32)
        chunk = int(min(predict_chunk, N))
        nbytes = int(self.lib.be_svgp_fit_workspace_bytes(N, D, M, batch, chunk))
        if nbytes == 0:
            raise ValueError(f"svgp_fit: unsupported shape (D = {D} must be 5 .. 36)")
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_svgp_fit(self.ctx, _ptr(X), _ptr(Y), N, D, M, batch, _ptr(idx), int(n_steps), float(gamma), float(lr),
                                  int(bool(train_hypers)), float(jitter), chunk, _ptr(Z), _ptr(var), _ptr(ls), _ptr(q_mu),
                                  _ptr(q_sqrt), _ptr(mu), _ptr(v), _ptr(info), _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_svgp_fit")
        return dict(mu=mu, var=v, Z=Z, variances=var, lengthscales=ls, q_mu=q_mu, q_sqrt=q_sqrt, info=info)

    # ------------------------------------------------------------------ SURVEY 8f "next" row 1: DTW barycentre averaging
    def dtw_barycenter_averaging_subgradient(self, reals, max_iter=30, initial_step_size=0.05, final_step_size=0.005,
                                             tol=1e-5, init_barycenter=None, want_info=False):
        """reals [B,R,T] -> barycentre [B,T] (tslearn semantics; models.py:176-178 passes max_iter=50, tol=1e-3)"""
        reals = self._in(reals)
        B, R, T = reals.shape
        init = None if init_barycenter is None else self._in(init_barycenter, (B, T), "init_barycenter")
        bary = self._new(B, T)
        n_iter = self._new(B, dtype=torch.int32)
        cost = self._new(B)
        nbytes = int(self.lib.be_dtw_dba_workspace_bytes(B, R, T))
        if nbytes == 0:
            raise ValueError(f"dtw_barycenter_averaging_subgradient: unsupported shape B={B} R={R} T={T} (T <= 4096)")
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_dtw_barycenter_averaging_subgradient(
            self.ctx, _ptr(reals), B, R, T, int(max_iter), float(initial_step_size), float(final_step_size),
            float(tol), _ptr(init), _ptr(bary), _ptr(n_iter), _ptr(cost), _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_dtw_barycenter_averaging_subgradient")
        return (bary, n_iter, cost) if want_info else bary

    def perform_dba(self, reals, n_iterations=10, want_medoid=False):
        """reals [B,R,T] -> centre [B,T] (ensembles/dtwa.py:6-20)"""
        reals = self._in(reals)
        B, R, T = reals.shape
        center = self._new(B, T)
        medoid = self._new(B, dtype=torch.int32)
        nbytes = int(self.lib.be_dtw_dba_workspace_bytes(B, R, T))
        if nbytes == 0:
            raise ValueError(f"perform_dba: unsupported shape B={B} R={R} T={T} (T <= 4096)")
        ws = self._ws(nbytes)
        self._sync_stream()
        rc = self.lib.be_perform_dba(self.ctx, _ptr(reals), B, R, T, int(n_iterations), _ptr(center), _ptr(medoid),
                                     _ptr(ws), nbytes)
        _lib.check(self.ctx, rc, "be_perform_dba")
        return (center, medoid) if want_medoid else center

    def dtw_squared(self, a, x):
        """a, x [P,T] -> squared DTW distance of each pair [P] (ensembles/dtwa.py:48-75)"""
        a = self._in(a)
        P, T = a.shape
        x = self._in(x, (P, T), "x")
        out = self._new(P)
        self._sync_stream()
        rc = self.lib.be_dtw_squared(self.ctx, _ptr(a), _ptr(x), P, T, _ptr(out))
        _lib.check(self.ctx, rc, "be_dtw_squared")
        return out

    # ------------------------------------------------------------------ SURVEY 8f "next": CRPS / similarity weights
    def crps_weights(self, loc, scale, obs, want_crps=False):
        """loc/scale [C,M,N], obs [C,Ro,N] -> weights [C,M,N] (+ crps_mean) (weights.py:444-515)"""
        loc = self._in(loc)
        C, M, N = loc.shape
        scale = self._in(scale, (C, M, N), "scale")
        obs = self._in(obs)
        Ro = obs.shape[1]
        w = self._new(C, M, N)
        cm = self._new(C, M, N) if want_crps else None
        self._sync_stream()
        rc = self.lib.be_crps_weights(self.ctx, _ptr(loc), _ptr(scale), _ptr(obs), C, M, Ro, N, _ptr(w), _ptr(cm))
        _lib.check(self.ctx, rc, "be_crps_weights")
        return (w, cm) if want_crps else w

    def ksd_weights(self, loc, scale, obs, want_ksd=False):
        """loc/scale [C,M,N], obs [C,Ro,N] -> weights [C,M,N] (+ ksd) (weights.py:336-441)"""
        loc = self._in(loc)
        C, M, N = loc.shape
        scale = self._in(scale, (C, M, N), "scale")
        obs = self._in(obs)
        Ro = obs.shape[1]
        w = self._new(C, M, N)
        k = self._new(C, M, N) if want_ksd else None
        self._sync_stream()
        rc = self.lib.be_ksd_weights(self.ctx, _ptr(loc), _ptr(scale), _ptr(obs), C, M, Ro, N, _ptr(w), _ptr(k))
        _lib.check(self.ctx, rc, "be_ksd_weights")
        return (w, k) if want_ksd else w

    def w2_collapse(self, w2):
        """w2 [C,M,M,N] -> weights [C,M,N]: nanmean over the second model, normalised over models"""
        w2 = self._in(w2)
        C, M, M2, N = w2.shape
        assert M == M2
        w = self._new(C, M, N)
        self._sync_stream()
        rc = self.lib.be_w2_collapse(self.ctx, _ptr(w2), C, M, N, _ptr(w))
        _lib.check(self.ctx, rc, "be_w2_collapse")
        return w

    def similarity_weights_pointwise(self, mean, var, want_w2=False):
        """mean/var [C,M,N] -> weights [C,M,N] (+ w2 [C,M,M,N]) (weights.py:302-325,331)"""
        mean = self._in(mean)
        C, M, N = mean.shape
        var = self._in(var, (C, M, N), "var")
        w = self._new(C, M, N)
        w2 = self._new(C, M, M, N) if want_w2 else None
        self._sync_stream()
        rc = self.lib.be_similarity_weights_pointwise(self.ctx, _ptr(mean), _ptr(var), C, M, N, _ptr(w), _ptr(w2))
        _lib.check(self.ctx, rc, "be_similarity_weights_pointwise")
        return (w, w2) if want_w2 else w
