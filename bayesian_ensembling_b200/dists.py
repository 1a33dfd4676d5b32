"""distrax-free stand-ins for the three distributions the reference builds through
``Distribution.__post_init__`` (ensembles/data.py:38-39): ``dist_type(mu, covariance)`` with
POSITIONAL arguments, so for ``Normal`` and ``MultivariateNormalDiag`` the second argument is
a *scale* (quirk Q-SCALE) while ``MultivariateNormalFullCovariance`` gets a true covariance.

State lives on the GPU; every number is produced by the C ABI (Cholesky, log-prob
statistics, log-densities).  ``mean() / variance() / covariance() / log_prob()`` return NumPy
arrays like the reference's ``jnp`` arrays do after ``np.asarray``.
"""
from __future__ import annotations

import numpy as np
import torch

from .backend import Backend


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


class MultivariateNormalFullCovariance:
    """``dx.MultivariateNormalFullCovariance(loc, covariance_matrix)``: the distrax constructor
    takes ``jnp.linalg.cholesky(covariance_matrix)`` as ``scale_tri`` (used at
    models.py:224-229)."""

    def __init__(self, loc=None, covariance_matrix=None, *, _device_state=None):
        be = Backend.get()
        if _device_state is not None:
            # already on device: a slice of a batched posterior (no recomputation)
            self._loc, self._cov, self._scale_tri, self._var_diag, self._stats, self._info = _device_state
        else:
            loc_t = be._in(np.asarray(loc, dtype=np.float64).reshape(1, -1))
            cov_t = be._in(np.asarray(covariance_matrix, dtype=np.float64)[None])
            tri, var_diag, stats, info = be.mvn_from_cov(loc_t, cov_t)
            self._loc, self._cov, self._scale_tri = loc_t[0], cov_t[0], tri[0]
            self._var_diag, self._stats, self._info = var_diag[0], stats[0], info[0]
        self.event_shape = (int(self._loc.shape[-1]),)

    # --- distrax surface used by the reference ------------------------------------------
    def mean(self):
        return _np(self._loc)

    def covariance(self):
        return _np(self._cov)

    def variance(self):
        return _np(self._var_diag)

    def stddev(self):
        return np.sqrt(self.variance())

    @property
    def scale_tri(self):
        return _np(self._scale_tri)

    @property
    def loc(self):
        return self.mean()

    def log_prob(self, value):
        """Event = last axis; ``value`` broadcasts against ``loc`` (distrax semantics).
        The reference's only full-covariance call passes ``obs.ravel()[:, None]``
        (weights.py:98-100): T constant vectors -> the constant-vector fast path (quirk Q-LL)."""
        value = np.asarray(value, dtype=np.float64)
        T = self.event_shape[0]
        if value.ndim >= 1 and value.shape[-1] == 1 and T > 1:
            be = Backend.get()
            flat = value.reshape(-1)  # each entry o -> log N(o * 1_T)
            # kernel layout obs [C=1, Ro=1, N]: treat the N constants as the "time" axis
            stats = self._stats.reshape(1, 4)
            n = flat.shape[0]
            st = stats.clone()
            ll = _constvec_ll(be, st, flat, T)
            return ll.reshape(value.shape[:-1])
        if value.shape[-1] != T:
            raise ValueError(f"log_prob: last axis must be {T} or 1, got {value.shape}")
        return self._log_prob_general(value)

    def _log_prob_general(self, value):
        # General vectors: ONE forward substitution per vector against the factor stored at construction
        # (be_mvn_log_prob), as distrax does -- no re-factorisation.
        be = Backend.get()
        T = self.event_shape[0]
        flat = np.ascontiguousarray(value.reshape(-1, T))
        ll = be.mvn_log_prob(self._loc, self._scale_tri, flat, float(self._stats[3]))
        return _np(ll).reshape(value.shape[:-1])

    def sample(self, seed=0, sample_shape=()):
        rng = np.random.default_rng(int(np.asarray(seed).ravel()[0]) if np.ndim(seed) else int(seed))
        shape = tuple(np.atleast_1d(sample_shape).astype(int)) if sample_shape != () else ()
        z = rng.standard_normal(shape + self.event_shape)
        return self.mean() + z @ self.scale_tri.T


def _constvec_ll(be: Backend, stats, constants: np.ndarray, T_event: int) -> np.ndarray:
    """log N(o * 1 | mu, Sigma) for every o in ``constants`` via be_mvn_constvec_logprob.
    The kernel's normaliser uses its T argument, so the event dimension is passed as T and the
    constants are chunked along the kernel's time axis."""
    n = constants.shape[0]
    out = np.empty(n)
    # kernel signature obs [C, Ro, T]: feed chunks of T_event constants (pad the tail)
    for s0 in range(0, n, T_event):
        chunk = constants[s0:s0 + T_event]
        buf = np.zeros(T_event)
        buf[: chunk.shape[0]] = chunk
        ll = be.mvn_constvec_logprob(stats, be._in(buf.reshape(1, 1, T_event)), 1)
        out[s0:s0 + chunk.shape[0]] = _np(ll).reshape(-1)[: chunk.shape[0]]
    return out


class MultivariateNormalDiag:
    """``dx.MultivariateNormalDiag(loc, scale_diag)``.  ``Barycentre`` passes a VARIANCE as
    ``scale_diag`` (ensemble_scheme.py:75-78), so ``variance()`` is that value squared."""

    def __init__(self, loc=None, scale_diag=None):
        self._loc = np.asarray(loc, dtype=np.float64)
        self._scale = np.asarray(scale_diag, dtype=np.float64)
        self.event_shape = (int(self._loc.shape[-1]),)

    def mean(self):
        return self._loc

    def stddev(self):
        return self._scale

    def variance(self):
        return self._scale * self._scale

    def covariance(self):
        return np.diag(self.variance())

    def log_prob(self, value):
        be = Backend.get()
        value = np.asarray(value, dtype=np.float64)
        x = np.broadcast_to(value, np.broadcast_shapes(value.shape, self._loc.shape))
        ll = be.normal_logprob(np.broadcast_to(self._loc, x.shape).copy(), np.broadcast_to(self._scale, x.shape).copy(),
                               np.ascontiguousarray(x))
        return _np(ll.sum(-1))

    def sample(self, seed=0, sample_shape=()):
        rng = np.random.default_rng(int(seed))
        return self._loc + self._scale * rng.standard_normal(self._loc.shape)


class Normal:
    """``dx.Normal(loc, scale)``; the reference passes a variance as ``scale``
    (models.py:129-131), kept as is."""

    def __init__(self, loc=None, scale=None):
        self._loc = np.asarray(loc, dtype=np.float64)
        self._scale = np.asarray(scale, dtype=np.float64)

    def mean(self):
        return self._loc

    def stddev(self):
        return self._scale

    def variance(self):
        return self._scale * self._scale

    def log_prob(self, value):
        be = Backend.get()
        value = np.asarray(value, dtype=np.float64)
        x = np.broadcast_to(value, np.broadcast_shapes(value.shape, self._loc.shape))
        ll = be.normal_logprob(np.broadcast_to(self._loc, x.shape).copy(), np.broadcast_to(self._scale, x.shape).copy(),
                               np.ascontiguousarray(x))
        return _np(ll)

    def sample(self, seed=0, sample_shape=()):
        rng = np.random.default_rng(int(seed))
        return self._loc + self._scale * rng.standard_normal(self._loc.shape)
