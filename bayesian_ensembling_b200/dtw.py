"""DTW barycentre averaging on the GPU, behind the two interfaces the reference touches:

* ``dtw_barycenter_averaging_subgradient`` -- tslearn 0.5.1.0's function of that name, which the
  reference imports at ensembles/models.py:15 and calls at :176-178 / :251-253
  (``max_iter=50, tol=1e-3``) to obtain ``y_mean``;
* ``performDBA`` -- the reference's own NumPy implementation, ensembles/dtwa.py:6-20
  (exported at ensembles/__init__.py:3).

Both run through the C ABI (``be_dtw_barycenter_averaging_subgradient`` / ``be_perform_dba``);
the batched forms on ``Backend`` take ``[B, R, T]`` and return device tensors, which is how
``GPDTW1D`` and ``grid`` use them (one (cell, member) problem per batch entry).  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from .backend import Backend


def dtw_barycenter_averaging_subgradient(X, barycenter_size=None, init_barycenter=None, max_iter=30,
                                         initial_step_size=0.05, final_step_size=0.005, tol=1e-5,
                                         random_state=None, weights=None, metric_params=None, verbose=False):
    """tslearn signature; ``X`` is ``[n_series, T]`` or ``[n_series, T, 1]``; returns ``[T, 1]``.

    ``barycenter_size`` other than T, ``weights``, ``metric_params`` and d > 1 are not on the
    reference's path (it passes none of them) and raise ``NotImplementedError``.
    ``random_state`` is accepted and unused: the batch sub-gradient iteration has no random step.
    """
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 3:
        if X.shape[2] != 1:
            raise NotImplementedError("multivariate series are not on the reference's path (d == 1 only)")
        X = X[:, :, 0]
    if X.ndim != 2:
        raise ValueError(f"X must be [n_series, T] or [n_series, T, 1], got {X.shape}")
    T = X.shape[1]
    if barycenter_size is not None and barycenter_size != T:
        raise NotImplementedError("barycenter_size != T is not on the reference's path")
    if weights is not None or metric_params:
        raise NotImplementedError("weights / metric_params are not on the reference's path")
    init = None
    if init_barycenter is not None:
        init = np.asarray(init_barycenter, dtype=np.float64).reshape(1, -1)
        if init.shape[1] != T:
            raise NotImplementedError("init_barycenter of a different length is not on the reference's path")
    be = Backend.get()
    bary = be.dtw_barycenter_averaging_subgradient(X[None], max_iter=max_iter, initial_step_size=initial_step_size,
                                                   final_step_size=final_step_size, tol=tol, init_barycenter=init)
    return bary[0].cpu().numpy().reshape(-1, 1)


def performDBA(series, n_iterations=10):
    """ensembles/dtwa.py:6-20 for equal-length series (<= 50 of them: beyond that the reference
    samples its medoid candidates from an unseeded RNG, dtwa.py:26)."""
    X = np.asarray(series, dtype=np.float64)
    if X.ndim != 2:
        raise NotImplementedError("performDBA: series of unequal length are not supported on the device path")
    if X.shape[0] > 50:
        raise NotImplementedError("performDBA: more than 50 series (unseeded candidate sampling in the reference)")
    be = Backend.get()
    return be.perform_dba(X[None], n_iterations=n_iterations)[0].cpu().numpy()
