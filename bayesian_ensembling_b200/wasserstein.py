"""ensembles/wasserstein.py entry points that sit on the hot path."""
from __future__ import annotations

import warnings

import numpy as np

from .backend import Backend


def gaussian_barycentre(means, std_devs, weights, tolerance: float = 1e-6, init_var=1.0):
    """wasserstein.py:61-100 (signed stop rule, quirk Q-BARY) for ONE point; returns (mu, sigma)."""
    be = Backend.get()
    means = np.asarray(means, dtype=np.float64).reshape(1, -1, 1)
    var = (np.asarray(std_devs, dtype=np.float64) ** 2).reshape(1, -1, 1)
    w = np.asarray(weights, dtype=np.float64).reshape(1, -1, 1)
    mu, sigma, iters = be.barycentre_1d(means, var, w, tolerance, init_var, 200)
    if int(iters.item()) > 200:
        warnings.warn("Barycentre not converged for 1 time step")
    return float(mu.item()), float(sigma.item())
