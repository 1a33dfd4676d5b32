"""ensembles/wasserstein.py entry points, computed on the GPU through the C ABI."""
from __future__ import annotations

import warnings

import numpy as np

from .backend import Backend


def sqrtm(A):
    """wasserstein.py:10-13.  The reference takes ``U diag(sqrt(s)) V^H`` from an SVD; for the
    symmetric positive definite matrices it is applied to, that is the principal square root,
    which the device computes (scaled Denman-Beavers).  Raises for non-SPD input."""
    be = Backend.get()
    A = np.asarray(A, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("sqrtm expects a square matrix")
    out, _, _, info = be.sqrtm_psd(A[None])
    if int(info[0]) != 0:
        raise ValueError("sqrtm: the device path needs a symmetric positive definite matrix")
    return out[0].cpu().numpy()


def wasserstien_distance(A, B):
    """wasserstein.py:15-19: tr A + tr B - 2 tr sqrtm(sqrtm(A) B sqrtm(A))."""
    be = Backend.get()
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    z = np.zeros((1, A.shape[0]))
    w2, _ = be.w2_distance(z, A[None], z, B[None])
    return float(w2.item())


def gaussian_w2_distance_distrax(alpha, beta, full_cov=True):
    """wasserstein.py:21-47 (quirk Q-W2: the location term is the UN-squared 2-norm)."""
    be = Backend.get()
    mu1 = np.asarray(alpha.mean(), dtype=np.float64).reshape(1, -1)
    mu2 = np.asarray(beta.mean(), dtype=np.float64).reshape(1, -1)
    if full_cov:
        s1 = np.asarray(alpha.covariance(), dtype=np.float64)[None]
        s2 = np.asarray(beta.covariance(), dtype=np.float64)[None]
        w2, _ = be.w2_distance(mu1, s1, mu2, s2)
    else:
        v1 = np.asarray(alpha.variance(), dtype=np.float64).reshape(1, -1)
        v2 = np.asarray(beta.variance(), dtype=np.float64).reshape(1, -1)
        w2 = be.w2_distance_diag(mu1, v1, mu2, v2)
    return float(w2.item())


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


def gaussian_barycentre_fullcov(means, covariances, weights, tolerance: float = 1e-6, init_var=1.0):
    """BASELINE config 5: the matrix generalisation of ``gaussian_barycentre`` with ``sqrtm`` as
    the square root (no reference code; defined by the oracle).  ``means [M,T]``,
    ``covariances [M,T,T]``, ``weights [M]`` (or with a leading cell axis) -> ``(mu, S)``."""
    be = Backend.get()
    import torch

    cov = covariances if isinstance(covariances, torch.Tensor) else np.asarray(covariances, dtype=np.float64)
    single = cov.ndim == 3
    if single:
        cov = cov[None]
        means = (means if isinstance(means, torch.Tensor) else np.asarray(means, dtype=np.float64))[None]
        weights = (weights if isinstance(weights, torch.Tensor) else np.asarray(weights, dtype=np.float64))[None]
    mu, S, iters, _ = be.barycentre_fullcov(means, cov, weights, tolerance, init_var, 200)
    if any(i > 200 for i in iters):
        warnings.warn(f"Barycentre not converged for {sum(i > 200 for i in iters)} cell")
    mu, S = mu.cpu().numpy(), S.cpu().numpy()
    return (mu[0], S[0]) if single else (mu, S)
