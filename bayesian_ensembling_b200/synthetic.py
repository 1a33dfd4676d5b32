"""Synthetic CMIP6 / HadCRUT5-shaped inputs for the five BASELINE.json configs.

Host NumPy, seeded (``numpy.random.default_rng(20240 + config index)``); the same arrays feed
the CUDA path and the CPU oracle.  Per (cell c, member m): a GMST-like trend
``a_m t + b_m t^2`` on ``t in [0,1]``, a cell offset, a small seasonal term for monthly
configs and independent AR(1) noise (phi=0.6, sigma=0.12) per realisation, so that the
across-realisation variance is ~0.02 as in the reference's fitted pickles.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# hyper-parameters used for fixed-theta (L1) runs: centre of the range recovered from the
# reference's pickled fits (tests/golden/make_golden_from_reference.py)
L1_VARIANCE = 0.5
L1_LENGTHSCALE = 6.0


@dataclass(frozen=True)
class Config:
    name: str
    index: int
    cells: int
    members: int
    realisations: int
    steps: int
    obs_realisations: int
    monthly: bool
    description: str


CONFIGS = {
    "cfg1": Config("cfg1", 1, 1, 10, 3, 251, 10, False, "10 models x 3 realisations x 251 annual steps, single location"),
    "cfg2": Config("cfg2", 2, 1, 24, 5, 3012, 10, True, "24 models x 5 realisations x 3012 months, single location"),
    "cfg3": Config("cfg3", 3, 2592, 24, 5, 1980, 10, True, "36x72 cells x 24 models x 5 realisations x 1980 months"),
    "cfg4": Config("cfg4", 4, 64800, 40, 10, 251, 10, False, "180x360 cells x 40 models x 10 realisations x 251 years"),
}


def _ar1(rng, shape, phi=0.6, sigma=0.12):
    """AR(1) along the last axis with stationary std ``sigma``."""
    eps = rng.standard_normal(shape) * (sigma * np.sqrt(1.0 - phi * phi))
    out = np.empty(shape)
    out[..., 0] = rng.standard_normal(shape[:-1]) * sigma
    for t in range(1, shape[-1]):
        out[..., t] = phi * out[..., t - 1] + eps[..., t]
    return out


def make_cells(cfg: Config, n_cells: int | None = None, cell_offset: int = 0, seed: int | None = None):
    """Returns ``(realisations [C,M,R,T], observations [C,Ro,T])`` for cells
    ``cell_offset .. cell_offset + n_cells`` of ``cfg`` (each cell has its own stream, so any
    subset / shard reproduces the same numbers)."""
    C = cfg.cells if n_cells is None else n_cells
    M, R, T, Ro = cfg.members, cfg.realisations, cfg.steps, cfg.obs_realisations
    base = 20240 + cfg.index if seed is None else seed
    tn = np.linspace(0.0, 1.0, T)
    season = 0.3 * np.sin(2.0 * np.pi * (np.arange(T) % 12) / 12.0) if cfg.monthly else np.zeros(T)
    # member trend coefficients are shared by all cells (one climate model = one sensitivity)
    rng_m = np.random.default_rng([base, 0])
    a = rng_m.uniform(0.5, 4.0, size=M + 1)
    b = rng_m.uniform(0.0, 2.0, size=M + 1)
    reals = np.empty((C, M, R, T))
    obs = np.empty((C, Ro, T))
    for ci in range(C):
        rng = np.random.default_rng([base, 1, cell_offset + ci])
        off = rng.normal(0.0, 0.5)
        trend = a[:M, None] * tn[None, :] + b[:M, None] * tn[None, :] ** 2
        reals[ci] = (off + trend + season[None, :])[:, None, :] + _ar1(rng, (M, R, T))
        otrend = a[M] * tn + b[M] * tn**2
        obs[ci] = (off + otrend + season)[None, :] + _ar1(rng, (Ro, T))
    return reals, obs
