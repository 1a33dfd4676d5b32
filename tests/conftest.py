"""pytest configuration: the ``gpu`` marker, import path, shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prefit_members.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _unpack_lower(packed, T):
    out = np.zeros((T, T))
    out[np.tril_indices(T)] = packed
    return out


class GoldenMember:
    """One fitted member of the reference's pickles (tests/golden/make_golden_from_reference.py)."""

    def __init__(self, z, key, name):
        self.key, self.name = key, name
        self.realisations = z[f"{key}.realisations"]
        self.mu = z[f"{key}.mu"]
        T = self.mu.shape[0]
        low = _unpack_lower(z[f"{key}.cov_lower"], T)
        self.cov = low + np.tril(low, -1).T
        self.cov_asym = float(z[f"{key}.cov_asym"])
        self.scale_tri = _unpack_lower(z[f"{key}.scale_tri_lower"], T)
        self.variance, self.lengthscale, self.hyper_resid = z[f"{key}.recovered_hypers"]
        self.tag = key.split(".")[0]


@pytest.fixture(scope="session")
def golden_members():
    z = np.load(GOLDEN)
    return [GoldenMember(z, n.split("|")[0], n.split("|")[1]) for n in z["names"]]


@pytest.fixture(scope="session")
def backend():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bayesian_ensembling_b200.backend import Backend

    return Backend.get()


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-300))
