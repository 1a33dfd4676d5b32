"""Checkpoint I/O for ``ModelCollection`` (the checkpoint half).

The reference pickles the whole collection (ensembles/data.py:397-404, loader utils.py:22-30); the
pickle embeds xarray, pandas and distrax objects and so only loads where those exact libraries are
installed.  Here a collection is stored as a plain ``.npz`` (arrays + a JSON manifest): the member
data with their dims and coordinates, and per fitted member the distribution type with ``mu`` and
``covariance`` exactly as ``Distribution`` received them.  Loading rebuilds the distributions through
the same constructors, i.e. the Cholesky factor and log-prob statistics are recomputed on the GPU.

``load_reference_pickle`` additionally reads a pickle WRITTEN BY THE REFERENCE (e.g. its
``experiments/pre_fit_models/*.pkl``) without xarray/distrax, so that already fitted members can be
weighted and combined on the GPU.  Only the arrays are recovered; coordinates that live in pandas
indexes are replaced by ``arange``.
"""
from __future__ import annotations

import io
import json
import pickle

import numpy as np

from . import dists
from .data import Distribution, ModelCollection, ProcessModel
from .labelled import DataArray, ones_like

_DIST_TYPES = {c.__name__: c for c in (dists.MultivariateNormalFullCovariance, dists.MultivariateNormalDiag, dists.Normal)}


def save_model_collection(mc: ModelCollection, path: str) -> None:
    arrays, manifest = {}, {"format": "bayesian_ensembling_b200.ModelCollection/1", "models": []}
    for i, pm in enumerate(mc.models):
        da = pm.model_data
        entry = {"name": pm.model_name, "dims": list(da.dims), "array_name": da.name, "coords": {}}
        arrays[f"m{i}.data"] = np.asarray(da.values)
        for d, v in da.coords.items():
            arrays[f"m{i}.coord.{d}"] = np.asarray(v)
            entry["coords"][d] = f"m{i}.coord.{d}"
        dist = pm.distribution
        if dist is not None:
            entry["dist_type"] = dist.dist_type.__name__
            arrays[f"m{i}.mu"] = np.asarray(dist.mu, dtype=np.float64)
            arrays[f"m{i}.covariance"] = np.asarray(dist.covariance, dtype=np.float64)
        manifest["models"].append(entry)
    arrays["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    with open(path, "wb") as f:
        np.savez(f, **arrays)


def load_model_collection(path: str) -> ModelCollection:
    z = np.load(path, allow_pickle=False)
    manifest = json.loads(bytes(z["manifest"]).decode())
    assert manifest["format"].startswith("bayesian_ensembling_b200.ModelCollection/"), manifest["format"]
    models = []
    for i, e in enumerate(manifest["models"]):
        coords = {d: z[k] for d, k in e["coords"].items()}
        da = DataArray(z[f"m{i}.data"], tuple(e["dims"]), coords, name=e.get("array_name"))
        pm = ProcessModel(da, e["name"])
        if "dist_type" in e:
            pm.distribution = _rebuild(pm, z[f"m{i}.mu"], z[f"m{i}.covariance"], _DIST_TYPES[e["dist_type"]])
        models.append(pm)
    return ModelCollection(models)


def _rebuild(pm: ProcessModel, mu, covariance, dist_type) -> Distribution:
    blank = ones_like(pm.model_data[0].drop_vars("realisation")) * np.nan
    return Distribution(mu=mu, covariance=covariance, dim_array=blank.rename("blank"), dist_type=dist_type)


# ---- the reference's own pickles -------------------------------------------------------------
class _Stub:
    def __init__(self, *a, **k):
        self._args, self._kwargs, self._state = a, k, None

    def __setstate__(self, state):
        self._state = state

    def __call__(self, *a, **k):
        return _Stub(*a, **k)


class _StubUnpickler(pickle.Unpickler):
    """Every global that is not numpy / builtin becomes a stub that only remembers its state."""

    def find_class(self, module, name):
        if module.startswith("numpy"):
            return super().find_class(module.replace("numpy.core", "numpy._core"), name)
        if module in ("builtins", "collections", "copyreg", "datetime"):
            return super().find_class(module, name)
        if name[:1].isupper():
            return type(name, (_Stub,), {"__module__": module})
        return lambda *a, **k: _Stub(*a, **k)


def _state(stub):
    st = getattr(stub, "_state", None)
    if isinstance(st, tuple):
        for s in st:
            if isinstance(s, dict):
                return s
    return st


def _first_array(obj, ndim, depth=0):
    if isinstance(obj, np.ndarray):
        return obj if obj.dtype == np.float64 and obj.ndim == ndim else None
    if depth > 12:
        return None
    if isinstance(obj, dict):
        it = obj.values()
    elif isinstance(obj, (list, tuple)):
        it = obj
    elif isinstance(obj, _Stub):
        it = (getattr(obj, "_state", None), getattr(obj, "_args", None), getattr(obj, "_kwargs", None))
    else:
        return None
    for v in it:
        r = _first_array(v, ndim, depth + 1)
        if r is not None:
            return r
    return None


def load_reference_pickle(path: str) -> ModelCollection:
    """A ``ModelCollection`` pickled by the reference (data.py:397-404) -> this package's
    ``ModelCollection`` of 1-D members with their fitted full-covariance posteriors."""
    with open(path, "rb") as f:
        mc = _StubUnpickler(io.BytesIO(f.read())).load()
    models = []
    for pm in _state(mc)["models"]:
        st = _state(pm)
        reals = np.array(_first_array(st["model_data"], 2))
        R, T = reals.shape
        da = DataArray(reals, ("realisation", "time"), {"realisation": np.arange(R), "time": np.arange(T)})
        out = ProcessModel(da, str(st["model_name"]))
        if st.get("_distribution") is not None:
            d = _state(st["_distribution"])
            out.distribution = _rebuild(out, np.asarray(d["mu"], dtype=np.float64),
                                        np.asarray(d["covariance"], dtype=np.float64),
                                        dists.MultivariateNormalFullCovariance)
        models.append(out)
    return ModelCollection(models)
