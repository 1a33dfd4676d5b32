"""``LogLikelihoodWeight`` (ensembles/weights.py:15-131) on the GPU; same call signature."""
from __future__ import annotations

import abc
import copy

import numpy as np
import torch

from . import dists
from .backend import Backend
from .data import ModelCollection, ProcessModel
from .labelled import DataArray, concat


class AbstractWeight:
    """ensembles/weights.py:15-53."""

    def __init__(self, name: str) -> None:
        self.name = name

    @abc.abstractmethod
    def _compute(self, process_models: ModelCollection, observations: ProcessModel) -> DataArray:
        raise NotImplementedError

    def __call__(self, process_models: ModelCollection, observations: ProcessModel = None, **kwargs) -> DataArray:
        if observations is not None:  # weights.py:44-46
            assert np.all(process_models.time.values == observations.time.values), \
                "Time coordinates do not match between models and observations"
            assert len(process_models.time) == len(observations.time), \
                "Time coordinates do not match between models and observations"
        for process_model in process_models.models:  # weights.py:48-49
            assert hasattr(process_model.distribution, "_dist"), "Distribution not defined - fit models first"
        return self._compute(process_models=process_models, observations=observations, **kwargs)


class LogLikelihoodWeight(AbstractWeight):
    """ensembles/weights.py:56-131.  ``standardisation_scheme`` must be ``exp`` (the reference's
    default ``jnp.exp``): the exponential is fused into the device kernel; pass ``np.exp`` or
    leave the default."""

    def __init__(self, name: str = "LogLikelihoodWeight") -> None:
        super().__init__(name)

    def _compute(self, process_models, observations, return_lls=False, standardisation_scheme=np.exp,
                 standardisation_constant=1.0):
        if standardisation_scheme not in (np.exp,) and getattr(standardisation_scheme, "__name__", "") != "exp":
            raise NotImplementedError("only exp is supported as standardisation_scheme on the device path")
        be = Backend.get()
        models = list(process_models.models)
        M = len(models)
        obs = np.stack([np.asarray(o.values, dtype=np.float64).ravel() for o in _realisations(observations)])  # [Ro,N]
        obs_shape = observations.model_data.shape[1:]
        obs_dev = be._in(obs[None])  # [C=1, Ro, N]
        is_normal = [m.distribution.dist_type == dists.Normal for m in models]
        if all(is_normal):  # weights.py:95-96
            loc = be._in(np.stack([m.distribution._dist.mean().ravel() for m in models])[None])
            scale = be._in(np.stack([m.distribution._dist.stddev().ravel() for m in models])[None])
            w, le, _ = be.loglik_weights_normal(loc, scale, obs_dev, standardisation_constant, want_lls=True)
        elif not any(is_normal):  # weights.py:97-100 (quirk Q-LL)
            stats = torch.stack([m.distribution._dist._stats for m in models])  # [M,4] on device
            w, le, _ = be.loglik_weights_mvn(stats, obs_dev, M, standardisation_constant, want_lls=True)
        else:
            raise NotImplementedError("mixed Normal / multivariate members")
        w = w[0].cpu().numpy()
        le = le[0].cpu().numpy()

        def wrap(vals, name):
            per_model = []
            for m, v in zip(models, vals):  # weights.py:110-115
                x = copy.deepcopy(m.model_data.isel(realisation=0)).drop_vars("realisation")
                x.data = v.reshape(x.shape)
                per_model.append(x.assign_coords(model=m.model_name))
            return concat(per_model, dim="model").rename(name)

        weights = wrap(w, "Log-likelihood weights")
        assert weights.shape == (len(process_models),) + tuple(obs_shape)  # weights.py:126
        if return_lls:
            return weights, wrap(le, "Log-likelihoods")
        return weights


class UniformWeight(AbstractWeight):
    """ensembles/weights.py:187-212: 1/M everywhere (no device work needed)."""

    def __init__(self, name: str = "UniformWeight") -> None:
        super().__init__(name)

    def _compute(self, process_models, observations=None):
        per_model = []
        M = len(process_models)
        for m in process_models.models:
            x = copy.deepcopy(m.model_data.isel(realisation=0)).drop_vars("realisation")
            x.data = np.full(x.shape, 1.0 / M)
            per_model.append(x.assign_coords(model=m.model_name))
        return concat(per_model, dim="model").rename("Uniform weights")


def _realisations(pm: ProcessModel):
    return [pm.model_data.isel(realisation=i) for i in range(pm.n_realisations)]
