"""``Barycentre`` (ensembles/ensemble_scheme.py:21-81) on the GPU; same call signature."""
from __future__ import annotations

import abc
import typing as tp

import numpy as np

from . import dists
from .backend import Backend
from .data import Distribution, ModelCollection
from .labelled import DataArray, as_labelled, ones_like


class AbstractEnsembleScheme:
    def __init__(self, name: str) -> None:
        self.name = name
        self.distributions = None

    @abc.abstractmethod
    def _compute(self, process_models: ModelCollection, weights: DataArray) -> Distribution:
        raise NotImplementedError

    def __call__(self, process_models: ModelCollection, weights: DataArray, **kwargs) -> tp.Any:
        return self._compute(process_models=process_models, weights=weights, **kwargs)


class Barycentre(AbstractEnsembleScheme):
    def __init__(self, name: str = "Barycentre") -> None:
        super().__init__(name)

    def _compute(self, process_models: ModelCollection, weights: DataArray, n_threads=2) -> Distribution:
        """ensemble_scheme.py:43-81: one device launch over all points instead of the
        reference's Python double loop (:54,57); ``n_threads`` is accepted and unused, as there."""
        be = Backend.get()
        models = list(process_models.models)
        M = process_models.number_of_models
        n_points = int(models[0].model_data.size / models[0].model_data.realisation.size)  # :51
        w = np.asarray(as_labelled(weights).values, dtype=np.float64).reshape(M, -1)  # :52
        means, variances = [], []
        for t_idx, process_model in enumerate(models):
            if not process_model.distribution:  # :58-61
                raise AttributeError(f"No posterior for model {t_idx}. Please run model.fit() first.")
            dist = process_model.distribution._dist
            means.append(_dev_or_np(be, dist, "mean"))
            variances.append(_dev_or_np(be, dist, "variance"))
        import torch

        mean_t = torch.stack(means)[None]
        var_t = torch.stack(variances)[None]
        mu, sigma, iters = be.barycentre_1d(mean_t, var_t, be._in(w[None]))  # :63-69
        mu, sigma, iters = mu[0].cpu().numpy(), sigma[0].cpu().numpy(), iters[0].cpu().numpy()
        if (iters > 200).any():  # wasserstein.py:94-97
            import warnings

            warnings.warn(f"Barycentre not converged for {int((iters > 200).sum())} time step")
        blank_array = ones_like(models[-1].model_data[0].drop_vars("realisation")) * np.nan  # :73-74
        blank_array = blank_array.rename("blank")
        assert mu.shape[0] == n_points
        return Distribution(mu=mu, covariance=sigma ** 2, dim_array=blank_array,
                            dist_type=dists.MultivariateNormalDiag)  # :75-78 (quirk Q-SCALE)


def _dev_or_np(be, dist, what):
    """mean / variance vector of a member on the device (no host round trip for GPU-resident fits)."""
    if isinstance(dist, dists.MultivariateNormalFullCovariance):
        return dist._loc if what == "mean" else dist._var_diag
    return be._in(np.asarray(getattr(dist, what)()).ravel())
