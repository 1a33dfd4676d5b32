"""``GPDTW1D`` -- the per-member GP posterior of the reference (ensembles/models.py:160-230),
same public signature, computed on the GPU through the C ABI.

Differences that are stated, not hidden:
* the DTW-barycentre-averaging mean (models.py:176-178: tslearn's
  ``dtw_barycenter_averaging_subgradient(realisation_set, max_iter=50, tol=1e-3)``) runs on the
  device too (``be_dtw_barycenter_averaging_subgradient``) and is the default,
  ``y_mean="dba"``; ``y_mean="mean"`` uses the arithmetic mean over realisations (tslearn's DBA
  initialiser) instead, and ``y_mean_fn`` supplies any other mean from the host;
* ``hyperparameters=(variance, lengthscale)`` selects the fixed-hyper-parameter posterior --
  the natural-gradient fixed point the reference's loop converges to -- without iterating;
  otherwise the natgrad(0.5)+Adam(0.01) loop of models.py:191-215 runs on the device for
  ``n_optim_nits`` iterations from GPflow's initial state.
"""
from __future__ import annotations

import numpy as np
import torch

from . import data as es_data
from . import dists
from .backend import Backend, DEFAULT_JITTER
from .labelled import ones_like


class GPDTW1D:
    def __init__(self, name: str = "GPRegressor", hyperparameters=None, y_mean_fn=None, y_mean: str = "dba") -> None:
        if y_mean not in ("dba", "mean"):
            raise ValueError(f"y_mean must be 'dba' or 'mean', got {y_mean!r}")
        self.name = name
        self.hyperparameters = hyperparameters
        self.y_mean_fn = y_mean_fn
        self.y_mean = y_mean

    # reference signature: models.py:164-170
    def fit(self, model, n_optim_nits: int = 500, compile_objective: bool = False, progress_bar: bool = True):
        return self.fit_batch([model], n_optim_nits=n_optim_nits, compile_objective=compile_objective,
                              progress_bar=progress_bar)[0]

    def fit_batch(self, models, n_optim_nits: int = 500, compile_objective: bool = False, progress_bar: bool = True):
        """Fits every ProcessModel in ``models`` (grouped by (R, T) shape) in batched device calls."""
        for m in models:
            if m.model_data.ndim > 2:
                raise NotImplementedError("Not implemented for more than temporal dimensions. Use GPDTW3D instead")
        be = Backend.get()
        out = [None] * len(models)
        groups = {}
        for i, m in enumerate(models):
            groups.setdefault(tuple(m.model_data.shape), []).append(i)
        for (R, T), idxs in groups.items():
            reals = np.stack([np.asarray(models[i].model_data.values, dtype=np.float64) for i in idxs])
            r_dev = be._in(reals)
            X, y_mean, y_var = be.gpdtw1d_inputs(r_dev)  # models.py:175-182
            if self.y_mean_fn is None and self.y_mean == "dba":  # models.py:176-178
                y_mean = be.dtw_barycenter_averaging_subgradient(r_dev, max_iter=50, tol=1e-3)
            if self.y_mean_fn is not None:
                y_mean = be._in(np.stack([np.asarray(self.y_mean_fn(reals[k])).ravel() for k in range(len(idxs))]))
            B = len(idxs)
            if self.hyperparameters is not None:
                var = torch.full((B,), float(self.hyperparameters[0]), dtype=torch.float64, device=be.device)
                ls = torch.full((B,), float(self.hyperparameters[1]), dtype=torch.float64, device=be.device)
                post = be.gp_posterior(X, y_mean, y_var, var, ls, DEFAULT_JITTER)
            else:
                post, _var, _ls = be.vgp_fit(X, y_mean, y_var, n_optim_nits)  # models.py:185-220
            for k, i in enumerate(idxs):
                pm = models[i]
                blank_array = ones_like(pm.model_data[0].drop_vars("realisation")) * np.nan
                blank_array = blank_array.rename("blank")
                dev = dists.MultivariateNormalFullCovariance(
                    _device_state=(post.mu[k], post.cov[k], post.scale_tri[k], post.var_diag[k], post.mvn_stats[k],
                                   post.info_dist[k]))
                out[i] = es_data.Distribution(
                    mu=post.mu[k].cpu().numpy(), covariance=post.cov[k].cpu().numpy(), dim_array=blank_array,
                    dist_type=dists.MultivariateNormalFullCovariance, _prebuilt=dev)
        return out


class MeanFieldApproximation:
    """ensembles/models.py:75-131.  The reference initialises ``mean`` / ``variance`` as the mean and the
    population variance over realisations (:104-105), runs ``n_optim_nits`` Adam steps on a copy of them
    (:116-121) and then returns the INITIAL values (:126-128: the loop's ``params`` are never read back), as
    ``Distribution(mu=mean, covariance=variance, dist_type=dx.Normal)`` -- the variance goes in as the scale
    (quirk Q-SCALE).  The dead loop is not run here; the moments come from the device kernel of
    models.py:175-182 (``be_gpdtw1d_inputs``) over the flattened trailing dimensions."""

    def __init__(self, name="MeanFieldModel"):
        self.name = name

    def fit(self, model, optimiser=None, n_optim_nits: int = 500, compile_objective: bool = False):
        if not optimiser:  # models.py:98-100
            import warnings

            warnings.warn("No optimiser specified, using Adam with learning rate 0.01")
        be = Backend.get()
        reals = np.asarray(model.model_data.values, dtype=np.float64).reshape(model.n_realisations, -1)  # :102-103
        _, mean, variance = be.gpdtw1d_inputs(be._in(reals[None]), want_X=False)
        blank_array = ones_like(model.model_data[0].drop_vars("realisation")) * np.nan
        blank_array = blank_array.rename("blank")
        return es_data.Distribution(mu=mean[0].cpu().numpy(), covariance=variance[0].cpu().numpy(),
                                    dim_array=blank_array, dist_type=dists.Normal)
