"""``PerfectModelTest`` -- the integration-level caller of the hot path (ensembles/utils.py:32-228): each model
in turn is removed from the ensemble and used as pseudo-observations; the remaining members are fitted,
weighted and combined, and the barycentre is scored against the held-out model next to the multi-model mean.

Same constructor, ``_run_single_test`` and ``run`` signatures and the same arithmetic (utils.py:102-158, quirks
included: the weights are averaged over time before the barycentre, :111,133; the multi-model-mean "Normal" gets
the variance as its scale, :149, Q-SCALE; the RMSE averages over realisations inside the square root, :141,152).
Every numerical step is a call into the mirrored classes, i.e. into the C ABI.  The two matplotlib figures the
reference writes per pseudo-truth (:120-131, :161-183) are NOT produced (plotting is out of scope); the CSV is.
The module also carries the reference's checkpoint loader name (utils.py:22-30 ``load_model_collection``) from
``checkpoint.py``.
"""
from __future__ import annotations

import copy
import csv
import os

import numpy as np

from . import dists
from .checkpoint import load_model_collection, load_reference_pickle  # noqa: F401  (utils.py:22-30)
from .data import ModelCollection, ProcessModel
from .wasserstein import gaussian_w2_distance_distrax
from .weights import ModelSimilarityWeight


class PerfectModelTest:
    def __init__(self, hindcast_models: ModelCollection, forecast_models: ModelCollection, emulate_method,
                 weight_method, ensemble_method, ssp: str, include_sim: bool = False, save_dir: str = None):
        self.hindcast_models = hindcast_models
        self.forecast_models = forecast_models
        self.emulate_method = emulate_method
        self.weight_method = weight_method
        self.ensemble_method = ensemble_method
        self.ssp = ssp
        self.save_dir = save_dir
        self.include_sim = include_sim
        os.makedirs(save_dir, exist_ok=True)  # utils.py:67-68 (a None save_dir fails there too)
        self.save_csv_dir = os.path.join(save_dir, "csvs")
        os.makedirs(self.save_csv_dir, exist_ok=True)

    def _run_single_test(self, hindcast_models: ModelCollection, forecast_models: ModelCollection,
                         pseudo_observations_past: ProcessModel, pseudo_observations_future: ProcessModel,
                         n_optim_nits: int = 1000, use_prefit_models: bool = False):
        if use_prefit_models is not True:  # utils.py:104-108
            hindcast_models.fit(model=self.emulate_method(), compile_objective=True, n_optim_nits=n_optim_nits,
                                progress_bar=False)
            forecast_models.fit(model=self.emulate_method(), compile_objective=True, n_optim_nits=n_optim_nits,
                                progress_bar=False)
            dist = self.emulate_method().fit(pseudo_observations_future, compile_objective=True,
                                             n_optim_nits=n_optim_nits)
            pseudo_observations_future.distribution = dist
        weight_function = self.weight_method()
        weights = weight_function(hindcast_models, pseudo_observations_past)  # :110
        mean_weights = weights.mean("time")  # :112 (xarray skips NaN)
        if self.include_sim:  # :113-118
            sim_weights = ModelSimilarityWeight()(hindcast_models, pseudo_observations_future)
            mean_sim_weights = sim_weights.mean("time")
            total_weights = mean_weights * mean_sim_weights
            total_weights = total_weights / total_weights.sum()
        else:
            total_weights = mean_weights
        weights_single = total_weights.expand_dims(time=forecast_models[0].model_data.time, axis=1)  # :133
        barycentre = self.ensemble_method()(forecast_models, weights_single)  # :134-135
        obs = np.asarray(pseudo_observations_future.model_data.values, dtype=np.float64)
        # :139-146
        nll_bary = -float(np.mean(barycentre._dist.log_prob(obs)))
        bmean = np.asarray(barycentre.mean.values, dtype=np.float64)
        # utils.py:141: xarray orders the dims of (barycentre.mean[time] - model_data[realisation,time]) by first
        # appearance = (time, realisation), so its axis 0 is TIME: the mean inside the root runs over time
        rmse_bary = float(np.mean(np.sqrt(np.mean((bmean[None, :] - obs) ** 2, axis=1))))
        truth = pseudo_observations_future.distribution._dist
        w2_bary = gaussian_w2_distance_distrax(barycentre._dist, truth, full_cov=hasattr(truth, "covariance"))
        # :148-155: the multi-model mean
        realisations = np.vstack([np.asarray(forecast_models[i].model_data.values, dtype=np.float64)
                                  for i in range(forecast_models.number_of_models)])
        mmm_dist = dists.Normal(np.mean(realisations, axis=0), np.var(realisations, axis=0))
        nll_mmm = -float(np.mean(mmm_dist.log_prob(obs)))
        rmse_mmm = float(np.mean(np.sqrt(np.mean((mmm_dist.mean() - obs) ** 2, axis=0))))
        w2_mmm = gaussian_w2_distance_distrax(mmm_dist, truth, full_cov=False)
        self.last_weights = total_weights
        self.last_barycentre = barycentre
        return nll_bary, rmse_bary, w2_bary, nll_mmm, rmse_mmm, w2_mmm

    def run(self, n_optim_nits: int = 1000, use_prefit_models=False):
        name = self.weight_method().name
        columns = ["model as psuedo obs", f"nll_bary_{name}", f"rmse_bary_{name}", f"w2_bary_{name}", "nll_mmm",
                   "rmse_mmm", "w2_mmm"]
        rows = []
        for i in range(self.hindcast_models.number_of_models):  # utils.py:196-214
            hindcast_model_list = copy.deepcopy(self.hindcast_models.models)
            pseudo_observations_past = hindcast_model_list.pop(i)
            forecast_model_list = copy.deepcopy(self.forecast_models.models)
            pseudo_observations_future = forecast_model_list.pop(i)
            metrics = self._run_single_test(ModelCollection(hindcast_model_list), ModelCollection(forecast_model_list),
                                            pseudo_observations_past, pseudo_observations_future, n_optim_nits,
                                            use_prefit_models=use_prefit_models)
            rows.append([pseudo_observations_past.model_name, *metrics])
        if self.include_sim:
            file_name = f"prefect_model_test_results_{name}_plus_sim_{self.ssp}.csv"
        else:
            file_name = f"prefect_model_test_results_{name}_{self.ssp}.csv"
        save_file = os.path.join(self.save_csv_dir, file_name)
        with open(save_file, "w", newline="") as fh:  # df.to_csv layout: leading index column
            wr = csv.writer(fh)
            wr.writerow([""] + columns)
            for k, r in enumerate(rows):
                wr.writerow([k] + r)
        print(f"Saved results to {save_file}")
        self.results = dict(columns=columns, rows=rows)
        return self.results
