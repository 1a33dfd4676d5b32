"""Host-side mirror of the reference's data containers (ensembles/data.py), restricted to what
the fit -> weight -> barycentre path uses.  Same names, argument meaning and error behaviour;
plotting, anomaly/climatology and pickling are out of scope.
"""
from __future__ import annotations

import typing as tp
import warnings
from dataclasses import dataclass

import numpy as np

from . import dists
from .labelled import DataArray, as_labelled


@dataclass
class Distribution:
    """ensembles/data.py:18-56,133-143.  ``dist_type(mu, covariance)`` is called POSITIONALLY
    (data.py:38-39), so the meaning of ``covariance`` depends on ``dist_type`` (quirk Q-SCALE)."""

    mu: np.ndarray
    covariance: np.ndarray
    dim_array: DataArray
    dist_type: tp.Any
    _prebuilt: tp.Any = None  # a distribution already resident on the GPU (batched fits)

    def __post_init__(self):
        self._dist = self._prebuilt if self._prebuilt is not None else self.dist_type(self.mu, self.covariance)

    def reshape(self, vals, name=False):
        reshaped_vals = np.asarray(vals).reshape(self.dim_array.shape)
        reshaped_array = self.dim_array.copy(data=reshaped_vals)
        if name:
            reshaped_array = reshaped_array.rename(name)
        return reshaped_array

    @property
    def mean(self):
        return self.reshape(self._dist.mean(), name="Distribution mean")

    @property
    def variance(self):
        return self.reshape(self._dist.variance(), name="Distribution variance")

    def sample(self):
        samples = np.asarray(self._dist.sample(seed=np.random.randint(0, 110000)))
        return self.reshape(samples, name="Distribution sample")


@dataclass
class ProcessModel:
    """ensembles/data.py:146-352 (data handling only)."""

    model_data: DataArray
    model_name: str
    idx: int = 0
    _distribution = None

    def __post_init__(self):
        try:
            self.model_data = as_labelled(self.model_data)
        except TypeError:
            raise AssertionError("Input must be xr.DataArray")
        self.model_mean = self.model_data.mean()
        self.model_std = self.model_data.std()
        self.climatology = None
        assert self.model_data.dims[0] == "realisation"
        assert np.any(~np.isnan(self.model_data.values)), "Input data must not contain NaN"

    @property
    def max_val(self):
        return self.model_data.max()

    @property
    def min_val(self):
        return self.model_data.min()

    @property
    def n_realisations(self) -> int:
        return self.model_data.realisation.size

    @property
    def time(self):
        return self.model_data.time

    @property
    def mean_across_realisations(self):
        return self.model_data.mean("realisation")

    @property
    def std_across_realisations(self):
        return self.model_data.std("realisation")

    @property
    def ndim(self):
        return self.model_data.ndim

    @property
    def distribution(self) -> Distribution:
        return self._distribution

    @distribution.setter
    def distribution(self, dist: Distribution):
        self._distribution = dist

    def __len__(self) -> int:
        return self.n_realisations

    def __iter__(self):
        return self

    def __next__(self):
        try:
            out = self.model_data.isel(realisation=self.idx)
            self.idx += 1
        except IndexError:
            self.idx = 0
            raise StopIteration
        return out


@dataclass
class ModelCollection:
    """ensembles/data.py:355-562 (data handling + ``fit``)."""

    models: tp.List[ProcessModel]
    idx: int = 0

    def __post_init__(self):
        self.check_time_axes()

    def __iter__(self):
        return self

    def __next__(self):
        try:
            out = self.models[self.idx]
            self.idx += 1
        except IndexError:
            self.idx = 0
            raise StopIteration
        return out

    def fit(self, model, **kwargs):
        """ensembles/data.py:385-395.  The reference loops members serially; when the
        statistical model offers ``fit_batch`` (GPDTW1D does) all members of equal shape are
        fitted in ONE batched device call -- same results, one launch sequence."""
        for process_model in self.models:
            if process_model.distribution != None:  # noqa: E711  (as the reference)
                warnings.warn("Removing the model's previously learnt distribution")
        if hasattr(model, "fit_batch"):
            dists_ = model.fit_batch(self.models, **kwargs)
            for process_model, dist in zip(self.models, dists_):
                process_model.distribution = dist
            return
        for process_model in self.models:
            dist = model.fit(process_model, **kwargs)
            process_model.distribution = dist

    def save(self, path: str):
        """ensembles/data.py:397-404, as a library-independent ``.npz`` instead of a pickle of
        xarray / distrax objects; reload with ``utils.load_model_collection(path)``."""
        from .checkpoint import save_model_collection

        save_model_collection(self, path)

    @property
    def time(self):
        return self.models[0].time

    @property
    def max_val(self):
        return np.max([model.max_val.values for model in self.models])

    @property
    def min_val(self):
        return np.min([model.min_val.values for model in self.models])

    @property
    def number_of_models(self):
        return len(self.models)

    @property
    def model_names(self):
        return [model.model_name for model in self.models]

    def __len__(self):
        return len(self.models)

    def __getitem__(self, item):
        return self.models[item]

    def distributions(self) -> tp.Dict[str, Distribution]:
        return {model.model_name: model.distribution for model in self.models}

    def check_time_axes(self):
        """ensembles/data.py:542-562."""
        time_axes_match = True
        for model1 in self.models:
            for model2 in self.models:
                t1, t2 = model1.model_data.time.values, model2.model_data.time.values
                if t1.shape != t2.shape or np.any(t1 != t2):
                    time_axes_match = False
        if time_axes_match == False:  # noqa: E712
            warnings.warn(
                "Time axes of models don't match: applying naive fix. Check models are collocated correctly in time!"
            )
            new_time = self.time
            for model in self:
                model.model_data["time"] = new_time
        return


# the distrax names the reference passes as ``dist_type``
MultivariateNormalFullCovariance = dists.MultivariateNormalFullCovariance
MultivariateNormalDiag = dists.MultivariateNormalDiag
Normal = dists.Normal
