"""Host-side containers with the reference's public names (ensembles/data.py: ``Distribution``, ``ProcessModel``,
``ModelCollection``), restricted to what the fit -> weight -> barycentre path touches.  Attribute names, argument
meaning and error behaviour follow the reference so that its call sites run unchanged; the bodies are this package's
own (forwarding properties are generated, iteration is one shared cursor, fitting is batched).  Plotting,
anomaly / climatology handling and pickling are out of scope.
"""
from __future__ import annotations

import typing as tp
import warnings
from dataclasses import dataclass

import numpy as np

from . import dists
from .labelled import DataArray, as_labelled


def _forwarded(call: tp.Callable[[tp.Any], tp.Any], doc: str) -> property:
    """A read-only property that forwards to the wrapped labelled array (or list of models)."""
    return property(call, doc=doc)


class _Cursor:
    """The reference's containers are their own iterators: ``__iter__`` returns the object and ``__next__`` walks an
    ``idx`` attribute that is rewound when the walk runs off the end (data.py:337-352, 369-383) -- so a ``break``
    leaves the cursor where it stopped.  Subclasses provide ``_at(i)`` raising ``IndexError`` past the end."""

    idx: int

    def __iter__(self):
        return self

    def __next__(self):
        try:
            item = self._at(self.idx)
        except IndexError:
            self.idx = 0
            raise StopIteration from None
        self.idx += 1
        return item


@dataclass
class Distribution:
    """ensembles/data.py:18-56,133-143.  ``dist_type(mu, covariance)`` is called POSITIONALLY (data.py:38-39), so what
    ``covariance`` means depends on ``dist_type`` (quirk Q-SCALE: a ``Normal`` takes it as the scale).  A batched fit
    hands over the distribution it already holds on the GPU through ``_prebuilt``."""

    mu: np.ndarray
    covariance: np.ndarray
    dim_array: DataArray
    dist_type: tp.Any
    _prebuilt: tp.Any = None

    def __post_init__(self):
        self._dist = self.dist_type(self.mu, self.covariance) if self._prebuilt is None else self._prebuilt

    def reshape(self, vals, name=False):
        """Flat values -> a labelled array with the coordinates of ``dim_array`` (data.py:41-56)."""
        labelled = self.dim_array.copy(data=np.asarray(vals).reshape(self.dim_array.shape))
        return labelled.rename(name) if name else labelled

    @property
    def mean(self):
        return self.reshape(self._dist.mean(), name="Distribution mean")

    @property
    def variance(self):
        return self.reshape(self._dist.variance(), name="Distribution variance")

    def sample(self):
        draw = self._dist.sample(seed=np.random.randint(0, 110000))  # the reference's seed range (data.py:141)
        return self.reshape(np.asarray(draw), name="Distribution sample")


@dataclass
class ProcessModel(_Cursor):
    """ensembles/data.py:146-352, data handling only: one climate model's realisations ``[realisation, time, ...]``."""

    model_data: DataArray
    model_name: str
    idx: int = 0
    _distribution = None

    def __post_init__(self):
        try:
            self.model_data = as_labelled(self.model_data)
        except TypeError:
            raise AssertionError("Input must be xr.DataArray")
        data = self.model_data
        assert data.dims[0] == "realisation"
        assert np.any(~np.isnan(data.values)), "Input data must not contain NaN"
        self.model_mean, self.model_std = data.mean(), data.std()
        self.climatology = None

    max_val = _forwarded(lambda self: self.model_data.max(), "largest value over all realisations")
    min_val = _forwarded(lambda self: self.model_data.min(), "smallest value over all realisations")
    time = _forwarded(lambda self: self.model_data.time, "the time coordinate")
    ndim = _forwarded(lambda self: self.model_data.ndim, "number of dimensions, the realisation axis included")
    n_realisations = _forwarded(lambda self: self.model_data.realisation.size, "number of realisations")
    mean_across_realisations = _forwarded(lambda self: self.model_data.mean("realisation"), "mean over realisations")
    std_across_realisations = _forwarded(lambda self: self.model_data.std("realisation"), "std over realisations")

    @property
    def distribution(self) -> Distribution:
        return self._distribution

    @distribution.setter
    def distribution(self, dist: Distribution):
        self._distribution = dist

    def __len__(self) -> int:
        return self.n_realisations

    def _at(self, i):
        return self.model_data.isel(realisation=i)


@dataclass
class ModelCollection(_Cursor):
    """ensembles/data.py:355-562: the ensemble's models, with ``fit`` and the time-axis check."""

    models: tp.List[ProcessModel]
    idx: int = 0

    def __post_init__(self):
        self.check_time_axes()

    def _at(self, i):
        return self.models[i]

    def __len__(self):
        return len(self.models)

    def __getitem__(self, item):
        return self.models[item]

    time = _forwarded(lambda self: self.models[0].time, "the first model's time coordinate")
    number_of_models = _forwarded(lambda self: len(self.models), "number of models")
    model_names = _forwarded(lambda self: [m.model_name for m in self.models], "the models' names, in order")
    max_val = _forwarded(lambda self: np.max([m.max_val.values for m in self.models]), "largest value of any model")
    min_val = _forwarded(lambda self: np.min([m.min_val.values for m in self.models]), "smallest value of any model")

    def distributions(self) -> tp.Dict[str, Distribution]:
        return {m.model_name: m.distribution for m in self.models}

    def fit(self, model, **kwargs):
        """ensembles/data.py:385-395.  The reference fits the members one after the other; a statistical model that
        offers ``fit_batch`` (GPDTW1D does) gets all members at once and fits those of equal shape in ONE batched
        device call -- same results, one launch sequence."""
        for m in self.models:
            if m.distribution != None:  # noqa: E711  (the reference's comparison)
                warnings.warn("Removing the model's previously learnt distribution")
        if hasattr(model, "fit_batch"):
            fitted = model.fit_batch(self.models, **kwargs)
        else:
            fitted = [model.fit(m, **kwargs) for m in self.models]
        for m, dist in zip(self.models, fitted):
            m.distribution = dist

    def save(self, path: str):
        """ensembles/data.py:397-404, as a library-independent ``.npz`` instead of a pickle of xarray / distrax
        objects; reload with ``utils.load_model_collection(path)``."""
        from .checkpoint import save_model_collection

        save_model_collection(self, path)

    def check_time_axes(self):
        """ensembles/data.py:542-562: if any two models' time coordinates differ, warn and give every model the first
        model's axis (the reference's "naive fix")."""
        axes = [m.model_data.time.values for m in self.models]
        first = axes[0]
        if all(t.shape == first.shape and not np.any(t != first) for t in axes[1:]):
            return
        warnings.warn(
            "Time axes of models don't match: applying naive fix. Check models are collocated correctly in time!")
        shared = self.time
        for m in self.models:
            m.model_data["time"] = shared


# the distrax names the reference passes as ``dist_type``
MultivariateNormalFullCovariance = dists.MultivariateNormalFullCovariance
MultivariateNormalDiag = dists.MultivariateNormalDiag
Normal = dists.Normal
