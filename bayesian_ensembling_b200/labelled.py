"""A small labelled-array stand-in for the subset of ``xarray.DataArray`` the hot path touches.

The reference keeps every field in an ``xr.DataArray`` with dims ``(realisation, time[, lat,
lon])`` (ensembles/data.py:158-170).  xarray is not available in this image, and the hot path
only needs: ``.values .dims .shape .ndim .size``, coordinate access by attribute
(``da.time``, ``da.realisation.size``), ``isel``, reductions over a named dim, ``copy(data=)``,
``rename``, ``drop_vars``, ``expand_dims``, ``assign_coords``, ``sel(model=...)``, arithmetic and
``concat``.  Real ``xarray.DataArray`` objects are accepted wherever this class is (see
``as_labelled``), so a user of the reference can pass their own arrays.
"""
from __future__ import annotations

import copy as _copy

import numpy as np


class DataArray:
    __array_priority__ = 50

    def __init__(self, data, dims, coords=None, name=None):
        self.values = np.asarray(data)
        self.dims = tuple(dims)
        if len(self.dims) != self.values.ndim:
            raise ValueError(f"dims {self.dims} do not match data of shape {self.values.shape}")
        self.coords = {}
        for k, v in (coords or {}).items():
            self.coords[k] = v if np.ndim(v) == 0 else np.asarray(v)
        for d, n in zip(self.dims, self.values.shape):
            if d not in self.coords:
                self.coords[d] = np.arange(n)
        self.name = name

    # ---- basic properties -------------------------------------------------------------
    @property
    def shape(self):
        return self.values.shape

    @property
    def ndim(self):
        return self.values.ndim

    @property
    def size(self):
        return self.values.size

    @property
    def data(self):
        return self.values

    @data.setter
    def data(self, v):
        v = np.asarray(v)
        if v.shape != self.values.shape:
            raise ValueError("replacement data must keep the shape")
        self.values = v

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __len__(self):
        return self.values.shape[0]

    def __getattr__(self, item):
        # coordinate access by attribute: da.time, da.realisation, da.model
        coords = self.__dict__.get("coords", {})
        if item in coords:
            c = coords[item]
            if np.ndim(c) == 0:
                return DataArray(np.asarray(c), (), name=item)
            return DataArray(c, (item,), {item: c}, name=item)
        raise AttributeError(item)

    def __getitem__(self, key):
        if isinstance(key, str):
            return getattr(self, key)
        if isinstance(key, (int, np.integer)):
            return self.isel(**{self.dims[0]: int(key)})
        raise TypeError("only integer indexing along the first dim (or coordinate names) is supported")

    def __setitem__(self, key, value):
        if isinstance(key, str):  # da['time'] = new_time  (data.py:560)
            self.coords[key] = np.asarray(getattr(value, "values", value))
            return
        raise TypeError("only coordinate assignment is supported")

    def __repr__(self):
        return f"<labelled.DataArray {self.name or ''} dims={self.dims} shape={self.shape}>"

    # ---- selection ----------------------------------------------------------------------
    def isel(self, **indexers):
        out = self
        for dim, idx in indexers.items():
            ax = out.dims.index(dim)
            vals = np.take(out.values, idx, axis=ax)  # raises IndexError when out of range
            coords = dict(out.coords)
            if np.ndim(idx) == 0:
                dims = out.dims[:ax] + out.dims[ax + 1:]
                coords[dim] = np.asarray(out.coords[dim])[idx]
            else:
                dims = out.dims
                coords[dim] = np.asarray(out.coords[dim])[idx]
            out = DataArray(vals, dims, coords, out.name)
        return out

    def sel(self, **indexers):
        out = self
        for dim, label in indexers.items():
            c = np.asarray(out.coords[dim])
            hits = np.nonzero(c == label)[0]
            if hits.size == 0:
                raise KeyError(label)
            out = out.isel(**{dim: int(hits[0])})
        return out

    def drop_vars(self, names, errors="raise"):
        names = [names] if isinstance(names, str) else list(names)
        coords = {k: v for k, v in self.coords.items() if k not in names or k in self.dims}
        return DataArray(self.values, self.dims, coords, self.name)

    drop = drop_vars

    def assign_coords(self, **kw):
        coords = dict(self.coords)
        for k, v in kw.items():
            coords[k] = v if np.ndim(v) == 0 else np.asarray(getattr(v, "values", v))
        return DataArray(self.values, self.dims, coords, self.name)

    def expand_dims(self, dim=None, axis=0, **dim_kwargs):
        if dim is not None and not dim_kwargs:
            dim_kwargs = {dim: 1} if isinstance(dim, str) else dict(dim)
        out = self
        for d, c in dim_kwargs.items():
            c = getattr(c, "values", c)
            n = int(c) if np.ndim(c) == 0 else len(c)
            coord = np.arange(n) if np.ndim(c) == 0 else np.asarray(c)
            vals = np.expand_dims(out.values, axis)
            vals = np.repeat(vals, n, axis=axis)
            dims = out.dims[:axis] + (d,) + out.dims[axis:]
            coords = {k: v for k, v in out.coords.items()}
            coords[d] = coord
            out = DataArray(vals, dims, coords, out.name)
        return out

    def rename(self, name):
        return DataArray(self.values, self.dims, self.coords, name)

    def copy(self, deep=True, data=None):
        vals = self.values if data is None else np.asarray(data)
        if data is not None and vals.shape != self.values.shape:
            raise ValueError("copy(data=...) must keep the shape")
        if deep and data is None:
            vals = vals.copy()
        return DataArray(vals, self.dims, _copy.deepcopy(self.coords) if deep else dict(self.coords), self.name)

    def __deepcopy__(self, memo):
        return self.copy(deep=True)

    def transpose(self, *dims):
        order = [self.dims.index(d) for d in dims]
        return DataArray(np.transpose(self.values, order), dims, self.coords, self.name)

    # ---- reductions (skip NaN like xarray's default skipna for floats) -------------------
    def _reduce(self, fn, nanfn, dim, **kw):
        if dim is None:
            vals = self.values
            if vals.dtype.kind == "f" and np.isnan(vals).any():
                return DataArray(nanfn(vals, **kw), (), name=self.name)
            return DataArray(fn(vals, **kw), (), name=self.name)
        dims = [dim] if isinstance(dim, str) else list(dim)
        axes = tuple(self.dims.index(d) for d in dims)
        use_nan = self.values.dtype.kind == "f" and np.isnan(self.values).any()
        with np.errstate(invalid="ignore", divide="ignore"):
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                vals = (nanfn if use_nan else fn)(self.values, axis=axes, **kw)
        new_dims = tuple(d for d in self.dims if d not in dims)
        coords = {k: v for k, v in self.coords.items() if k not in dims}
        return DataArray(vals, new_dims, coords, self.name)

    def mean(self, dim=None):
        return self._reduce(np.mean, np.nanmean, dim)

    def std(self, dim=None):
        return self._reduce(np.std, np.nanstd, dim)

    def var(self, dim=None):
        return self._reduce(np.var, np.nanvar, dim)

    def sum(self, dim=None):
        return self._reduce(np.sum, np.nansum, dim)

    def max(self, dim=None):
        return self._reduce(np.max, np.nanmax, dim)

    def min(self, dim=None):
        return self._reduce(np.min, np.nanmin, dim)

    # ---- arithmetic (broadcast by dim name) ------------------------------------------------
    def _binary(self, other, op, reflexive=False):
        if isinstance(other, DataArray):
            dims = list(self.dims) + [d for d in other.dims if d not in self.dims]
            a = _align(self, dims)
            b = _align(other, dims)
            coords = dict(other.coords)
            coords.update(self.coords)
            vals = op(b, a) if reflexive else op(a, b)
            return DataArray(vals, dims, {k: v for k, v in coords.items() if k in dims or np.ndim(v) == 0}, self.name)
        o = np.asarray(other)
        vals = op(o, self.values) if reflexive else op(self.values, o)
        return DataArray(vals, self.dims, self.coords, self.name)

    def __add__(self, o):
        return self._binary(o, np.add)

    def __radd__(self, o):
        return self._binary(o, np.add, True)

    def __sub__(self, o):
        return self._binary(o, np.subtract)

    def __rsub__(self, o):
        return self._binary(o, np.subtract, True)

    def __mul__(self, o):
        return self._binary(o, np.multiply)

    def __rmul__(self, o):
        return self._binary(o, np.multiply, True)

    def __truediv__(self, o):
        with np.errstate(invalid="ignore", divide="ignore"):
            return self._binary(o, np.divide)

    def __rtruediv__(self, o):
        with np.errstate(invalid="ignore", divide="ignore"):
            return self._binary(o, np.divide, True)

    def __pow__(self, o):
        return self._binary(o, np.power)

    def __neg__(self):
        return DataArray(-self.values, self.dims, self.coords, self.name)

    def __eq__(self, o):  # elementwise, as xarray
        return self._binary(o, np.equal)

    def __ne__(self, o):
        return self._binary(o, np.not_equal)

    __hash__ = None


def _align(da: DataArray, dims):
    vals = da.values
    shape = []
    src = []
    for d in dims:
        if d in da.dims:
            src.append(da.dims.index(d))
            shape.append(vals.shape[da.dims.index(d)])
        else:
            shape.append(1)
    vals = np.transpose(vals, src) if src else vals
    return vals.reshape(shape)


def concat(arrays, dim):
    """``xr.concat(list_of_arrays, dim=...)`` along a NEW dim whose scalar coord each array carries
    (weights.py:118)."""
    arrays = list(arrays)
    vals = np.stack([a.values for a in arrays], axis=0)
    labels = []
    for i, a in enumerate(arrays):
        c = a.coords.get(dim, i)
        labels.append(c.item() if isinstance(c, np.ndarray) and c.ndim == 0 else c)
    coords = {k: v for k, v in arrays[0].coords.items() if k != dim}
    coords[dim] = np.asarray(labels)
    return DataArray(vals, (dim,) + arrays[0].dims, coords, arrays[0].name)


def ones_like(da: DataArray):
    return DataArray(np.ones_like(da.values, dtype=np.float64), da.dims, da.coords, da.name)


def as_labelled(obj) -> DataArray:
    """Accepts this module's DataArray or a real ``xarray.DataArray`` (duck-typed)."""
    if isinstance(obj, DataArray):
        return obj
    if hasattr(obj, "dims") and hasattr(obj, "values") and hasattr(obj, "coords"):
        coords = {}
        for k in obj.coords:
            v = np.asarray(obj.coords[k].values)
            coords[str(k)] = v
        return DataArray(np.asarray(obj.values), [str(d) for d in obj.dims], coords, getattr(obj, "name", None))
    raise TypeError("expected a labelled DataArray (bayesian_ensembling_b200.labelled.DataArray or xarray.DataArray)")
