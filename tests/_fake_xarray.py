"""A small stand-in for the parts of xarray (and of a dask-backed array) that ``marex_b200.xr_api`` touches, so that the
xarray boundary of the drop-in -- DataArray in, Dataset out, dims / dtypes / coords / attrs / chunking as
``marEx.preprocess_data`` produces them (detect.py:679-683, 718-828) and as ``marEx.tracker`` demands them
(track.py:411-418, 505-591, 640-668) -- executes in an image that has neither xarray nor dask.  Test infrastructure only.

Semantics follow xarray where the adapter relies on them: named dimensions, coordinate variables that travel through
``transpose`` / ``isel``, ``Dataset.__setitem__`` with ``(dims, data)`` tuples, ``.chunk`` producing a lazily evaluated
("dask-backed") variable, ``.values`` computing it.
"""
from collections import OrderedDict
from typing import Any, Dict, Iterable, Optional, Tuple

import numpy as np


class FakeDaskArray:
    """Lazy, chunked view of a numpy array: what ``da.data`` is for a dask-backed DataArray.  ``computed`` counts how
    often (and how much of) the array was materialised, so tests can see that the adapter streams chunk by chunk."""

    def __init__(self, array: np.ndarray, chunks: Tuple[Tuple[int, ...], ...], log: Optional[list] = None):
        self._a = array
        self.chunks = chunks
        self.dask = {"fake-graph": None}  # the attribute xr_api._is_dask looks for
        self.log = log if log is not None else []

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)
    ndim = property(lambda self: self._a.ndim)

    def __array__(self, dtype=None, copy=None):
        self.log.append(self._a.shape)
        return np.asarray(self._a, dtype=dtype)

    def compute(self):
        return np.asarray(self)

    def transpose(self, axes):
        return FakeDaskArray(self._a.transpose(axes), tuple(self.chunks[i] for i in axes), self.log)

    def __getitem__(self, key):
        sub = self._a[key]
        chunks = tuple((n,) for n in sub.shape)
        return FakeDaskArray(sub, chunks, self.log)

    def astype(self, dtype):
        return FakeDaskArray(self._a.astype(dtype), self.chunks, self.log)


def _normalise_chunks(shape, dims, spec: Dict[str, int]):
    out = []
    for n, d in zip(shape, dims):
        c = spec.get(d, -1) if isinstance(spec, dict) else -1
        if c in (-1, None) or c >= n:
            out.append((n,))
        else:
            full, rest = divmod(n, c)
            out.append((c,) * full + ((rest,) if rest else ()))
    return tuple(out)


class Coords(OrderedDict):
    pass


class DataArray:
    def __init__(self, data, dims: Iterable[str] = None, coords: Optional[Dict[str, Any]] = None, attrs=None, name=None):
        self.data = data
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(np.ndim(data)))
        assert len(self.dims) == np.ndim(data), (self.dims, np.shape(data))
        self.attrs = dict(attrs or {})
        self.name = name
        self.coords = Coords()
        for k, v in (coords or {}).items():
            self.coords[k] = v if isinstance(v, DataArray) else DataArray(np.asarray(v), dims=(k,))

    # ---- array protocol
    shape = property(lambda self: tuple(self.data.shape))
    dtype = property(lambda self: self.data.dtype)
    ndim = property(lambda self: len(self.dims))

    @property
    def values(self):
        return np.asarray(self.data)

    @property
    def chunks(self):
        return getattr(self.data, "chunks", None)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.coords[key]
        raise TypeError("positional indexing is not part of the fake")

    def __getattr__(self, name):  # `neighbours.nv`: coordinate access by attribute
        coords = self.__dict__.get("coords", {})
        if name in coords:
            return coords[name]
        raise AttributeError(name)

    def _coords_for(self, dims, indexers=None):
        out = {}
        for k, v in self.coords.items():
            if all(d in dims for d in v.dims):
                if indexers and any(d in indexers for d in v.dims):
                    idx = tuple(indexers.get(d, slice(None)) for d in v.dims)
                    v = DataArray(np.asarray(v.data)[idx], dims=v.dims, attrs=v.attrs)
                out[k] = v
        return out

    def transpose(self, *dims):
        axes = tuple(self.dims.index(d) for d in dims)
        data = self.data.transpose(axes)
        return DataArray(data, dims=dims, coords=self._coords_for(dims), attrs=self.attrs, name=self.name)

    def isel(self, indexers: Dict[str, Any]):
        key = tuple(indexers.get(d, slice(None)) for d in self.dims)
        key = tuple(np.nonzero(k)[0] if isinstance(k, np.ndarray) and k.dtype == bool else k for k in key)
        idx = {d: k for d, k in zip(self.dims, key)}
        return DataArray(self.data[key], dims=self.dims, coords=self._coords_for(self.dims, idx), attrs=self.attrs, name=self.name)

    def astype(self, dtype):
        return DataArray(self.data.astype(dtype), dims=self.dims, coords=dict(self.coords), attrs=self.attrs, name=self.name)

    def chunk(self, chunks):
        a = np.asarray(self.data) if not isinstance(self.data, FakeDaskArray) else self.data._a
        return DataArray(FakeDaskArray(a, _normalise_chunks(a.shape, self.dims, chunks)), dims=self.dims,
                         coords=dict(self.coords), attrs=self.attrs, name=self.name)  # fmt: skip

    def compute(self):
        return DataArray(np.asarray(self.data), dims=self.dims, coords=dict(self.coords), attrs=self.attrs, name=self.name)


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.coords = Coords()
        for k, v in (coords or {}).items():
            self.coords[k] = v if isinstance(v, DataArray) else DataArray(np.asarray(v), dims=(k,))
        self.data_vars: "OrderedDict[str, DataArray]" = OrderedDict()
        self.attrs = dict(attrs or {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    def __setitem__(self, name, value):
        if isinstance(value, tuple):
            dims, data = value[0], value[1]
            value = DataArray(data, dims=dims)
        assert isinstance(value, DataArray), type(value)
        for d, n in zip(value.dims, value.shape):
            if d in self.coords and self.coords[d].dims == (d,):
                assert self.coords[d].shape[0] == n, f"conflicting sizes for dimension {d!r}: {n} vs {self.coords[d].shape[0]}"
        for k, v in value.coords.items():  # a DataArray brings its coordinates along
            if k not in self.coords:
                self.coords[k] = v
        self.data_vars[name] = DataArray(value.data, dims=value.dims, attrs=value.attrs, name=name)

    def __getitem__(self, name):
        if name in self.data_vars:
            v = self.data_vars[name]
            v.coords = Coords((k, c) for k, c in self.coords.items() if all(d in v.dims for d in c.dims))
            return v
        return self.coords[name]

    def __contains__(self, name):
        return name in self.data_vars or name in self.coords

    def __getattr__(self, name):
        dv = self.__dict__.get("data_vars", {})
        if name in dv:
            return self[name]
        raise AttributeError(name)

    def assign_coords(self, **kw):
        for k, v in kw.items():
            self.coords[k] = v if isinstance(v, DataArray) else DataArray(np.asarray(v), dims=(k,))
        return self

    @property
    def dims(self):
        out = OrderedDict()
        for v in self.data_vars.values():
            for d, n in zip(v.dims, v.shape):
                out[d] = n
        return out
