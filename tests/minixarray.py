"""A very small stand-in for the slice of xarray that pyremap's in-memory remap
path touches.  TEST INFRASTRUCTURE ONLY (xarray is not installed in this image).

It exists so that (a) the *reference's* ``_remap_numpy`` / ``_load_mapping`` /
``_remap_data_array`` (``/root/reference/pyremap/remapper/remap_numpy.py``) can be
executed end to end in the authoring container to produce golden fixtures, and
(b) the product's xarray-facing plumbing can be tested on the GPU box with the
very same container types.  With real xarray installed, the product uses real
xarray and never sees this file.

Semantics implemented (and nothing more):

* ``DataArray``: ``dims``, ``shape``, ``sizes``, ``values``, ``coords`` (mapping of
  name -> DataArray), ``attrs``, ``name``, ``from_dict`` including xarray's
  MaskedArray -> NaN-filled conversion (``xarray.core.variable.as_compatible_data``).
* ``Dataset``: ``data_vars``, ``coords``, ``attrs``, ``sizes``, ``__getitem__``,
  ``__contains__``, ``drop_vars``, ``map(func, keep_attrs, args)``.
* ``open_dataset``: ``.npz`` map files (the synthetic maps of this repo) and
  NetCDF-3 files through ``scipy.io.netcdf_file``.
"""

from __future__ import annotations

import numpy as np

# dimension names of the variables of an ESMF/SCRIP mapping file
_MAP_DIMS = {
    'S': ('n_s',), 'row': ('n_s',), 'col': ('n_s',),
    'frac_a': ('n_a',), 'area_a': ('n_a',), 'mask_a': ('n_a',),
    'yc_a': ('n_a',), 'xc_a': ('n_a',),
    'frac_b': ('n_b',), 'area_b': ('n_b',), 'mask_b': ('n_b',),
    'yc_b': ('n_b',), 'xc_b': ('n_b',),
    'src_grid_dims': ('src_grid_rank',), 'dst_grid_dims': ('dst_grid_rank',),
}


def _compatible_data(data):
    if isinstance(data, np.ma.MaskedArray):
        mask = np.ma.getmaskarray(data)
        if mask.any():
            raw = np.ma.getdata(data)
            if raw.dtype.kind in 'iub':
                raw = raw.astype(np.float64)
            return np.where(~mask, raw, np.array(np.nan, dtype=raw.dtype))
        return np.asarray(data)
    return np.asarray(data)


class DataArray:
    def __init__(self, data, dims=None, coords=None, attrs=None, name=None):
        self._data = _compatible_data(data)
        if dims is None:
            dims = tuple(f'dim_{i}' for i in range(self._data.ndim))
        self.dims = tuple(dims)
        if len(self.dims) != self._data.ndim:
            raise ValueError(f'{len(self.dims)} dims for {self._data.ndim}-d data')
        self.attrs = dict(attrs) if attrs else {}
        self.name = name
        self.coords = {}
        for key, val in (coords or {}).items():
            if not isinstance(val, DataArray):
                val = DataArray(val, dims=(key,), name=key)
            self.coords[key] = val

    @property
    def values(self):
        return self._data

    @property
    def shape(self):
        return self._data.shape

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def ndim(self):
        return self._data.ndim

    @property
    def sizes(self):
        return dict(zip(self.dims, self._data.shape))

    @classmethod
    def from_dict(cls, d):
        coords = {}
        for key, val in d.get('coords', {}).items():
            coords[key] = DataArray(val['data'], dims=val['dims'],
                                    attrs=val.get('attrs'), name=key)
        return cls(d['data'], dims=d.get('dims'), coords=coords,
                   attrs=d.get('attrs'), name=d.get('name'))

    def __repr__(self):
        return f'<mini DataArray {self.name} {self.sizes}>'


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self.data_vars = {}
        self.coords = {}
        self.attrs = dict(attrs) if attrs else {}
        for key, val in (coords or {}).items():
            if not isinstance(val, DataArray):
                val = DataArray(val, dims=(key,), name=key)
            self.coords[key] = val
        for key, val in (data_vars or {}).items():
            if isinstance(val, tuple):
                val = DataArray(val[1], dims=val[0],
                                attrs=val[2] if len(val) > 2 else None)
            if not isinstance(val, DataArray):
                raise TypeError('data_vars values must be DataArray or tuple')
            val.name = key
            for ckey, cval in val.coords.items():
                self.coords.setdefault(ckey, cval)
            self.data_vars[key] = val

    @property
    def sizes(self):
        out = {}
        for var in list(self.coords.values()) + list(self.data_vars.values()):
            for dim, size in var.sizes.items():
                if out.setdefault(dim, size) != size:
                    raise ValueError(f'conflicting sizes for dimension {dim}')
        return out

    def __contains__(self, key):
        return key in self.data_vars or key in self.coords

    def __getitem__(self, key):
        if key in self.data_vars:
            var = self.data_vars[key]
            coords = {k: v for k, v in self.coords.items()
                      if set(v.dims) <= set(var.dims)}
            return DataArray(var.values, dims=var.dims, coords=coords,
                             attrs=var.attrs, name=key)
        return self.coords[key]

    def drop_vars(self, names):
        if isinstance(names, str):
            names = [names]
        keep = {k: v for k, v in self.data_vars.items() if k not in names}
        coords = {k: v for k, v in self.coords.items() if k not in names}
        return Dataset(keep, coords=coords, attrs=self.attrs)

    def map(self, func, keep_attrs=None, args=(), **kwargs):
        out = {}
        for key in self.data_vars:
            res = func(self[key], *args, **kwargs)
            if not isinstance(res, DataArray):
                res = DataArray(res, dims=self.data_vars[key].dims)
            if keep_attrs:
                res.attrs = dict(self.data_vars[key].attrs)
            out[key] = res
        return Dataset(out, attrs=self.attrs if keep_attrs else None)

    def close(self):
        pass

    def __repr__(self):
        return f'<mini Dataset {list(self.data_vars)} {self.sizes}>'


def open_dataset(filename, **_):
    filename = str(filename)
    if filename.endswith('.npz'):
        with np.load(filename, allow_pickle=False) as npz:
            data = {}
            for key in npz.files:
                arr = npz[key]
                dims = _MAP_DIMS.get(key)
                if dims is None or len(dims) != arr.ndim:
                    dims = tuple(f'{key}_dim{i}' for i in range(arr.ndim))
                data[key] = DataArray(arr, dims=dims)
        return Dataset(data)
    from scipy.io import netcdf_file
    with netcdf_file(filename, 'r', mmap=False) as nc:
        data = {}
        for key, var in nc.variables.items():
            arr = np.array(var[...])
            scale = getattr(var, 'scale_factor', None)
            fill = getattr(var, '_FillValue', None)
            if arr.dtype.kind == 'f' and fill is not None:
                arr = np.where(arr == fill, np.nan, arr).astype(arr.dtype)
            if scale is not None:
                arr = arr * scale
            attrs = {k: v for k, v in var._attributes.items()}
            data[key] = DataArray(arr, dims=var.dimensions, attrs=attrs)
        attrs = {k: v for k, v in nc._attributes.items()}
    return Dataset(data, attrs=attrs)
