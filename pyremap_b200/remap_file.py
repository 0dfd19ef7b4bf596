"""File-to-file remap on the GPU.

The reference offers a file-to-file path through the NCO subprocess,
``Remapper.ncremap(in_filename, out_filename, variable_list, overwrite, renormalize,
logger, replace_mpas_fill, parallel_exec)``
(``/root/reference/pyremap/remapper/remapper.py:434-506`` ->
``/root/reference/pyremap/remapper/ncremap.py:117-145``).  :func:`remap_file` takes the
same arguments and keeps the same pre-flight behaviour (``ValueError`` without a map,
silent return when the output exists and ``overwrite`` is false, ``TypeError`` for a
point-collection source, ``ncremap.py:15-28``), but applies the weights with this
package's kernels:

* every variable that carries all source dims goes through ONE streamed
  H2D / kernel / D2H pipeline (``engine.apply_weights_many``); the arithmetic is the
  in-memory path's (``remap_numpy.py:223-297``): ``renormalize`` is
  ``renormalization_threshold``, without it results are divided by ``frac_b``;
* missing values become NaN before the product -- a variable's own ``_FillValue`` /
  ``missing_value`` always, and MPAS' undeclared fill (-9.99999979021477e+33) when
  ``replace_mpas_fill`` is set (what ``ncremap -P mpas`` without ``-C`` does,
  ``ncremap.py:53-56``);
* variables without source dims are copied, variables with some but not all of them are
  dropped (``remap_numpy.py:142-147``);
* the destination dims replace the source dims at the position of the first one and the
  destination descriptor's coordinates are added (``remap_numpy.py:171-198``);
* on output a numeric variable gets a ``_FillValue`` attribute iff it holds NaN, taken by
  dtype from ``netCDF4.default_fillvals`` (reproduced below; the rule of
  ``/root/reference/pyremap/utility.py:35-51``), and its NaNs are stored as that value.

Containers: NetCDF-3 classic / 64-bit offset through ``scipy.io.netcdf_file`` (memory
mapped input: a variable is paged in as the pipeline walks its slices); NetCDF-4 input
through ``netCDF4`` when that package is importable.  Output is NetCDF-3 64-bit offset.
"""

from __future__ import annotations

import os
import sys

import numpy as np

from . import engine

#: ``netCDF4.default_fillvals`` (netCDF-C's NC_FILL_*), keyed like there
DEFAULT_FILLVALS = {
    'S1': '\x00', 'i1': -127, 'u1': 255, 'i2': -32767, 'u2': 65535,
    'i4': -2147483647, 'u4': 4294967295, 'i8': -9223372036854775806,
    'u8': 18446744073709551614, 'f4': 9.969209968386869e+36,
    'f8': 9.969209968386869e+36,
}

#: the value MPAS writes where a field is undefined, without declaring it
MPAS_FILL = -9.99999979021477e+33


def fill_value_for(array, fillvalues=None):
    """``_FillValue`` the reference's writer would attach to ``array``
    (``utility.py:35-51``): numeric with at least one NaN -> the default fill of its dtype,
    anything else -> None."""
    fillvalues = DEFAULT_FILLVALS if fillvalues is None else fillvalues
    array = np.asarray(array)
    if not np.issubdtype(array.dtype, np.number):
        return None
    if array.dtype.kind != 'f' or not np.any(np.isnan(array)):
        return None
    for fill_type, value in fillvalues.items():
        if fill_type != 'S1' and array.dtype == np.dtype(fill_type):
            return value
    return None


def decode_missing(array, attrs, replace_mpas_fill=False):
    """Native-endian copy of a floating-point ``array`` with its missing values as NaN."""
    out = np.array(array, dtype=array.dtype.newbyteorder('='), order='C')
    if out.dtype.kind != 'f':
        return out
    for key in ('_FillValue', 'missing_value'):
        if key in attrs:
            fill = np.asarray(attrs[key]).ravel()
            if fill.size and np.isfinite(fill[0]):
                out[out == out.dtype.type(fill[0])] = np.nan
    if replace_mpas_fill:
        out[out == out.dtype.type(MPAS_FILL)] = np.nan
    return out


class _Input:
    """Variables, dims and attributes of the input file behind one small interface."""

    def __init__(self, filename):
        self.filename = filename
        with open(filename, 'rb') as f:
            magic = f.read(4)
        if magic[:3] == b'CDF':
            from scipy.io import netcdf_file
            self._nc = netcdf_file(filename, 'r', mmap=True)
            self.dims = dict(self._nc.dimensions)       # None = the record dimension
            self.attrs = dict(self._nc._attributes)
            self.names = list(self._nc.variables)
            self._kind = 'scipy'
        else:
            try:
                import netCDF4
            except ImportError as exc:
                raise OSError(
                    f'{filename} is not a NetCDF-3 file and the netCDF4 package is not '
                    f'installed: convert it with "ncks -6" or install netCDF4') from exc
            self._nc = netCDF4.Dataset(filename, 'r')
            self._nc.set_auto_maskandscale(False)
            self.dims = {k: (None if d.isunlimited() else len(d))
                         for k, d in self._nc.dimensions.items()}
            self.attrs = {k: self._nc.getncattr(k) for k in self._nc.ncattrs()}
            self.names = list(self._nc.variables)
            self._kind = 'netcdf4'

    def dim_size(self, dim):
        size = self.dims[dim]
        if size is None:                                 # the record dimension
            size = self._nc._recs if self._kind == 'scipy' else len(self._nc.dimensions[dim])
        return int(size)

    def var_dims(self, name):
        return tuple(self._nc.variables[name].dimensions)

    def var_dtype(self, name):
        v = self._nc.variables[name]
        return np.dtype(v.data.dtype if self._kind == 'scipy' else v.dtype)

    def var_attrs(self, name):
        v = self._nc.variables[name]
        if self._kind == 'scipy':
            return dict(v._attributes)
        return {k: v.getncattr(k) for k in v.ncattrs()}

    def var_shape(self, name):
        return tuple(int(s) for s in self._nc.variables[name].shape)

    def read(self, name):
        v = self._nc.variables[name]
        if self._kind == 'scipy' and v.shape == ():
            return np.asarray(v.getValue())
        return np.asarray(v[...])

    def close(self):
        import warnings
        with warnings.catch_warnings():
            # scipy warns when views of the memory map are still alive: every array taken from
            # it above was copied (decode_missing, the copies of passthrough variables)
            warnings.simplefilter('ignore', RuntimeWarning)
            try:
                self._nc.close()
            except Exception:
                pass


def _validate(remapper, out_filename, overwrite):
    """Pre-flight of ``ncremap.py:15-28`` (minus the search for the NCO executable)."""
    if remapper.map_filename is None:
        raise ValueError('No mapping file has been defined')
    if not overwrite and os.path.exists(out_filename):
        return False
    if type(remapper.src_descriptor).__name__ == 'PointCollectionDescriptor':
        raise TypeError('Source grid is a point collection, which is not supported.')
    return True


def _to_nc3(data):
    """``data`` in a type NetCDF-3 can hold (no unsigned or 64-bit integers there: such
    arrays are stored as int32 when their values fit, as xarray's scipy backend does)."""
    data = np.asarray(data)
    kind, size = data.dtype.kind, data.dtype.itemsize
    if kind == 'f':
        return data if size in (4, 8) else data.astype(np.float32)
    if kind == 'S' and size == 1:
        return data
    if kind == 'i' and size <= 4:
        return data
    if kind in 'iub':
        if data.size == 0 or (data.min() >= -2 ** 31 and data.max() < 2 ** 31):
            return data.astype(np.int32)
        raise ValueError(f'{data.dtype} values do not fit a NetCDF-3 integer')
    raise TypeError(f'no NetCDF-3 type for dtype {data.dtype}')


def remap_file(remapper, in_filename, out_filename, variable_list=None, overwrite=False,
               renormalize=None, logger=None, replace_mpas_fill=False, parallel_exec=None,
               fillvalues=None):
    """Remap the variables of ``in_filename`` into ``out_filename`` (see the module text).

    Arguments as ``Remapper.ncremap`` (``remapper.py:434-506``); ``parallel_exec`` is accepted
    and ignored (there is no subprocess to launch).  Returns the list of variables written,
    or None when the call was skipped because the output exists."""
    from .remap_numpy import _dst_dims, _load_mapping

    if not _validate(remapper, out_filename, overwrite):
        return None
    _load_mapping(remapper)
    src_dims = list(remapper.src_descriptor.dims)
    dst_dims = list(remapper.dst_descriptor.dims)
    dst_sizes = _dst_dims(remapper)
    src_grid_dims = np.asarray(remapper._ds_map['src_grid_dims'].values)[::-1]

    src = _Input(str(in_filename))
    try:
        for index, dim in enumerate(src_dims):
            if dim not in src.dims:
                raise ValueError(f'source dimension {dim} is not in {in_filename}')
            size = src.dim_size(dim)
            if src_grid_dims[index] != size:
                raise ValueError(
                    f"data set and remapping source dimension {dim} don't "
                    f'have the same size: {src_grid_dims[index]} != {size}')

        wanted = list(src.names) if variable_list is None else list(variable_list)
        missing = [v for v in wanted if v not in src.names]
        if missing:
            raise KeyError(f'variables not in {in_filename}: {missing}')

        plan = []          # (name, dims, 'remap' | 'copy', src axes)
        for name in wanted:
            dims = src.var_dims(name)
            present = [d in dims for d in src_dims]
            if name in dst_dims and not any(present):
                continue                              # replaced by the destination coordinate
            if any(present) and not all(present):
                continue                              # remap_numpy.py:142-147
            if not any(present):
                plan.append((name, dims, 'copy', None))
                continue
            if src.var_dtype(name).kind not in 'fiu':
                continue                              # characters on the source grid
            axes = [i for i, d in enumerate(dims) if d in src_dims]
            plan.append((name, dims, 'remap', axes))

        fields, attrs_of = [], {}
        for name, dims, kind, axes in plan:
            attrs_of[name] = src.var_attrs(name)
            if kind == 'remap':
                raw = src.read(name)
                field = decode_missing(raw, attrs_of[name], replace_mpas_fill)
                if field.dtype.kind != 'f':
                    field = field.astype(np.float64)
                fields.append((field, axes))
        if logger is not None:
            logger.info('remap_file: %d variable(s) on the source grid, %d copied',
                        len(fields), len(plan) - len(fields))
        results = engine.apply_weights_many(
            remapper._matrix, dst_sizes, fields, renormalize,
            device=getattr(remapper, 'device', None)) if fields else []

        written = _write(remapper, src, out_filename, plan, results, attrs_of, src_dims,
                         dst_dims, dst_sizes, fillvalues)
    finally:
        src.close()
    return written


def _out_dims(dims, src_dims, dst_dims):
    out, added = [], False
    for d in dims:
        if d in src_dims:
            if not added:
                out.extend(dst_dims)
                added = True
        else:
            out.append(d)
    return tuple(out)


def _write(remapper, src, out_filename, plan, results, attrs_of, src_dims, dst_dims, dst_sizes,
           fillvalues):
    from scipy.io import netcdf_file

    tmp = f'{out_filename}.tmp{os.getpid()}'
    written = []
    out = netcdf_file(tmp, 'w', version=2)
    try:
        # dimensions: the destination dims, then every other dim a written variable uses
        sizes = dict(zip(dst_dims, dst_sizes))
        record = [d for d, n in src.dims.items() if n is None]
        coords = dict(remapper.dst_descriptor.coords or {})
        needed = list(dst_dims)
        for name, dims, kind, _ in plan:
            for d in (_out_dims(dims, src_dims, dst_dims) if kind == 'remap' else dims):
                if d not in needed:
                    needed.append(d)
        for cname, c in coords.items():
            for d, n in zip(c['dims'], np.shape(c['data'])):
                if d not in needed:
                    needed.append(d)
                sizes.setdefault(d, int(n))
        needed.sort(key=lambda d: d not in record)        # the record dimension comes first
        for d in needed:
            if d in record:
                out.createDimension(d, None)
            else:
                if d not in sizes:
                    sizes[d] = src.dim_size(d)
                out.createDimension(d, sizes[d])

        def put(name, data, dims, attrs, fill):
            data = np.asarray(data)
            if data.dtype.kind == 'f' and fill is not None:
                data = np.where(np.isnan(data), data.dtype.type(fill), data)
            data = _to_nc3(data)
            var = out.createVariable(name, data.dtype.newbyteorder('='), dims)
            for k, v in attrs.items():
                if k not in ('_FillValue', 'missing_value'):
                    setattr(var, k, v)
            if fill is not None:
                setattr(var, '_FillValue', np.asarray(fill, dtype=data.dtype))
            if tuple(dims) == ():
                var.data[...] = data          # (scipy's assignValue indexes 0-d data with [:])
            else:
                var[:] = data
            written.append(name)

        # destination coordinates first (remap_numpy.py:198)
        for cname, c in coords.items():
            data = np.asarray(c['data'])
            put(cname, data, tuple(c['dims']), dict(c.get('attrs', {})),
                fill_value_for(data, fillvalues))

        k = 0
        for name, dims, kind, _ in plan:
            if name in coords:
                if kind == 'remap':
                    k += 1
                continue
            if kind == 'remap':
                data = results[k]
                k += 1
                put(name, data, _out_dims(dims, src_dims, dst_dims), attrs_of[name],
                    fill_value_for(data, fillvalues))
            else:
                raw = np.asarray(src.read(name))
                if raw.dtype.kind == 'f':
                    # a declared fill value becomes NaN and is encoded again by the writer's rule
                    data = decode_missing(raw, attrs_of[name])
                    fill = fill_value_for(data, fillvalues)
                else:
                    data = np.array(raw, dtype=raw.dtype.newbyteorder('='))
                    declared = attrs_of[name].get('_FillValue')
                    fill = None if declared is None or data.dtype.kind == 'S' else \
                        np.asarray(declared).ravel()[0]
                put(name, data, dims, attrs_of[name], fill)

        # global attributes: the input's, plus history / mesh_name as remap_numpy.py:59-67
        for key, value in src.attrs.items():
            if key not in ('history', 'mesh_name'):
                setattr(out, key, value)
        current_hist = ' '.join(sys.argv[:])
        hist = src.attrs.get('history')
        if isinstance(hist, bytes):
            hist = hist.decode()
        out.history = '\n'.join([hist, current_hist]) if hist else current_hist
        out.mesh_name = str(remapper.dst_descriptor.mesh_name)
        out.close()
        os.replace(tmp, out_filename)
    except BaseException:
        try:
            out.close()
        except Exception:
            pass
        if os.path.exists(tmp):
            os.remove(tmp)
        raise
    return written
