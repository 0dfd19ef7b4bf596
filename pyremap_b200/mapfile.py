"""Mapping-file loading: SCRIP/ESMF weights -> row-sorted canonical CSR on the host.

Mirrors the one-time part of the reference's ``_load_mapping``
(``/root/reference/pyremap/remapper/remap_numpy.py:72-139``): read ``n_a``, ``n_b``,
``S``, ``row``, ``col`` (1-based), ``frac_b``, ``src_grid_dims``, ``dst_grid_dims``
and build the matrix ``csr_matrix((S, (row-1, col-1)), shape=(n_b, n_a))``
(``:134-137``) -- i.e. rows = destination cells, columns sorted within a row,
duplicate ``(row, col)`` entries summed.  The CSR builder here is our own (NumPy);
it does not call scipy.

Readers: real xarray if importable (any NetCDF flavour it supports), ``.npz``
(synthetic maps of this repo), NetCDF-3 via ``scipy.io.netcdf_file``, NetCDF-4
via ``netCDF4``/``h5py`` when present.
"""

from __future__ import annotations

import numpy as np


class _Var:
    """Minimal variable view (``.values``) so ``remapper._ds_map['frac_b'].values``
    keeps working when xarray is not installed (reference remap_numpy.py:252,270)."""

    def __init__(self, values):
        self.values = values


class MapDataset:
    """Dictionary-backed stand-in for the opened map ``xr.Dataset``."""

    def __init__(self, variables, sizes):
        self._vars = {k: _Var(v) for k, v in variables.items()}
        self.sizes = dict(sizes)

    def __getitem__(self, key):
        return self._vars[key]

    def __contains__(self, key):
        return key in self._vars

    def close(self):
        pass


_NEEDED = ('S', 'row', 'col', 'frac_b', 'src_grid_dims', 'dst_grid_dims')


def _from_arrays(get):
    arrays = {k: np.asarray(get(k)) for k in _NEEDED}
    sizes = {
        'n_s': arrays['S'].size,
        'n_b': arrays['frac_b'].size,
        'src_grid_rank': arrays['src_grid_dims'].size,
        'dst_grid_rank': arrays['dst_grid_dims'].size,
    }
    return arrays, sizes


def open_map(filename):
    """Open a mapping file; returns an object offering ``ds['S'].values`` and
    ``ds.sizes['n_a']`` (an ``xr.Dataset`` when xarray is installed)."""
    filename = str(filename)
    if filename.endswith('.npz'):
        with np.load(filename, allow_pickle=False) as npz:
            arrays, sizes = _from_arrays(lambda k: npz[k])
            if 'frac_a' in npz.files:
                sizes['n_a'] = int(npz['frac_a'].size)
            elif 'n_a' in npz.files:
                sizes['n_a'] = int(npz['n_a'])
            else:
                sizes['n_a'] = int(np.prod(arrays['src_grid_dims']))
        return MapDataset(arrays, sizes)
    try:
        import xarray as xr
        if hasattr(xr, 'open_dataset'):
            return xr.open_dataset(filename)
    except ImportError:
        pass
    with open(filename, 'rb') as fh:
        magic = fh.read(4)
    if magic[:3] == b'CDF':
        from scipy.io import netcdf_file
        with netcdf_file(filename, 'r', mmap=False) as nc:
            arrays, sizes = _from_arrays(lambda k: np.array(nc.variables[k][...]))
            sizes['n_a'] = int(nc.dimensions['n_a'])
        return MapDataset(arrays, sizes)
    try:
        import netCDF4
        with netCDF4.Dataset(filename) as nc:
            arrays, sizes = _from_arrays(lambda k: np.array(nc.variables[k][...]))
            sizes['n_a'] = int(nc.dimensions['n_a'].size)
        return MapDataset(arrays, sizes)
    except ImportError:
        pass
    try:
        import h5py
        with h5py.File(filename, 'r') as h5:
            arrays, sizes = _from_arrays(lambda k: np.array(h5[k]))
            sizes['n_a'] = int(h5['n_a'].shape[0]) if 'n_a' in h5 else int(
                np.prod(arrays['src_grid_dims']))
        return MapDataset(arrays, sizes)
    except ImportError:
        pass
    raise OSError(
        f'cannot read {filename}: it is NetCDF-4/HDF5 and none of xarray, '
        'netCDF4 or h5py is installed')


def coo_to_csr(S, row0, col0, n_row, n_col):
    """Canonical CSR (sorted columns, duplicates summed) from 0-based triplets.

    Same result as ``scipy.sparse.csr_matrix((S, (row, col)))`` used by the
    reference (remap_numpy.py:137): entries are ordered by (row, col) with a
    *stable* sort and equal (row, col) pairs are added left to right in file
    order.  (scipy's in-row sort is std::sort; for three or more duplicates of
    one (row, col) in a row longer than 16 its addition order is unspecified, so
    bit-equality with scipy is guaranteed for up to two duplicates per pair --
    map files written by ESMF/MOAB contain none.)
    """
    S = np.asarray(S, dtype=np.float64).ravel()
    row0 = np.asarray(row0).ravel()
    col0 = np.asarray(col0).ravel()
    if not (S.size == row0.size == col0.size):
        raise ValueError('S, row and col must have the same length')
    n_row, n_col = int(n_row), int(n_col)
    if S.size:
        if row0.min() < 0 or row0.max() >= n_row:
            raise ValueError('row index out of range for n_b')
        if col0.min() < 0 or col0.max() >= n_col:
            raise ValueError('col index out of range for n_a')
    if S.size >= 2**31 - 1 or n_row >= 2**31 - 1 or n_col >= 2**31 - 1:
        raise ValueError('int32 CSR only: sizes must be below 2**31')
    key = row0.astype(np.int64) * n_col + col0.astype(np.int64)
    if S.size > 1:
        step = np.diff(key)
        if not np.all(step > 0):                       # not already canonical
            order = np.argsort(key, kind='stable')
            key = key[order]
            S = S[order]
            dup = np.concatenate([[False], np.diff(key) == 0])
            if dup.any():
                starts = np.nonzero(~dup)[0]
                # left-to-right sums within each run of equal keys
                summed = S[starts].copy()
                run = np.cumsum(~dup) - 1
                for j in np.nonzero(dup)[0]:
                    summed[run[j]] = summed[run[j]] + S[j]
                S = summed
                key = key[starts]
    rows = (key // n_col).astype(np.int64)
    indices = (key - rows * n_col).astype(np.int32)
    counts = np.bincount(rows, minlength=n_row)
    indptr = np.zeros(n_row + 1, dtype=np.int32)
    np.cumsum(counts, out=indptr[1:])
    return indptr, indices, np.ascontiguousarray(S, dtype=np.float64)


def coo_to_csr_gpu(S, row0, col0, n_row, n_col, device=None):
    """:func:`coo_to_csr` on the GPU (``b200remap_coo_to_csr``): counting pass, scan, one warp
    per row ranking its entries by (col, file position), duplicate runs summed in file order.
    Bit-identical to :func:`coo_to_csr`; ~50x faster for maps with tens of millions of weights."""
    import torch

    from . import _cabi
    S = np.ascontiguousarray(S, dtype=np.float64).ravel()
    row0 = np.asarray(row0).ravel()
    col0 = np.asarray(col0).ravel()
    if not (S.size == row0.size == col0.size):
        raise ValueError('S, row and col must have the same length')
    n_row, n_col = int(n_row), int(n_col)
    if S.size >= 2**31 - 1 or n_row >= 2**31 - 1 or n_col >= 2**31 - 1:
        raise ValueError('int32 CSR only: sizes must be below 2**31')
    # range checks on the ORIGINAL integers: an int64 index of a corrupt map must not wrap
    # into the valid range when it is narrowed to the int32 the device builder takes
    if S.size:
        if row0.min() < 0 or row0.max() >= n_row:
            raise ValueError('row index out of range for n_b')
        if col0.min() < 0 or col0.max() >= n_col:
            raise ValueError('col index out of range for n_a')
    row32 = np.ascontiguousarray(row0, dtype=np.int32)
    col32 = np.ascontiguousarray(col0, dtype=np.int32)
    # the current device unless told otherwise (one process per GPU: never touch device 0
    # from every rank)
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))
    with torch.cuda.device(dev):
        indptr = torch.empty(n_row + 1, dtype=torch.int32, device=dev)
        indices = torch.empty(max(S.size, 1), dtype=torch.int32, device=dev)
        data = torch.empty(max(S.size, 1), dtype=torch.float64, device=dev)
        try:
            nnz = _cabi.coo_to_csr_device(dev.index, n_row, n_col, row32, col32, S,
                                          indptr.data_ptr(), indices.data_ptr(), data.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream)
        except _cabi.B200RemapError as exc:
            if 'out of range' in str(exc):
                raise ValueError(str(exc).split(': ', 1)[-1]) from exc
            raise
        return (indptr.cpu().numpy(), indices[:nnz].cpu().numpy(), data[:nnz].cpu().numpy())


class WeightMatrix:
    """Host-side canonical CSR of the map plus lazily created device copies.

    Takes the place of ``remapper._matrix`` (a ``scipy.sparse.csr_matrix`` in the
    reference, remap_numpy.py:137).  ``shape``, ``nnz``, ``indptr``, ``indices``,
    ``data`` are spelled like scipy's so diagnostics keep working; ``dot`` is
    intentionally absent -- products run on the GPU only.
    """

    def __init__(self, indptr, indices, data, shape, frac_b=None):
        self.indptr = indptr
        self.indices = indices
        self.data = data
        self.shape = (int(shape[0]), int(shape[1]))
        self.frac_b = None if frac_b is None else np.ascontiguousarray(
            frac_b, dtype=np.float64)
        self._device = {}

    @property
    def nnz(self):
        return int(self.data.size)

    def on_device(self, device=0):
        """The ``DeviceCSR`` handle for CUDA device ``device`` (created once)."""
        from ._cabi import DeviceCSR
        device = int(device)
        if device not in self._device:
            self._device[device] = DeviceCSR(self.indptr, self.indices,
                                             self.data, self.frac_b,
                                             self.shape[1], device)
        return self._device[device]

    def _touched_rows(self):
        """Sorted distinct source rows the map references (O(nnz), no sort)."""
        if getattr(self, '_touched', None) is None:
            seen = np.zeros(self.shape[1], dtype=bool)
            seen[self.indices] = True
            self._touched = np.flatnonzero(seen)
        return self._touched

    def cover(self, max_runs=64, worthwhile=0.7):
        """Source rows a host->device copy has to bring over, as a few contiguous runs.

        Regional maps touch a small part of the source mesh (BASELINE config 3: 9 %).
        Returns ``None`` when (nearly) everything is touched, else a dict with
        ``runs`` = list of ``(start, length, position)`` covering every touched source
        row (gaps are bridged until at most ``max_runs`` runs remain), ``n_cover`` and
        ``indices`` (column indices renumbered into the covered rows; the mapping is
        monotonic, so rows stay in canonical order).
        """
        if getattr(self, '_cover', False) is not False:
            return self._cover
        self._cover = None
        n_a = self.shape[1]
        touched = self._touched_rows()
        if touched.size and touched.size <= worthwhile * n_a:
            gaps = np.diff(touched) - 1
            cut = 0
            if np.count_nonzero(gaps) + 1 > max_runs:
                cut = int(np.sort(gaps)[::-1][max_runs - 1])     # bridge gaps up to this size
            split = np.nonzero(gaps > cut)[0]
            starts = np.concatenate([[touched[0]], touched[split + 1]]).astype(np.int64)
            ends = np.concatenate([touched[split], [touched[-1]]]).astype(np.int64) + 1
            lengths = ends - starts
            positions = np.concatenate([[0], np.cumsum(lengths)[:-1]])
            n_cover = int(lengths.sum())
            if n_cover <= worthwhile * n_a:
                run = np.searchsorted(starts, self.indices, side='right') - 1
                new_idx = (self.indices - starts[run] + positions[run]).astype(np.int32)
                self._cover = {'runs': [(int(a), int(b), int(c)) for a, b, c in
                                        zip(starts, lengths, positions)],
                               'n_cover': n_cover, 'indices': new_idx}
        return self._cover

    def cover_exact(self, worthwhile=0.7, slack=0.015):
        """The touched source rows (sorted) for transfers that move row by row or run by run
        (pinned host memory): ``{'rows', 'n_cover', 'indices', 'run_start', 'run_len',
        'run_pos'}`` or None.  Short gaps between runs are bridged as long as the untouched rows
        this adds stay below ``slack`` of the touched ones: a DMA submission is cheaper with
        fewer, longer runs (C3: 1004 -> ~880 runs, +0.4 % bytes, 4.8 -> 4.2 ms per slice)."""
        if getattr(self, '_cover_exact', False) is not False:
            return self._cover_exact
        self._cover_exact = None
        touched = self._touched_rows()
        if touched.size and touched.size <= worthwhile * self.shape[1]:
            gaps = np.diff(touched) - 1                       # untouched rows between neighbours
            holes = np.sort(gaps[gaps > 0])
            bridge = 0
            if holes.size:
                ok = np.nonzero(np.cumsum(holes) <= slack * touched.size)[0]
                if ok.size:
                    bridge = int(holes[ok[-1]])
                    # the cumulative bound must hold for every gap of that size
                    if np.sum(holes[holes <= bridge]) > slack * touched.size:
                        smaller = holes[holes < bridge]
                        bridge = int(smaller[-1]) if smaller.size else 0
            split = np.nonzero(gaps > bridge)[0]
            starts = np.concatenate([[touched[0]], touched[split + 1]]).astype(np.int64)
            ends = np.concatenate([touched[split], [touched[-1]]]).astype(np.int64) + 1
            lengths = ends - starts
            positions = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64)
            rows = np.concatenate([np.arange(a, b) for a, b in zip(starts, ends)]) \
                if bridge else touched
            run = np.searchsorted(starts, self.indices, side='right') - 1
            self._cover_exact = {
                'rows': rows.astype(np.int32), 'n_cover': int(rows.size),
                'n_touched': int(touched.size), 'bridged_gap': bridge,
                'indices': (self.indices - starts[run] + positions[run]).astype(np.int32),
                'run_start': starts, 'run_len': lengths, 'run_pos': positions}
        return self._cover_exact

    def on_device_cover(self, device=0, exact=False):
        """``DeviceCSR`` over the covered source rows only (see :meth:`cover`)."""
        from ._cabi import DeviceCSR
        cov = self.cover_exact() if exact else self.cover()
        key = ('cover-exact' if exact else 'cover', int(device))
        if key not in self._device:
            self._device[key] = DeviceCSR(self.indptr, cov['indices'], self.data, self.frac_b,
                                          cov['n_cover'], int(device))
        return self._device[key]

    def release(self):
        for h in self._device.values():
            h.close()
        self._device = {}
