"""Host-side logic that needs no GPU: CSR construction, layout planning, the
Remapper facade's error behaviour, map readers, the C ABI's export table."""

import ctypes
import os
import re

import numpy as np
import pytest
from scipy.sparse import csr_matrix

import pyremap_b200
from pyremap_b200 import _cabi, build, mapfile, synthetic as syn
from pyremap_b200.engine import Layout
from pyremap_b200.remap_numpy import _check_drop, _load_mapping

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- CSR builder == scipy's canonical CSR --------------------------------
@pytest.mark.parametrize('seed', range(5))
def test_coo_to_csr_equals_scipy(seed):
    rng = np.random.default_rng(seed)
    n_row, n_col, n = 50, 40, 600
    row = rng.integers(0, n_row, n)
    col = rng.integers(0, n_col, n)          # plenty of duplicates
    S = rng.normal(size=n)
    # at most two duplicates per (row, col): order-independent sums
    key = row * n_col + col
    _, first, counts = np.unique(key, return_index=True, return_counts=True)
    keep = np.ones(n, bool)
    for k in np.unique(key)[counts > 2]:
        idx = np.nonzero(key == k)[0]
        keep[idx[2:]] = False
    row, col, S = row[keep], col[keep], S[keep]
    ref = csr_matrix((S, (row, col)), shape=(n_row, n_col))
    indptr, indices, data = mapfile.coo_to_csr(S, row, col, n_row, n_col)
    assert indptr.dtype == np.int32 and indices.dtype == np.int32
    np.testing.assert_array_equal(indptr, ref.indptr)
    np.testing.assert_array_equal(indices, ref.indices)
    assert np.array_equal(data.view(np.uint64), ref.data.view(np.uint64))


def test_coo_to_csr_fast_path_and_errors():
    m = syn.make_c1(20.0, 10.0)
    ip, ix, d = mapfile.coo_to_csr(m.S, m.row - 1, m.col - 1, m.n_b, m.n_a)
    ref = csr_matrix((m.S, (m.row - 1, m.col - 1)), shape=(m.n_b, m.n_a))
    np.testing.assert_array_equal(ix, ref.indices)
    with pytest.raises(ValueError):
        mapfile.coo_to_csr([1.0], [5], [0], 3, 3)
    with pytest.raises(ValueError):
        mapfile.coo_to_csr([1.0], [0], [-1], 3, 3)
    ip, ix, d = mapfile.coo_to_csr([], [], [], 3, 3)
    assert ip.tolist() == [0, 0, 0, 0] and ix.size == 0


# ---- layout planning -------------------------------------------------------
def test_layout_adjacent_batched():
    lay = Layout((12, 235, 60), [1], [36, 72])
    assert (lay.adjacent, lay.B, lay.n_src, lay.L, lay.K) == (True, 12, 235, 60, 720)
    assert lay.out_shape == (12, 36, 72, 60)
    lay = Layout((3, 51, 61, 2), [1, 2], [7, 6])
    assert (lay.B, lay.n_src, lay.L) == (3, 51 * 61, 2) and lay.out_shape == (3, 7, 6, 2)
    lay = Layout((51, 61), [0, 1], [9])
    assert (lay.B, lay.L, lay.K) == (1, 1, 1) and lay.out_shape == (9,)


def test_layout_nonadjacent_matches_reference_permutation():
    shape = (9, 4, 18, 3)
    lay = Layout(shape, [0, 2], [5, 6])
    assert not lay.adjacent and lay.K == 12
    flat = np.arange(5 * 6 * 12).reshape(5 * 6, 12)
    mine = flat.reshape(lay.dst_dims + lay.extra_shape).transpose(lay.unpermute_axes())
    assert mine.shape == lay.out_shape == (5, 6, 4, 3)


def test_layout_rejects_bad_axes():
    with pytest.raises(ValueError):
        Layout((3, 4), [2], [5])
    with pytest.raises(ValueError):
        Layout((3, 4), [], [5])


# ---- Remapper facade -----------------------------------------------------
def test_remapper_signature_matches_reference():
    import inspect
    sig = inspect.signature(pyremap_b200.Remapper.__init__)
    assert list(sig.parameters)[1:] == ['ntasks', 'map_filename', 'method', 'src_descriptor',
                                        'dst_descriptor', 'map_tool', 'parallel_exec', 'use_tmp']
    r = pyremap_b200.Remapper()
    assert (r.ntasks, r.map_filename, r.method, r.map_tool, r.parallel_exec, r.use_tmp) == \
        (1, None, 'bilinear', 'esmf', 'mpirun', True)
    assert r._ds_map is None and r._matrix is None
    assert pyremap_b200.Remapper.remap is pyremap_b200.Remapper.remap_numpy
    sig = inspect.signature(pyremap_b200.Remapper.remap_numpy)
    assert list(sig.parameters) == ['self', 'ds', 'renormalization_threshold']
    assert sig.parameters['renormalization_threshold'].default is None


def test_remapper_errors_before_any_arithmetic(tmp_path):
    import xarray as xr
    m = syn.make_c1(20.0, 10.0)
    path = str(tmp_path / 'map.npz')
    m.save_npz(path)
    da = xr.DataArray(np.zeros((9, 18)), dims=('lat', 'lon'))

    r = pyremap_b200.Remapper(src_descriptor=m.src_descriptor, dst_descriptor=m.dst_descriptor)
    with pytest.raises(ValueError, match='No mapping file has been defined'):
        r.remap_numpy(da)

    r = pyremap_b200.Remapper(map_filename=path, src_descriptor=m.src_descriptor,
                              dst_descriptor=m.dst_descriptor)
    with pytest.raises(TypeError, match='ds not an xarray Dataset or DataArray'):
        class Sized:
            sizes = {'lat': 9, 'lon': 18}
        r.remap_numpy(Sized())
    bad = xr.DataArray(np.zeros((9, 17)), dims=('lat', 'lon'))
    with pytest.raises(ValueError, match="don't have the same size"):
        r.remap_numpy(bad)

    wrong_rank = syn.SimpleDescriptor(['nCells'], [m.n_a])
    r2 = pyremap_b200.Remapper(map_filename=path, src_descriptor=wrong_rank,
                               dst_descriptor=m.dst_descriptor)
    with pytest.raises(ValueError, match='number of source and/or destination dimensions'):
        r2.remap_numpy(da)
    wrong_size = syn.SimpleDescriptor(['lat', 'lon'], [9, 19])
    r3 = pyremap_b200.Remapper(map_filename=path, src_descriptor=wrong_size,
                               dst_descriptor=m.dst_descriptor)
    with pytest.raises(ValueError, match='source mesh descriptor and remapping source'):
        r3.remap_numpy(da)
    wrong_dst = syn.SimpleDescriptor(['lat', 'lon'], [18, 35])
    r4 = pyremap_b200.Remapper(map_filename=path, src_descriptor=m.src_descriptor,
                               dst_descriptor=wrong_dst)
    with pytest.raises(ValueError, match='dest. mesh descriptor and remapping dest.'):
        r4.remap_numpy(da)
    with pytest.raises(NotImplementedError):
        r.build_map()


def test_load_mapping_builds_canonical_csr_once(tmp_path):
    m = syn.make_c1(20.0, 10.0, shuffle_triplets=True)
    path = str(tmp_path / 'map.npz')
    m.save_npz(path)
    r = pyremap_b200.Remapper(map_filename=path, src_descriptor=m.src_descriptor,
                              dst_descriptor=m.dst_descriptor)
    _load_mapping(r)
    first = r._matrix
    _load_mapping(r)
    assert r._matrix is first                      # memoised (remap_numpy.py:81-82)
    ref = csr_matrix((m.S, (m.row - 1, m.col - 1)), shape=(m.n_b, m.n_a))
    np.testing.assert_array_equal(first.indices, ref.indices)
    assert np.array_equal(first.data.view(np.uint64), ref.data.view(np.uint64))
    assert first.shape == (m.n_b, m.n_a) and first.nnz == ref.nnz
    np.testing.assert_array_equal(r._ds_map['dst_grid_dims'].values, m.dst_grid_dims)
    np.testing.assert_array_equal(r._ds_map['frac_b'].values, m.frac_b)


def test_check_drop():
    import xarray as xr

    class R:
        src_descriptor = syn.SimpleDescriptor(['lat', 'lon'], [2, 3])
    assert not _check_drop(R, xr.DataArray(np.zeros((2, 3)), dims=('lat', 'lon')))
    assert not _check_drop(R, xr.DataArray(np.zeros((4,)), dims=('time',)))
    assert _check_drop(R, xr.DataArray(np.zeros((2,)), dims=('lat',)))


def test_netcdf3_map_reader(tmp_path):
    from scipy.io import netcdf_file
    m = syn.make_c1(30.0, 15.0)
    path = str(tmp_path / 'map.nc')
    with netcdf_file(path, 'w') as nc:
        for dim, n in (('n_a', m.n_a), ('n_b', m.n_b), ('n_s', m.n_s),
                       ('src_grid_rank', 2), ('dst_grid_rank', 2)):
            nc.createDimension(dim, n)
        for name, dim, typ, val in (('S', 'n_s', 'd', m.S), ('row', 'n_s', 'i', m.row),
                                    ('col', 'n_s', 'i', m.col), ('frac_b', 'n_b', 'd', m.frac_b),
                                    ('frac_a', 'n_a', 'd', np.ones(m.n_a)),
                                    ('src_grid_dims', 'src_grid_rank', 'i', m.src_grid_dims),
                                    ('dst_grid_dims', 'dst_grid_rank', 'i', m.dst_grid_dims)):
            v = nc.createVariable(name, typ, (dim,))
            v[:] = val
    ds = mapfile.open_map(path)
    assert ds.sizes['n_a'] == m.n_a and ds.sizes['n_b'] == m.n_b
    np.testing.assert_array_equal(ds['row'].values, m.row)
    np.testing.assert_array_equal(ds['S'].values, m.S)


# ---- the C ABI -----------------------------------------------------------
def test_library_builds_loads_and_exports_every_declared_symbol():
    path = build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, 'include', 'b200remap.h')).read()
    declared = set(re.findall(r'B200REMAP_API[^;(]*?\b(b200remap_\w+)\s*\(', header))
    assert declared == set(_cabi.EXPORTED_SYMBOLS), declared ^ set(_cabi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f'{name} not exported'
    assert _cabi.load_library().b200remap_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(pyremap_b200.B200RemapError, match='no CPU fallback'):
        _cabi.DeviceCSR(np.array([0, 1], np.int32), np.array([0], np.int32),
                        np.array([1.0]), None, 1)
    m = syn.make_c1(30.0, 15.0)
    r = pyremap_b200.Remapper(map_filename='unused.npz', src_descriptor=m.src_descriptor,
                              dst_descriptor=m.dst_descriptor)
    ip, ix, d = mapfile.coo_to_csr(m.S, m.row - 1, m.col - 1, m.n_b, m.n_a)
    r._matrix = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
    r._ds_map = mapfile.MapDataset({'dst_grid_dims': m.dst_grid_dims, 'frac_b': m.frac_b}, {})
    with pytest.raises(pyremap_b200.B200RemapError, match='no CPU fallback'):
        r.remap_array(np.zeros((6, 12)), [0, 1])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'pyremap_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), f
                assert 'scipy.sparse' not in text or f == 'mapfile.py', f


def test_host_any_nan_native_cpu():
    rng = np.random.default_rng(0)
    for dtype in (np.float64, np.float32):
        a = rng.normal(size=300_001).astype(dtype)
        assert _cabi.host_any_nan(a) is False
        for pos in (0, 150_000, 300_000):
            b = a.copy()
            b[pos] = np.nan
            assert _cabi.host_any_nan(b) is True and _cabi.host_any_nan(b, threads=1) is True
    assert _cabi.host_any_nan(np.zeros(0)) is False
    with pytest.raises(ValueError):
        _cabi.host_any_nan(np.zeros(4, dtype=np.int32))


def test_cover_renumbering_is_monotonic_and_complete():
    m = syn.make_c3(scale=0.02)
    ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1,
                                   m.n_b, m.n_a)
    W = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
    cov = W.cover()
    assert cov is not None and len(cov['runs']) <= 64
    lut = np.full(m.n_a, -1)
    for start, length, pos in cov['runs']:
        lut[start:start + length] = np.arange(pos, pos + length)
    assert np.array_equal(lut[ix], cov['indices']) and cov['indices'].min() >= 0
    assert cov['n_cover'] == sum(r[1] for r in cov['runs']) < 0.7 * m.n_a
    full = syn.make_c1(30.0, 15.0)
    ip, ix, d = mapfile.coo_to_csr(full.S, full.row - 1, full.col - 1, full.n_b, full.n_a)
    assert mapfile.WeightMatrix(ip, ix, d, (full.n_b, full.n_a), full.frac_b).cover() is None


@pytest.mark.parametrize('slack', [0.0, 0.015, 0.2])
def test_cover_exact_runs_bridge_short_gaps_only(slack):
    """``cover_exact``: every referenced source row keeps its data under the renumbering, the
    runs tile the covered rows in order, and bridging adds at most ``slack`` untouched rows."""
    from pyremap_b200 import mapfile
    rng = np.random.default_rng(5)
    n_a, n_b = 5000, 300
    # touched rows: a few dense bands with small holes, the rest of the mesh untouched
    touched = np.unique(np.concatenate([np.arange(100, 400), np.arange(402, 600),
                                        np.arange(640, 900), rng.integers(2000, 2300, 200)]))
    rows = np.repeat(np.arange(n_b), 4)
    cols = rng.choice(touched, size=rows.size)
    cols[:touched.size] = touched                      # every touched row is referenced
    ip, ix, d = mapfile.coo_to_csr(rng.normal(size=rows.size), rows, cols, n_b, n_a)
    W = mapfile.WeightMatrix(ip, ix, d, (n_b, n_a), np.ones(n_b))
    cov = W.cover_exact(slack=slack)
    assert cov is not None and cov['n_touched'] == touched.size
    np.testing.assert_array_equal(cov['rows'][cov['indices']], ix)       # renumbering is faithful
    np.testing.assert_array_equal(np.unique(cov['rows']), cov['rows'])    # sorted, no repeats
    assert cov['n_cover'] - cov['n_touched'] <= slack * cov['n_touched']
    starts, lens, pos = cov['run_start'], cov['run_len'], cov['run_pos']
    assert np.all(lens > 0) and np.all(starts[1:] > starts[:-1] + lens[:-1])   # disjoint, ordered
    np.testing.assert_array_equal(pos, np.concatenate([[0], np.cumsum(lens)[:-1]]))
    np.testing.assert_array_equal(np.concatenate([np.arange(s, s + n) for s, n in zip(starts, lens)]),
                                  cov['rows'])
    if slack == 0.0:
        assert cov['bridged_gap'] == 0 and cov['n_cover'] == touched.size
    # a map that touches (nearly) everything has no worthwhile cover
    ip2, ix2, d2 = mapfile.coo_to_csr(np.ones(n_a), np.arange(n_a) % n_b, np.arange(n_a), n_b, n_a)
    assert mapfile.WeightMatrix(ip2, ix2, d2, (n_b, n_a), np.ones(n_b)).cover_exact() is None


def test_engine_host_helpers():
    from pyremap_b200 import engine
    assert engine._wants_f32(None) is False and engine._wants_f32(np.float64) is False
    assert engine._wants_f32(np.float32) is True and engine._wants_f32('float32') is True
    with pytest.raises(ValueError, match='out_dtype'):
        engine._wants_f32(np.int16)
    off, ln = engine._chunks(2 * (1 << 20) + 5)
    np.testing.assert_array_equal(off, [0, 1 << 20, 2 << 20])
    np.testing.assert_array_equal(ln, [1 << 20, 1 << 20, 5])
    off, ln = engine._chunks(7, piece=4)
    assert off.tolist() == [0, 4] and ln.tolist() == [4, 3]
