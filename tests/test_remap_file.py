"""File-to-file remap: ``Remapper.remap_file`` / ``ncremap``.

CPU tests cover the file walk (which variables are remapped, copied, dropped; dims,
coordinates, attributes, ``_FillValue`` rule of the reference's writer,
``/root/reference/pyremap/utility.py:35-51``; pre-flight behaviour of
``/root/reference/pyremap/remapper/ncremap.py:15-28``) with the oracle standing in for the
kernels; the GPU test runs the real path and compares it with the oracle bit for bit.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest
from scipy.io import netcdf_file

import pyremap_b200
from oracle import remap_oracle
from pyremap_b200 import engine, remap_file as rf, synthetic as syn

FILL = 9.969209968386869e+36


def _write_input(path, m, with_fill_attr=True, mpas_fill=False):
    """(time=2 record, depth=3, lat, lon) temperature with land NaNs stored as a fill value,
    a clean ssh field, a bounds variable, a lat-only variable and a char variable."""
    rng = np.random.default_rng(11)
    nlat, nlon = m.src_descriptor.dim_sizes
    temp = rng.normal(10.0, 5.0, size=(2, 3, nlat, nlon))
    land = rng.random((nlat, nlon)) < 0.3
    missing = syn_fill = rf.MPAS_FILL if mpas_fill else -1.0e34
    temp_file = np.where(land[None, None], missing, temp)
    temp_nan = np.where(land[None, None], np.nan, temp)
    ssh = rng.normal(0.0, 1.0, size=(2, nlat, nlon)).astype(np.float32)
    with netcdf_file(path, 'w', version=2) as nc:
        nc.createDimension('time', None)
        nc.createDimension('depth', 3)
        nc.createDimension('lat', nlat)
        nc.createDimension('lon', nlon)
        nc.createDimension('nbnd', 2)
        nc.createDimension('strlen', 4)
        nc.title = 'tiny'
        nc.history = 'made by the test'
        v = nc.createVariable('time', 'f8', ('time',))
        v[:] = [10.0, 20.0]
        v.units = 'days'
        v = nc.createVariable('lat', 'f8', ('lat',))
        v[:] = m.src_descriptor.coords['lat']['data']
        v = nc.createVariable('lon', 'f8', ('lon',))
        v[:] = m.src_descriptor.coords['lon']['data']
        v = nc.createVariable('temperature', 'f8', ('time', 'depth', 'lat', 'lon'))
        v[:] = temp_file
        v.units = 'C'
        if with_fill_attr and not mpas_fill:
            v._FillValue = np.float64(syn_fill)
        v = nc.createVariable('ssh', 'f4', ('time', 'lat', 'lon'))
        v[:] = ssh
        v.units = 'm'
        v = nc.createVariable('time_bnds', 'f8', ('time', 'nbnd'))
        v[:] = np.arange(4.0).reshape(2, 2)
        v = nc.createVariable('lat_only', 'f8', ('lat',))
        v[:] = np.arange(float(nlat))
        v = nc.createVariable('xtime', 'c', ('time', 'strlen'))
        v[:] = np.array([list('0001'), list('0002')], dtype='S1')
        v = nc.createVariable('count', 'i4', ('time',))
        v[:] = [3, 4]
    return temp_nan, ssh


def _oracle_many(m):
    A = remap_oracle.build_matrix(m.S, m.row, m.col, m.n_b, m.n_a)

    def many(matrix, dst_dims, fields, threshold=None, *, device=None, kernel=0):
        out = []
        for field, axes in fields:
            arg = field
            nan = np.isnan(field)
            if nan.any():
                arg = np.ma.masked_array(field, nan)
            res = remap_oracle.remap_array(A, m.frac_b, m.dst_grid_dims, arg, axes, threshold)
            out.append(remap_oracle.nanfilled(res))
        return out
    return many


def _remapper(tmp_path, m):
    path = str(tmp_path / 'map.npz')
    m.save_npz(path)
    return pyremap_b200.Remapper(map_filename=path, src_descriptor=m.src_descriptor,
                                 dst_descriptor=m.dst_descriptor)


def _check_output(out_path, m, temp_nan, ssh, thr, expect_vars=None):
    many = _oracle_many(m)
    ref_t, ref_s = many(None, None, [(temp_nan, [2, 3]), (ssh.astype(np.float32), [1, 2])], thr)
    nlat, nlon = m.dst_descriptor.dim_sizes
    with netcdf_file(out_path, 'r', mmap=False) as nc:
        names = set(nc.variables)
        if expect_vars is None:
            assert names == {'time', 'lat', 'lon', 'temperature', 'ssh', 'time_bnds', 'xtime',
                             'count'}          # 'lat_only' dropped, src coords replaced
        else:
            assert names == expect_vars
        assert nc.dimensions['lat'] == nlat and nc.dimensions['lon'] == nlon
        assert nc.dimensions['time'] is None           # still the record dimension
        t = nc.variables['temperature']
        assert t.dimensions == ('time', 'depth', 'lat', 'lon')
        assert t.units == b'C'
        got = np.array(t[...], dtype=np.float64)
        if np.isnan(ref_t).any():
            assert float(t._FillValue) == FILL
            got = np.where(got == FILL, np.nan, got)
        else:
            assert not hasattr(t, '_FillValue')
        np.testing.assert_array_equal(np.isnan(got), np.isnan(ref_t))
        ok = ~np.isnan(ref_t)
        assert np.array_equal(got[ok].view(np.uint64), ref_t[ok].view(np.uint64))
        if 'ssh' in names:
            s = nc.variables['ssh']
            assert s.dimensions == ('time', 'lat', 'lon') and s.data.dtype.itemsize == 8
            gs = np.array(s[...], dtype=np.float64)
            if np.isnan(ref_s).any():
                gs = np.where(gs == FILL, np.nan, gs)
            else:
                assert not hasattr(s, '_FillValue')    # utility.py:47-50: no NaN -> no fill
            ok = ~np.isnan(ref_s)
            np.testing.assert_array_equal(np.isnan(gs), np.isnan(ref_s))
            assert np.array_equal(gs[ok].view(np.uint64), ref_s[ok].view(np.uint64))
        if 'time_bnds' in names:
            np.testing.assert_array_equal(nc.variables['time_bnds'][...],
                                          np.arange(4.0).reshape(2, 2))
        if 'xtime' in names:
            assert nc.variables['xtime'][...].tobytes() == b'00010002'
        np.testing.assert_array_equal(nc.variables['lat'][...],
                                      m.dst_descriptor.coords['lat']['data'])
        assert nc.title == b'tiny'
        assert nc.history.decode().startswith('made by the test\n')
        assert nc.mesh_name.decode() == m.dst_descriptor.mesh_name


@pytest.mark.parametrize('thr', [0.01, None])
def test_remap_file_walk_with_oracle_kernels(tmp_path, monkeypatch, thr):
    m = syn.make_c1(20.0, 10.0)
    r = _remapper(tmp_path, m)
    src = str(tmp_path / 'in.nc')
    temp_nan, ssh = _write_input(src, m)
    monkeypatch.setattr(engine, 'apply_weights_many', _oracle_many(m))
    out_path = str(tmp_path / 'out.nc')
    written = r.remap_file(src, out_path, renormalize=thr)
    assert 'temperature' in written and 'lat_only' not in written
    _check_output(out_path, m, temp_nan, ssh, thr)
    assert not [f for f in os.listdir(tmp_path) if '.tmp' in f]


def test_remap_file_variable_list_overwrite_and_mpas_fill(tmp_path, monkeypatch):
    m = syn.make_c1(20.0, 10.0)
    r = _remapper(tmp_path, m)
    src = str(tmp_path / 'in.nc')
    temp_nan, ssh = _write_input(src, m, mpas_fill=True)
    monkeypatch.setattr(engine, 'apply_weights_many', _oracle_many(m))
    out_path = str(tmp_path / 'out.nc')
    # the reference's name of the call, a variable list, MPAS' undeclared fill value
    r.ncremap(src, out_path, variable_list=['temperature', 'time'], renormalize=0.05,
              replace_mpas_fill=True)
    _check_output(out_path, m, temp_nan, ssh, 0.05,
                  expect_vars={'time', 'lat', 'lon', 'temperature'})
    # an existing output is left alone unless overwrite is set (ncremap.py:18-19)
    stamp = os.path.getmtime(out_path)
    assert r.remap_file(src, out_path, renormalize=0.05) is None
    assert os.path.getmtime(out_path) == stamp
    assert r.remap_file(src, out_path, variable_list=['temperature'], renormalize=0.05,
                        replace_mpas_fill=True, overwrite=True) is not None
    # without replace_mpas_fill the fill value is data: finite numbers, no _FillValue
    r.remap_file(src, out_path, variable_list=['temperature'], overwrite=True)
    with netcdf_file(out_path, 'r', mmap=False) as nc:
        assert not hasattr(nc.variables['temperature'], '_FillValue')
        assert np.nanmin(nc.variables['temperature'][...]) < -1e30


def test_remap_file_classic_format_scalar_and_nonadjacent_dims(tmp_path, monkeypatch):
    """NetCDF-3 classic input (version 1), a scalar variable, an int variable on the source grid
    and a variable whose source dims are not adjacent, ``(lat, depth, lon)``: the destination
    dims take the place of the first source dim (remap_numpy.py:171-182)."""
    m = syn.make_c1(20.0, 10.0)
    r = _remapper(tmp_path, m)
    nlat, nlon = m.src_descriptor.dim_sizes
    rng = np.random.default_rng(3)
    odd = rng.normal(size=(nlat, 3, nlon))
    mask = rng.integers(0, 2, size=(nlat, nlon)).astype(np.int32)
    src = str(tmp_path / 'in.nc')
    with netcdf_file(src, 'w', version=1) as nc:
        nc.createDimension('lat', nlat)
        nc.createDimension('depth', 3)
        nc.createDimension('lon', nlon)
        v = nc.createVariable('odd', 'f8', ('lat', 'depth', 'lon'))
        v[:] = odd
        v = nc.createVariable('landmask', 'i4', ('lat', 'lon'))
        v[:] = mask
        v = nc.createVariable('year', 'i4', ())
        v.data[...] = 2001
    monkeypatch.setattr(engine, 'apply_weights_many', _oracle_many(m))
    out_path = str(tmp_path / 'out.nc')
    written = r.remap_file(src, out_path)
    assert set(written) >= {'odd', 'landmask', 'year'}
    many = _oracle_many(m)
    ref_odd, ref_mask = many(None, None, [(odd, [0, 2]), (mask.astype(np.float64), [0, 1])], None)
    with netcdf_file(out_path, 'r', mmap=False) as nc:
        v = nc.variables['odd']
        assert v.dimensions == ('lat', 'lon', 'depth')
        got = np.array(v[...], dtype=np.float64)
        assert got.shape == ref_odd.shape
        ok = ~np.isnan(ref_odd)
        assert np.array_equal(got[ok].view(np.uint64), ref_odd[ok].view(np.uint64))
        lm = nc.variables['landmask']
        assert lm.dimensions == ('lat', 'lon') and lm.data.dtype.kind == 'f'     # float64 result
        np.testing.assert_array_equal(np.array(lm[...], dtype=np.float64)[~np.isnan(ref_mask)],
                                      ref_mask[~np.isnan(ref_mask)])
        assert int(nc.variables['year'].getValue()) == 2001
        assert 'time' not in nc.dimensions


def test_remap_file_preflight_errors(tmp_path):
    m = syn.make_c1(20.0, 10.0)
    src = str(tmp_path / 'in.nc')
    _write_input(src, m)
    r = pyremap_b200.Remapper(src_descriptor=m.src_descriptor, dst_descriptor=m.dst_descriptor)
    with pytest.raises(ValueError, match='No mapping file has been defined'):
        r.remap_file(src, str(tmp_path / 'o.nc'))
    r = _remapper(tmp_path, m)

    class PointCollectionDescriptor(syn.SimpleDescriptor):
        pass
    r.src_descriptor = PointCollectionDescriptor(m.src_descriptor.dims,
                                                 m.src_descriptor.dim_sizes)
    with pytest.raises(TypeError, match='point collection'):
        r.remap_file(src, str(tmp_path / 'o.nc'))
    r = _remapper(tmp_path, m)
    with pytest.raises(KeyError, match='nope'):
        r.remap_file(src, str(tmp_path / 'o.nc'), variable_list=['nope'])
    bad = syn.make_c1(30.0, 10.0)
    bad_src = str(tmp_path / 'bad.nc')
    _write_input(bad_src, bad)
    with pytest.raises(ValueError, match="don't have the same size"):
        r.remap_file(bad_src, str(tmp_path / 'o.nc'))
    not_nc = tmp_path / 'x.nc'
    not_nc.write_bytes(b'\x89HDF\r\n\x1a\n' + b'\0' * 64)
    try:
        import netCDF4  # noqa: F401
    except ImportError:
        with pytest.raises(OSError, match='not a NetCDF-3 file'):
            r.remap_file(str(not_nc), str(tmp_path / 'o.nc'))
    assert not os.path.exists(tmp_path / 'o.nc')


def test_fill_value_rule_and_missing_decoding():
    a = np.array([1.0, np.nan])
    assert rf.fill_value_for(a) == FILL
    assert rf.fill_value_for(a.astype(np.float32)) == FILL
    assert rf.fill_value_for(np.array([1.0, 2.0])) is None
    assert rf.fill_value_for(np.array([1, 2])) is None
    assert rf.fill_value_for(np.array([b'a'])) is None
    assert rf.fill_value_for(a, {'f8': -9.0}) == -9.0
    big = np.array([1.0, -1e34, rf.MPAS_FILL], dtype='>f8')
    out = rf.decode_missing(big, {'_FillValue': np.array([-1e34])})
    assert out.dtype.isnative and np.isnan(out[1]) and out[2] == rf.MPAS_FILL
    out = rf.decode_missing(big, {}, replace_mpas_fill=True)
    assert np.isnan(out[2]) and out[1] == -1e34
    ints = rf.decode_missing(np.array([1, 2], dtype='>i4'), {'_FillValue': 1})
    assert ints.dtype == np.int32 and ints.tolist() == [1, 2]
    assert rf._to_nc3(np.array([1, 2], dtype=np.int64)).dtype == np.int32
    with pytest.raises(ValueError):
        rf._to_nc3(np.array([2 ** 40]))


@pytest.mark.gpu
def test_remap_file_on_the_gpu_is_bitwise_the_oracle(tmp_path):
    m = syn.make_c1(20.0, 10.0)
    r = _remapper(tmp_path, m)
    src = str(tmp_path / 'in.nc')
    temp_nan, ssh = _write_input(src, m)
    for thr in (0.01, None):
        out_path = str(tmp_path / f'out_{thr}.nc')
        r.remap_file(src, out_path, renormalize=thr)
        _check_output(out_path, m, temp_nan, ssh, thr)
