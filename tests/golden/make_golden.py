"""Generate the golden fixtures in this directory by running the REAL reference.

Run in the authoring container only (needs ``/root/reference``)::

    python tests/golden/make_golden.py

For every case the inputs (map triplets, field, mask, axes, threshold) are fed to
the reference's own ``_remap_numpy_array`` (``/root/reference/pyremap/remapper/
remap_numpy.py:223``), executed from its own file by ``oracle/ref_loader.py``; the
dataset-level cases run the reference's ``_remap_numpy`` (``:19``) end to end on
``tests/minixarray.py`` containers.  Inputs and reference outputs are stored
together so the fixtures are self-contained: the GPU box has no reference tree.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import minixarray  # noqa: E402
from oracle import ref_loader  # noqa: E402
from pyremap_b200 import synthetic as syn  # noqa: E402


def _map_dict(m):
    return dict(S=m.S, row=m.row, col=m.col, frac_b=m.frac_b,
                src_grid_dims=m.src_grid_dims, dst_grid_dims=m.dst_grid_dims,
                n_a=np.int64(m.n_a), n_b=np.int64(m.n_b))


def _custom_map(n_a, n_b, rows, cols, vals, frac_b, src_dims, dst_dims):
    return dict(S=np.asarray(vals, np.float64),
                row=np.asarray(rows, np.int32) + 1,
                col=np.asarray(cols, np.int32) + 1,
                frac_b=np.asarray(frac_b, np.float64),
                src_grid_dims=np.asarray(src_dims, np.int32),
                dst_grid_dims=np.asarray(dst_dims, np.int32),
                n_a=np.int64(n_a), n_b=np.int64(n_b))


def array_cases():
    rng = np.random.default_rng(20261017)
    cases = []

    # -- C1-like bilinear, 10 levels first, unmasked, no threshold
    m = syn.make_c1(src_res=20.0, dst_res=10.0, shuffle_triplets=True)
    f = rng.normal(size=(10, 9, 18))
    cases.append(dict(name='c1_bilinear_lev10_unmasked', map=_map_dict(m),
                      field=f, remap_axes=[1, 2], thr=None))

    # -- SST-like: float32 (time=1, lat, lon), NaN over "land", thr 0.01 (K=1)
    f = rng.uniform(-2, 30, size=(1, 9, 18)).astype(np.float32)
    f[0, 2:5, 3:9] = np.nan
    f[0, 7, :] = np.nan
    cases.append(dict(name='sst_like_f32_masked_k1', map=_map_dict(m), field=f,
                      remap_axes=[1, 2], thr=0.01, wrap_nan=True))

    # -- (time=3, lat, lon): source dims last, B=3
    f = rng.normal(size=(3, 9, 18))
    f[1, 4, 4:12] = np.nan
    cases.append(dict(name='time_lat_lon_masked_b3', map=_map_dict(m), field=f,
                      remap_axes=[1, 2], thr=0.3, wrap_nan=True))

    # -- non-adjacent remap axes (lat, time, lon)
    f = rng.normal(size=(9, 4, 18))
    cases.append(dict(name='nonadjacent_axes_lat_time_lon', map=_map_dict(m),
                      field=f, remap_axes=[0, 2], thr=None))

    # -- C2-like: (Time, nCells, nVertLevels) with bathymetry NaNs, thr 0.01
    m2 = syn.make_c2(scale=0.002)
    lv = syn.bathymetry_levels(m2.n_a, 6, seed=7)
    f = np.stack([syn.ocean_field(m2.n_a, 6, seed=8 + t, max_level=lv)
                  for t in range(2)])
    cases.append(dict(name='c2_time_cells_levels_masked', map=_map_dict(m2),
                      field=f, remap_axes=[1], thr=0.01, wrap_nan=True))

    # -- same map: (Time=1, nCells) float64 without NaN but with a threshold -> frac_b branch
    f = rng.normal(size=(1, m2.n_a))
    cases.append(dict(name='ssh_like_no_nan_with_thr', map=_map_dict(m2),
                      field=f, remap_axes=[1], thr=0.01, wrap_nan=True))

    # -- NaN in the field but no threshold: NaN propagates through the frac_b branch
    f = rng.normal(size=(m2.n_a, 3))
    f[::17, 1] = np.nan
    cases.append(dict(name='nan_no_threshold_propagates', map=_map_dict(m2),
                      field=f, remap_axes=[0], thr=None, wrap_nan=True))

    # -- C3-like, 8 levels, unmasked and masked
    m3 = syn.make_c3(scale=0.001)
    lv = syn.bathymetry_levels(m3.n_a, 8, seed=11)
    f = syn.ocean_field(m3.n_a, 8, seed=12)
    cases.append(dict(name='c3_cells_levels_unmasked', map=_map_dict(m3),
                      field=f, remap_axes=[0], thr=None))
    f = syn.ocean_field(m3.n_a, 8, seed=13, max_level=lv)
    cases.append(dict(name='c3_cells_levels_masked', map=_map_dict(m3), field=f,
                      remap_axes=[0], thr=0.01, wrap_nan=True))

    # -- C4-like (long rows): (dim0, y, x, dim3) extra dims on both sides
    m4 = syn.make_c4(scale=0.01, ratio=4)
    ny, nx = m4.src_descriptor.dim_sizes
    f = rng.normal(size=(3, ny, nx, 2))
    f[:, 5:9, 3:12, :] = np.nan
    cases.append(dict(name='c4_extra_dims_both_sides_masked', map=_map_dict(m4),
                      field=f, remap_axes=[1, 2], thr=0.01, wrap_nan=True))
    # 2-D field, K = 1, long rows, NaN disc
    f = rng.normal(size=(ny, nx))
    yy, xx = np.mgrid[0:ny, 0:nx]
    f[(yy - ny / 2) ** 2 + (xx - nx / 2) ** 2 < 30] = np.nan
    cases.append(dict(name='c4_2d_k1_masked', map=_map_dict(m4), field=f,
                      remap_axes=[0, 1], thr=0.01, wrap_nan=True))

    # -- hand-made map: duplicates, shuffled triplets, negative weights,
    #    frac_b zero / negative / NaN, an empty row, +-inf in the field
    rows = [0, 0, 0, 1, 1, 2, 2, 2, 4, 4, 5, 0, 2]
    cols = [3, 1, 3, 0, 2, 4, 4, 1, 5, 0, 2, 1, 4]
    vals = [0.25, 0.5, 0.125, -0.75, 1.5, 0.1, 0.2, 0.7, 0.3, 0.3, 1.0, 0.0625, 0.3]
    frac = [1.0, 0.75, 0.0, 1.0, -0.5, np.nan]
    hm = _custom_map(6, 6, rows, cols, vals, frac, [6], [3, 2])
    f = rng.normal(size=(6, 5))
    f[2, 0] = np.inf
    f[4, 1] = -np.inf
    cases.append(dict(name='handmade_duplicates_fracb_edges', map=hm, field=f,
                      remap_axes=[0], thr=None))
    f2 = f.copy()
    f2[1, 2] = np.nan
    f2[3, :] = np.nan
    cases.append(dict(name='handmade_duplicates_masked', map=hm, field=f2,
                      remap_axes=[0], thr=0.2, wrap_nan=True))

    # -- threshold is strict: den == thr exactly must be masked
    hm2 = _custom_map(4, 3, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 3, 0],
                      [0.25, 0.25, 0.5, 0.5, 0.5, 1.0], [1, 1, 1], [4], [3])
    f = np.array([[1.0, 2.0], [np.nan, 3.0], [np.nan, 4.0], [5.0, np.nan]])
    cases.append(dict(name='threshold_is_strict', map=hm2, field=f,
                      remap_axes=[0], thr=0.5, wrap_nan=True))
    cases.append(dict(name='threshold_zero', map=hm2, field=f, remap_axes=[0],
                      thr=0.0, wrap_nan=True))

    # -- integer field (upcast to float64 through the weights)
    f = rng.integers(-50, 50, size=(m2.n_a, 4)).astype(np.int32)
    cases.append(dict(name='int32_field', map=_map_dict(m2), field=f,
                      remap_axes=[0], thr=None))

    # -- MaskedArray whose mask is NOT isnan(data): finite junk under the mask,
    #    and one unmasked NaN (which then poisons its destination rows)
    f = rng.normal(size=(m2.n_a, 3))
    mask = rng.random(f.shape) < 0.3
    f[5, 0] = np.nan
    mask[5, 0] = False
    cases.append(dict(name='explicit_mask_not_isnan', map=_map_dict(m2), field=f,
                      mask=mask, remap_axes=[0], thr=0.05))
    # MaskedArray given but threshold None -> unmasked branch on the raw data
    cases.append(dict(name='masked_array_no_threshold', map=_map_dict(m2),
                      field=f, mask=mask, remap_axes=[0], thr=None))
    return cases


def run_array_case(case):
    from oracle.remap_oracle import build_matrix
    mp = case['map']
    matrix = build_matrix(mp['S'], mp['row'], mp['col'], mp['n_b'], mp['n_a'])
    field = case['field']
    if 'mask' in case:
        arg = np.ma.masked_array(field, mask=case['mask'])
    elif case.get('wrap_nan'):
        # what _remap_data_array does (remap_numpy.py:201-204)
        nanmask = np.isnan(field)
        arg = np.ma.masked_array(field, nanmask) if nanmask.any() else field
    else:
        arg = field
    out = ref_loader.reference_remap_array(matrix, mp['frac_b'],
                                           mp['dst_grid_dims'], arg,
                                           case['remap_axes'], case['thr'])
    assert isinstance(out, np.ma.MaskedArray) and out.dtype == np.float64
    return np.ma.getdata(out), np.ma.getmaskarray(out)


def dataset_case(tmpdir):
    """Reference `_remap_numpy` end to end on a small Dataset."""
    ref = ref_loader.load(minixarray)
    xr = minixarray
    m = syn.make_c1(src_res=20.0, dst_res=10.0)
    map_file = os.path.join(tmpdir, 'map_ds_case.npz')
    m.save_npz(map_file)
    rng = np.random.default_rng(99)
    temp = rng.uniform(-2, 30, size=(2, 3, 9, 18))
    temp[:, :, 3:5, 2:8] = np.nan
    ssh = rng.normal(size=(2, 9, 18)).astype(np.float32)
    ds = xr.Dataset(
        {'temperature': (('time', 'depth', 'lat', 'lon'), temp, {'units': 'C'}),
         'ssh': (('time', 'lat', 'lon'), ssh, {'units': 'm'}),
         'time_bnds': (('time', 'nbnd'), np.arange(4.0).reshape(2, 2)),
         'lat_only': (('lat',), np.arange(9.0)),
         'xtime': (('time', 'strlen'), np.zeros((2, 4), dtype='S1'))},
        coords={'time': np.array([10.0, 20.0]), 'depth': np.array([5., 15., 25.]),
                'lat': m.src_descriptor.coords['lat']['data'],
                'lon': m.src_descriptor.coords['lon']['data']},
        attrs={'history': 'created by make_golden', 'title': 'tiny'})

    class R:
        pass
    r = R()
    r.map_filename = map_file
    r.src_descriptor = m.src_descriptor
    r.dst_descriptor = m.dst_descriptor
    r._ds_map = None
    r._matrix = None
    saved = sys.argv[:]
    sys.argv = ['golden_prog', '--flag']
    try:
        out = ref._remap_numpy(r, ds, 0.01)
    finally:
        sys.argv = saved
    arrays = dict(temperature_in=temp, ssh_in=ssh)
    meta = {'data_vars': {}, 'attrs': out.attrs, 'coords': sorted(out.coords),
            'argv': ['golden_prog', '--flag']}
    for name, var in out.data_vars.items():
        meta['data_vars'][name] = {'dims': list(var.dims),
                                   'attrs': {k: str(v) for k, v in var.attrs.items()},
                                   'dtype': str(var.dtype)}
        if var.dtype.kind == 'f':
            arrays[f'out__{name}'] = var.values
    np.savez_compressed(os.path.join(HERE, 'dataset_case.npz'), **arrays)
    with open(os.path.join(HERE, 'dataset_case.json'), 'w') as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    os.remove(map_file)
    return meta


def main():
    if not ref_loader.available():
        raise SystemExit('the reference tree is not present; fixtures unchanged')
    ref_loader.load(minixarray)
    index = []
    for case in array_cases():
        data, mask = run_array_case(case)
        payload = {f'map__{k}': v for k, v in case['map'].items()}
        payload['field'] = case['field']
        if 'mask' in case:
            payload['field_mask'] = case['mask']
        payload['remap_axes'] = np.asarray(case['remap_axes'], np.int64)
        payload['thr'] = np.float64(np.nan if case['thr'] is None else case['thr'])
        payload['has_thr'] = np.bool_(case['thr'] is not None)
        payload['wrap_nan'] = np.bool_(case.get('wrap_nan', False))
        payload['out_data'] = data
        payload['out_mask'] = mask
        np.savez_compressed(os.path.join(HERE, f"{case['name']}.npz"), **payload)
        index.append(case['name'])
        print(f"{case['name']:40s} out{data.shape} masked={int(mask.sum())}/{mask.size}")
    meta = dataset_case(HERE)
    print('dataset_case', json.dumps(meta['data_vars'])[:200])
    with open(os.path.join(HERE, 'INDEX.json'), 'w') as fh:
        json.dump({'array_cases': index, 'dataset_cases': ['dataset_case'],
                   'reference': 'pyremap 2.4.0 remap_numpy.py executed in place; '
                                'numpy %s' % np.__version__}, fh, indent=1)


if __name__ == '__main__':
    main()
