"""Shared helpers of the test-suite (golden loading, exact comparison)."""

from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_index():
    with open(os.path.join(GOLDEN, 'INDEX.json')) as fh:
        return json.load(fh)


ARRAY_CASES = golden_index()['array_cases']


def load_case(name):
    with np.load(os.path.join(GOLDEN, f'{name}.npz')) as z:
        d = {k: z[k] for k in z.files}
    case = {
        'name': name,
        'map': {k[5:]: v for k, v in d.items() if k.startswith('map__')},
        'field': d['field'],
        'mask': d.get('field_mask'),
        'remap_axes': [int(a) for a in d['remap_axes']],
        'thr': float(d['thr']) if bool(d['has_thr']) else None,
        'wrap_nan': bool(d['wrap_nan']),
        'out_data': d['out_data'],
        'out_mask': d['out_mask'],
    }
    return case


def reference_argument(case):
    """The object the reference's ``_remap_numpy_array`` was given."""
    f = case['field']
    if case['mask'] is not None:
        return np.ma.masked_array(f, mask=case['mask'])
    if case['wrap_nan']:
        nanmask = np.isnan(f)
        return np.ma.masked_array(f, nanmask) if nanmask.any() else f
    return f


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bitwise(values, keep, ref_values, ref_keep, what=''):
    """Mask placement identical; kept values bit-identical (NaNs: same places)."""
    values = np.asarray(values)
    assert values.shape == ref_values.shape, (what, values.shape, ref_values.shape)
    assert values.dtype == np.float64, (what, values.dtype)
    np.testing.assert_array_equal(np.asarray(keep, bool), np.asarray(ref_keep, bool),
                                  err_msg=f'{what}: mask placement differs')
    k = np.asarray(ref_keep, bool)
    a, b = values[k], ref_values[k]
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    np.testing.assert_array_equal(nan_a, nan_b, err_msg=f'{what}: NaN placement differs')
    ok = ~nan_a
    diff = bits(a[ok]) != bits(b[ok])
    assert not diff.any(), (f'{what}: {int(diff.sum())} of {diff.size} kept values differ '
                            f'bitwise; max abs diff {np.abs(a[ok] - b[ok]).max()}')


def assert_nanfilled_bitwise(out, ref_data, ref_mask, what=''):
    """``out`` is NaN-filled: NaN exactly where masked (or where the reference
    itself holds NaN), bit-identical elsewhere."""
    out = np.asarray(out)
    assert out.shape == ref_data.shape, (what, out.shape, ref_data.shape)
    assert out.dtype == np.float64, (what, out.dtype)
    expect_nan = np.asarray(ref_mask, bool) | np.isnan(ref_data)
    np.testing.assert_array_equal(np.isnan(out), expect_nan,
                                  err_msg=f'{what}: NaN placement differs')
    ok = ~expect_nan
    diff = bits(out[ok]) != bits(ref_data[ok])
    assert not diff.any(), (f'{what}: {int(diff.sum())} of {diff.size} values differ '
                            f'bitwise; max abs diff {np.abs(out[ok] - ref_data[ok]).max()}')
