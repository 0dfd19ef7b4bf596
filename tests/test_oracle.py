"""The oracle against the reference's own outputs (golden fixtures) and against
scipy's kernel.  CPU only."""

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from _util import ARRAY_CASES, assert_bitwise, bits, load_case, reference_argument
from oracle import c_oracle, remap_oracle


@pytest.mark.parametrize('name', ARRAY_CASES)
def test_oracle_matches_reference_golden(name):
    case = load_case(name)
    mp = case['map']
    matrix = remap_oracle.build_matrix(mp['S'], mp['row'], mp['col'], mp['n_b'], mp['n_a'])
    out = remap_oracle.remap_array(matrix, mp['frac_b'], mp['dst_grid_dims'],
                                   reference_argument(case), case['remap_axes'],
                                   case['thr'])
    assert isinstance(out, np.ma.MaskedArray)
    assert_bitwise(np.ma.getdata(out), ~np.ma.getmaskarray(out), case['out_data'],
                   ~case['out_mask'], name)


@pytest.mark.parametrize('name', ARRAY_CASES)
def test_stepwise_port_matches_reference_golden(name):
    case = load_case(name)
    if case['field'].dtype.kind in 'iu':
        pytest.skip('in-place divide of the stepwise port needs a float product')
    mp = case['map']
    matrix = remap_oracle.build_matrix(mp['S'], mp['row'], mp['col'], mp['n_b'], mp['n_a'])
    out = remap_oracle.remap_array_stepwise(matrix, mp['frac_b'], mp['dst_grid_dims'],
                                            reference_argument(case),
                                            case['remap_axes'], case['thr'])
    assert_bitwise(np.ma.getdata(out), ~np.ma.getmaskarray(out), case['out_data'],
                   ~case['out_mask'], name)


def _random_csr(rng, n_row, n_col, max_nnz):
    counts = rng.integers(0, max_nnz + 1, size=n_row)
    rows = np.repeat(np.arange(n_row), counts)
    cols = np.concatenate([np.sort(rng.choice(n_col, c, replace=False)) for c in counts]
                          + [np.zeros(0, np.int64)]).astype(np.int64)
    vals = rng.normal(size=rows.size)
    return csr_matrix((vals, (rows, cols)), shape=(n_row, n_col))


@pytest.mark.parametrize('k', [1, 3, 16])
def test_restated_csr_matvecs_is_bitwise_scipy(k):
    rng = np.random.default_rng(k)
    A = _random_csr(rng, 200, 150, 40)
    X = rng.normal(size=(150, k)) * 10.0 ** rng.integers(-8, 8, size=(150, k))
    Y = A.dot(X)
    for mine in (remap_oracle.spmm_ordered(A.indptr, A.indices, A.data, X),
                 remap_oracle.spmm_rowloop(A.indptr, A.indices, A.data, X),
                 c_oracle.csr_matvecs(A, X), c_oracle.csr_matvecs(A, X, threads=4)):
        assert np.array_equal(bits(mine), bits(Y))


@pytest.mark.parametrize('name', ARRAY_CASES)
def test_c_oracle_fused_matches_golden(name):
    case = load_case(name)
    if len(case['remap_axes']) + 1 < case['field'].ndim and case['remap_axes'][0] != 0:
        pytest.skip('C oracle is checked on flat [n_a, K] layouts')
    mp = case['map']
    matrix = remap_oracle.build_matrix(mp['S'], mp['row'], mp['col'], mp['n_b'], mp['n_a'])
    arg = reference_argument(case)
    flat, extra = remap_oracle._flatten(arg, case['remap_axes'])
    masked = isinstance(flat, np.ma.MaskedArray) and case['thr'] is not None
    X = np.ma.getdata(flat).astype(np.float64)
    if masked:
        valid = ~np.ma.getmaskarray(flat)
        y, keep = c_oracle.remap_fused(matrix, mp['frac_b'], X, 2, case['thr'],
                                       valid=valid, want_keep=True, threads=2)
    else:
        y, keep = c_oracle.remap_fused(matrix, mp['frac_b'], X, 1, want_keep=True)
    ref = remap_oracle._flatten(np.ma.masked_array(case['out_data'], case['out_mask']),
                                list(range(len(mp['dst_grid_dims']))))[0] \
        if case['remap_axes'][0] == 0 else None
    if ref is None:
        pytest.skip('layout not flat')
    assert_bitwise(y, keep, np.ma.getdata(ref), ~np.ma.getmaskarray(ref), name)


def test_c_oracle_isnan_validity_equals_explicit_mask():
    rng = np.random.default_rng(5)
    A = _random_csr(rng, 120, 90, 12)
    X = rng.normal(size=(90, 7))
    X[rng.random(X.shape) < 0.25] = np.nan
    fb = rng.uniform(0.1, 1, size=120)
    y1, k1 = c_oracle.remap_fused(A, fb, X, 2, 0.05, want_keep=True)
    y2, k2 = c_oracle.remap_fused(A, fb, X, 2, 0.05, valid=~np.isnan(X), want_keep=True)
    assert np.array_equal(k1, k2)
    assert np.array_equal(bits(y1[k1]), bits(y2[k2]))
    vals, keep = remap_oracle.remap_flat(A, fb, X, ~np.isnan(X), 0.05)
    assert np.array_equal(keep, k1)
    assert np.array_equal(bits(vals[keep]), bits(y1[k1]))
