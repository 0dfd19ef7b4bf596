"""Differential test of the oracle against the LIVE reference module.  Runs only
where ``/root/reference`` exists (the authoring container); elsewhere the
committed golden fixtures stand in (tests/test_oracle.py)."""

import numpy as np
import pytest

from _util import assert_bitwise
from oracle import ref_loader, remap_oracle
from pyremap_b200 import synthetic as syn

pytestmark = pytest.mark.skipif(not ref_loader.available(),
                                reason='reference tree not present')


def _both(m, arg, axes, thr):
    matrix = remap_oracle.build_matrix(m.S, m.row, m.col, m.n_b, m.n_a)
    ref = ref_loader.reference_remap_array(matrix, m.frac_b, m.dst_grid_dims, arg, axes, thr)
    mine = remap_oracle.remap_array(matrix, m.frac_b, m.dst_grid_dims, arg, axes, thr)
    assert_bitwise(np.ma.getdata(mine), ~np.ma.getmaskarray(mine), np.ma.getdata(ref),
                   ~np.ma.getmaskarray(ref))
    step = remap_oracle.remap_array_stepwise(matrix, m.frac_b, m.dst_grid_dims, arg, axes, thr)
    assert_bitwise(np.ma.getdata(step), ~np.ma.getmaskarray(step), np.ma.getdata(ref),
                   ~np.ma.getmaskarray(ref))


@pytest.mark.parametrize('seed', range(4))
@pytest.mark.parametrize('thr', [None, 0.01, 0.6])
def test_random_c2_like(seed, thr):
    rng = np.random.default_rng(seed)
    m = syn.make_c2(scale=0.004, seed=seed + 10)
    f = rng.normal(size=(3, m.n_a, 5)) * 10.0 ** rng.integers(-3, 4, size=(3, 1, 5))
    nanmask = rng.random(f.shape) < 0.2
    f[nanmask] = np.nan
    _both(m, np.ma.masked_array(f, nanmask), [1], thr)
    _both(m, f, [1], thr)


@pytest.mark.parametrize('scale', [0.002, 0.01])
def test_random_c3_c4_like(scale):
    rng = np.random.default_rng(3)
    m = syn.make_c3(scale=scale)
    f = rng.normal(size=(m.n_a, 9)).astype(np.float32)
    _both(m, f, [0], None)
    m4 = syn.make_c4(scale=0.02, ratio=6)
    ny, nx = m4.src_descriptor.dim_sizes
    g = rng.normal(size=(2, ny, nx))
    g[:, :7, :] = np.nan
    _both(m4, np.ma.masked_array(g, np.isnan(g)), [1, 2], 0.01)
