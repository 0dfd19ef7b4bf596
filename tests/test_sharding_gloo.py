"""The N>1 path on CPU: world_size-2 ``gloo`` run of the K-sharding driver.  The
per-rank compute is the ORACLE here (tests may use it as the checker); on the GPU
box the same driver runs the CUDA path (tests/test_gpu_multi.py)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyremap_b200.sharding import shard_bounds, shard_counts


@pytest.mark.parametrize('n,world', [(365, 1), (365, 2), (365, 4), (365, 8), (3, 8), (0, 2), (12, 5)])
def test_shard_bounds_partition_everything_once(n, world):
    spans = [shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b
    counts = shard_counts(n, world)
    assert sum(counts) == n and max(counts) - min(counts) <= 1


def test_shard_bounds_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_slices, out_dir, nan_where='all'):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import remap_oracle
        from pyremap_b200 import synthetic as syn
        from pyremap_b200.sharding import ShardedRemap
        m = syn.make_c3(scale=0.002)
        A = remap_oracle.build_matrix(m.S, m.row, m.col, m.n_b, m.n_a)
        lv = syn.bathymetry_levels(m.n_a, 6, seed=1)
        # nan_where='last': only the last slice (= the last rank's shard) holds NaNs, so a rank
        # deciding the branch from its own block would take frac_b where the reference, which
        # decides per variable (remap_numpy.py:202-204), takes the masked branch
        field = np.stack([syn.ocean_field(m.n_a, 6, seed=10 + t,
                                          max_level=lv if (nan_where == 'all' or t == n_slices - 1)
                                          else None)
                          for t in range(n_slices)])

        def compute(local, masked):
            if local.shape[0] == 0:
                return torch.empty((0,) + tuple(m.dst_descriptor.dim_sizes) + (6,),
                                   dtype=torch.float64)
            # the reference wraps the WHOLE variable in a MaskedArray iff it holds any NaN
            arg = np.ma.masked_array(local, np.isnan(local)) if masked else local
            out = remap_oracle.remap_array(A, m.frac_b, m.dst_grid_dims, arg, [1], 0.01)
            return torch.from_numpy(remap_oracle.nanfilled(out))

        sh = ShardedRemap(compute=compute, renormalization_threshold=0.01)
        assert (sh.world, sh.rank) == (world, rank)
        local = sh.remap_local(field)
        lo, hi = sh.local_slices(n_slices)
        assert local.shape[0] == hi - lo
        full = sh.gather(local, n_slices)
        whole = compute(field, bool(np.isnan(field).any()))
        if nan_where == 'last' and n_slices > 1 and rank == 0:
            # the trap this guards against: rank 0's block is NaN-free, yet it must have
            # taken the masked branch (S@1 > thr), not frac_b
            wrong = compute(field[lo:hi], False)
            assert not torch.equal(torch.isnan(wrong), torch.isnan(local)) or \
                not torch.equal(torch.nan_to_num(wrong), torch.nan_to_num(local))
        assert full.shape == whole.shape
        same = torch.equal(torch.nan_to_num(full, nan=-1.0).view(torch.int64),
                           torch.nan_to_num(whole, nan=-1.0).view(torch.int64))
        assert same and torch.equal(torch.isnan(full), torch.isnan(whole))
        open(os.path.join(out_dir, f'ok{rank}'), 'w').close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_slices,nan_where', [(5, 'all'), (1, 'all'), (4, 'last')])
def test_world_size_2_gloo_sharded_equals_unsharded(tmp_path, n_slices, nan_where):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_slices, str(tmp_path), nan_where),
             nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']


def test_one_argument_compute_is_still_accepted():
    from pyremap_b200.sharding import ShardedRemap
    sh = ShardedRemap(compute=lambda block: block * 2)
    assert sh.remap_local(np.arange(4.0)).tolist() == [0.0, 2.0, 4.0, 6.0]


def test_block_has_nan_covers_numpy_and_tensors():
    from pyremap_b200.sharding import block_has_nan
    a = np.zeros((3, 5))
    assert not block_has_nan(a) and not block_has_nan(torch.from_numpy(a))
    a[2, 4] = np.nan
    assert block_has_nan(a) and block_has_nan(torch.from_numpy(a))
    assert block_has_nan(a[:, ::2])                       # non-contiguous view
    assert not block_has_nan(np.arange(6)) and not block_has_nan(np.empty((0, 4)))
