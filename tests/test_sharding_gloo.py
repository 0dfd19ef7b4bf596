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


def _worker(rank, world, port, n_slices, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import remap_oracle
        from pyremap_b200 import synthetic as syn
        from pyremap_b200.sharding import ShardedRemap
        m = syn.make_c3(scale=0.002)
        A = remap_oracle.build_matrix(m.S, m.row, m.col, m.n_b, m.n_a)
        lv = syn.bathymetry_levels(m.n_a, 6, seed=1)
        field = np.stack([syn.ocean_field(m.n_a, 6, seed=10 + t, max_level=lv)
                          for t in range(n_slices)])

        def compute(local):
            if local.shape[0] == 0:
                return torch.empty((0,) + tuple(m.dst_descriptor.dim_sizes) + (6,),
                                   dtype=torch.float64)
            arg = np.ma.masked_array(local, np.isnan(local))
            out = remap_oracle.remap_array(A, m.frac_b, m.dst_grid_dims, arg, [1], 0.01)
            return torch.from_numpy(remap_oracle.nanfilled(out))

        sh = ShardedRemap(compute=compute)
        assert (sh.world, sh.rank) == (world, rank)
        local = sh.remap_local(field)
        lo, hi = sh.local_slices(n_slices)
        assert local.shape[0] == hi - lo
        full = sh.gather(local, n_slices)
        whole = compute(field)
        assert full.shape == whole.shape
        same = torch.equal(torch.nan_to_num(full, nan=-1.0).view(torch.int64),
                           torch.nan_to_num(whole, nan=-1.0).view(torch.int64))
        assert same and torch.equal(torch.isnan(full), torch.isnan(whole))
        open(os.path.join(out_dir, f'ok{rank}'), 'w').close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_slices', [5, 1])
def test_world_size_2_gloo_sharded_equals_unsharded(tmp_path, n_slices):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_slices, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']
