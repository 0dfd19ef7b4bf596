"""Multi-GPU K-sharding on real devices (needs >= 2 GPUs; skipped on a 1-GPU box, where the
world_size-2 gloo test covers the driver logic): every rank remaps its block of time slices
with the replicated weights, no collective in the data path; the optional NCCL all-gather of
the outputs reproduces the single-GPU result bit for bit."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        import pyremap_b200
        from pyremap_b200 import mapfile, synthetic as syn
        from pyremap_b200.sharding import ShardedRemap
        m = syn.make_c3(scale=0.05)
        ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1,
                                       m.n_b, m.n_a)
        r = pyremap_b200.Remapper(map_filename='in-memory', src_descriptor=m.src_descriptor,
                                  dst_descriptor=m.dst_descriptor)
        r._matrix = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
        r._ds_map = mapfile.MapDataset({'dst_grid_dims': m.dst_grid_dims, 'frac_b': m.frac_b,
                                        'src_grid_dims': m.src_grid_dims}, {})
        r.device = rank
        T, L = 7, 16
        lv = syn.bathymetry_levels(m.n_a, L, seed=3)
        field = np.stack([syn.ocean_field(m.n_a, L, seed=40 + t, max_level=lv) for t in range(T)])
        sh = ShardedRemap(r, [1], 0.01)
        lo, hi = sh.local_slices(T)
        local = sh.remap_local(torch.from_numpy(field).cuda(rank))
        assert local.shape[0] == hi - lo and local.device.index == rank
        full = sh.gather(local, T)
        whole = r.remap_array(torch.from_numpy(field).cuda(rank), [1], 0.01, return_torch=True)
        assert torch.equal(torch.isnan(full), torch.isnan(whole))
        assert torch.equal(torch.nan_to_num(full).view(torch.int64),
                           torch.nan_to_num(whole).view(torch.int64))
        open(os.path.join(out_dir, f'ok{rank}'), 'w').close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_gpu_sharded_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']
