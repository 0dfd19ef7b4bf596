#!/usr/bin/env python
"""End to end with pageable vs pinned input (development tool, one B200)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyremap_b200
from pyremap_b200 import mapfile, synthetic as syn

def timed(fn, n=3):
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - a)
    return min(ts) * 1e3

m = syn.make_c3()
ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1, m.n_b, m.n_a)
W = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
r = pyremap_b200.Remapper(map_filename='x', src_descriptor=m.src_descriptor, dst_descriptor=m.dst_descriptor)
r._matrix = W
r._ds_map = mapfile.MapDataset({'dst_grid_dims': m.dst_grid_dims, 'frac_b': m.frac_b, 'src_grid_dims': m.src_grid_dims}, {})
T, L = 4, 80
pinned = torch.empty((T, m.n_a, L), dtype=torch.float64, pin_memory=True)
g = torch.empty((m.n_a, L), dtype=torch.float64, device='cuda')
for i in range(T):
    g.uniform_(-2, 30); g[::3, 40:] = float('nan'); pinned[i].copy_(g)
torch.cuda.synchronize()
pageable = pinned.numpy().copy()
for name, arr in (('pinned', pinned.numpy()), ('pageable', pageable)):
    r.remap_array(arr, [1], 0.01)
    t = timed(lambda: r.remap_array(arr, [1], 0.01))
    print(f'{name}: remap_array(T={T}) {t:.1f} ms = {t / T:.2f} ms/slice', flush=True)
os.environ['B200REMAP_TRACE'] = '1'
r.remap_array(pageable, [1], 0.01)
os.environ.pop('B200REMAP_TRACE', None)
keep = []
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    keep.append(r.remap_array(pinned.numpy(), [1], 0.01))
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3 * 1e3
print(f'pinned input, results kept alive (fresh pinned result block per call): {dt:.1f} ms per call = {dt / T:.2f} ms/slice')
t = timed(lambda: torch.empty((T, m.n_b, L), dtype=torch.float64, pin_memory=True), n=2)
keep2 = [torch.empty((T, m.n_b, L), dtype=torch.float64, pin_memory=True) for _ in range(2)]
torch.cuda.synchronize(); t0 = time.perf_counter(); keep2.append(torch.empty((T, m.n_b, L), dtype=torch.float64, pin_memory=True)); dt = (time.perf_counter() - t0) * 1e3
print(f'fresh pinned allocation of {T * m.n_b * L * 8 / 1e6:.0f} MB: {dt:.1f} ms')
t0 = time.perf_counter(); z = np.empty((T, m.n_b, L)); z[...] = 0; dt = (time.perf_counter() - t0) * 1e3
print(f'fresh pageable allocation + first touch of the same size: {dt:.1f} ms')
