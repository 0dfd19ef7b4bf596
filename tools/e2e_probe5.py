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
