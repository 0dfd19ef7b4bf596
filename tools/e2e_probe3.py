#!/usr/bin/env python
"""A/B of the bridged-gap run cover in one process (development tool, one B200)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyremap_b200
from pyremap_b200 import _cabi, mapfile, synthetic as syn


def timed(fn, n=5):
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - a)
    return min(ts) * 1e3, float(np.median(ts)) * 1e3


m = syn.make_c3()
ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1, m.n_b, m.n_a)
T, L = 8, 80
host_t = torch.empty((T, m.n_a, L), dtype=torch.float64, pin_memory=True)
g = torch.empty((m.n_a, L), dtype=torch.float64, device='cuda')
for i in range(T):
    g.uniform_(-2, 30); g[::3, 40:] = float('nan'); host_t[i].copy_(g)
torch.cuda.synchronize()
host = host_t.numpy()
for rep in range(2):
    for slack in (0.0, 0.015, 0.05):
        W = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
        cov = W.cover_exact(slack=slack)
        r = pyremap_b200.Remapper(map_filename='x', src_descriptor=m.src_descriptor, dst_descriptor=m.dst_descriptor)
        r._matrix = W
        r._ds_map = mapfile.MapDataset({'dst_grid_dims': m.dst_grid_dims, 'frac_b': m.frac_b, 'src_grid_dims': m.src_grid_dims}, {})
        r.remap_array(host, [1], 0.01); r.remap_array(host, [1], 0.01)
        best, med = timed(lambda: r.remap_array(host, [1], 0.01), n=5)
        print(f'slack={slack}: runs {cov["run_start"].size} rows {cov["n_cover"]}: remap_array(T=8) best {best:.1f} ms median {med:.1f} ms = {best / T:.2f} ms/slice', flush=True)
        W.release()
