#!/usr/bin/env python
"""Where does end-to-end time go?  (development probe, one B200)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyremap_b200
from pyremap_b200 import _cabi, mapfile, synthetic as syn

def t(fn, n=3):
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - a)
    return min(ts) * 1e3

m = syn.make_c3()
ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1, m.n_b, m.n_a)
W = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
cov = W.cover()
print('cover runs', len(cov['runs']), 'n_cover', cov['n_cover'], 'touched', np.unique(ix).size)
for mr in (64, 256, 1024, 4096):
    W2 = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b); c2 = W2.cover(max_runs=mr)
    print('  max_runs', mr, '->', len(c2['runs']), 'runs, n_cover', c2['n_cover'])
T, L = 4, 80
host_t = torch.empty((T, m.n_a, L), dtype=torch.float64, pin_memory=True)
g = torch.empty((m.n_a, L), dtype=torch.float64, device='cuda')
for i in range(T):
    g.uniform_(-2, 30); g[::3, 40:] = float('nan'); host_t[i].copy_(g)
torch.cuda.synchronize()
host = host_t.numpy()
print('is_pinned(from_numpy):', torch.from_numpy(host).is_pinned())
n_x = cov['n_cover']
xd = torch.empty((n_x, L), dtype=torch.float64, device='cuda')
big = host_t[0, :n_x]
print('H2D contiguous %d MB pinned: %.2f ms' % (big.numel() * 8 / 1e6, t(lambda: xd.copy_(big, non_blocking=True))))
def runs_copy():
    for s, l, p in cov['runs']:
        xd[p:p + l].copy_(host_t[0, s:s + l], non_blocking=True)
print('H2D in %d runs: %.2f ms' % (len(cov['runs']), t(runs_copy)))
yd = torch.empty((m.n_b, L), dtype=torch.float64, device='cuda')
yo = torch.empty((m.n_b, L), dtype=torch.float64, pin_memory=True)
print('D2H %d MB pinned: %.2f ms' % (yd.numel() * 8 / 1e6, t(lambda: yo.copy_(yd, non_blocking=True))))
print('pinned alloc out (T slices): %.2f ms' % t(lambda: torch.empty((T, m.n_b, L), dtype=torch.float64, pin_memory=True), n=2))
print('host_any_nan early: %.3f ms' % t(lambda: _cabi.host_any_nan(host), n=3))
clean = np.nan_to_num(host[0])
print('host_any_nan full scan of one slice: %.1f ms' % t(lambda: _cabi.host_any_nan(clean), n=2))
r = pyremap_b200.Remapper(map_filename='x', src_descriptor=m.src_descriptor, dst_descriptor=m.dst_descriptor)
r._matrix = W; r._ds_map = mapfile.MapDataset({'dst_grid_dims': m.dst_grid_dims, 'frac_b': m.frac_b, 'src_grid_dims': m.src_grid_dims}, {})
r.remap_array(host, [1], 0.01)
print('remap_array(T=%d) total: %.1f ms' % (T, t(lambda: r.remap_array(host, [1], 0.01), n=3)))
keep = r.remap_array(host, [1], 0.01)
print('remap_array(T=%d) with previous result alive: %.1f ms' % (T, t(lambda: r.remap_array(host, [1], 0.01), n=3)))
print('remap_array(T=1): %.1f ms' % t(lambda: r.remap_array(host[:1], [1], 0.01), n=3))
