#!/usr/bin/env python
"""PCIe overlap probe for the streamed host path (development tool, one B200)."""
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
    return min(ts) * 1e3


m = syn.make_c3()
ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1, m.n_b, m.n_a)
W = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
cov = W.cover_exact()
rows_dev = torch.from_numpy(cov['rows']).cuda()
n_x = cov['n_cover']
T, L = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 80
host_t = torch.empty((T, m.n_a, L), dtype=torch.float64, pin_memory=True)
g = torch.empty((m.n_a, L), dtype=torch.float64, device='cuda')
for i in range(T):
    g.uniform_(-2, 30); g[::3, 40:] = float('nan'); host_t[i].copy_(g)
torch.cuda.synchronize()
xd = torch.empty((n_x, L), dtype=torch.float64, device='cuda')
yd = torch.empty((m.n_b, L), dtype=torch.float64, device='cuda')
yo = torch.empty((m.n_b, L), dtype=torch.float64, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
mb_in, mb_out = n_x * L * 8 / 1e6, m.n_b * L * 8 / 1e6


def gather(stream):
    _cabi.gather_rows(host_t[0].data_ptr(), xd.data_ptr(), rows_dev.data_ptr(), n_x, L * 8, L * 8, stream.cuda_stream)


def d2h(stream):
    with torch.cuda.stream(stream):
        yo.copy_(yd, non_blocking=True)


t = timed(lambda: gather(s1)); print(f'gather {mb_in:.0f} MB from pinned (zero-copy kernel): {t:.2f} ms = {mb_in / t:.1f} GB/s')
big = host_t[0, :n_x]
t = timed(lambda: xd.copy_(big, non_blocking=True)); print(f'H2D contiguous DMA {mb_in:.0f} MB: {t:.2f} ms = {mb_in / t:.1f} GB/s')
t = timed(lambda: d2h(s2)); print(f'D2H DMA {mb_out:.0f} MB: {t:.2f} ms = {mb_out / t:.1f} GB/s')
t = timed(lambda: (gather(s1), d2h(s2))); print(f'gather || D2H: {t:.2f} ms = {(mb_in + mb_out) / t:.1f} GB/s aggregate')
def both_dma():
    with torch.cuda.stream(s1):
        xd.copy_(big, non_blocking=True)
    d2h(s2)
t = timed(both_dma); print(f'H2D DMA || D2H DMA: {t:.2f} ms = {(mb_in + mb_out) / t:.1f} GB/s aggregate')
ident = torch.arange(m.n_b, dtype=torch.int32, device='cuda')
def kd2h(stream):
    _cabi.gather_rows(yd.data_ptr(), yo.data_ptr(), ident.data_ptr(), m.n_b, L * 8, L * 8, stream.cuda_stream)
t = timed(lambda: kd2h(s2)); print(f'kernel D2H (SM stores to pinned) {mb_out:.0f} MB: {t:.2f} ms = {mb_out / t:.1f} GB/s')
t = timed(lambda: (gather(s1), kd2h(s2))); print(f'gather || kernel D2H: {t:.2f} ms = {(mb_in + mb_out) / t:.1f} GB/s aggregate')
for blocks in (64, 32):
    _cabi.set_tunable(10, blocks)
    t = timed(lambda: (gather(s1), kd2h(s2))); print(f'blocks={blocks}: gather || kernel D2H: {t:.2f} ms = {(mb_in + mb_out) / t:.1f} GB/s aggregate')
_cabi.set_tunable(10, 0)
for blocks in ():
    _cabi.set_tunable(10, blocks)
    ta = timed(lambda: gather(s1))
    tb = timed(lambda: (gather(s1), d2h(s2)))
    print(f'gather blocks={blocks}: alone {ta:.2f} ms ({mb_in / ta:.1f} GB/s)   || D2H {tb:.2f} ms ({(mb_in + mb_out) / tb:.1f} GB/s aggregate)')
_cabi.set_tunable(10, 0)
r = pyremap_b200.Remapper(map_filename='x', src_descriptor=m.src_descriptor, dst_descriptor=m.dst_descriptor)
r._matrix = W
r._ds_map = mapfile.MapDataset({'dst_grid_dims': m.dst_grid_dims, 'frac_b': m.frac_b, 'src_grid_dims': m.src_grid_dims}, {})
host = host_t.numpy()
rb = L * 8
so, do, nb = cov['run_start'] * rb, cov['run_pos'] * rb, cov['run_len'] * rb
print('runs', so.size)
t = timed(lambda: _cabi.copy_runs(host_t[0].data_ptr(), xd.data_ptr(), so, do, nb, s1.cuda_stream)); print(f'copy_runs batch, 1 stream: {t:.2f} ms = {mb_in / t:.1f} GB/s')
t = timed(lambda: _cabi.copy_runs(host_t[0].data_ptr(), xd.data_ptr(), so, do, nb, s1.cuda_stream, use_batch=False)); print(f'copy_runs loop, 1 stream: {t:.2f} ms = {mb_in / t:.1f} GB/s')
s3 = torch.cuda.Stream()
for parts in (2, 4):
    streams = [s1, s3, torch.cuda.Stream(), torch.cuda.Stream()][:parts]
    # split runs into `parts` groups of about equal bytes
    cum = np.cumsum(nb); cuts = [0] + [int(np.searchsorted(cum, cum[-1] * k / parts)) for k in range(1, parts)] + [so.size]
    def multi():
        for k, stq in enumerate(streams):
            a, b = cuts[k], cuts[k + 1]
            _cabi.copy_runs(host_t[0].data_ptr(), xd.data_ptr(), so[a:b], do[a:b], nb[a:b], stq.cuda_stream)
    t = timed(multi); print(f'copy_runs batch, {parts} streams: {t:.2f} ms = {mb_in / t:.1f} GB/s')
    t = timed(lambda: (multi(), d2h(s2))); print(f'copy_runs batch, {parts} streams || D2H: {t:.2f} ms = {(mb_in + mb_out) / t:.1f} GB/s aggregate')
t = timed(lambda: (_cabi.copy_runs(host_t[0].data_ptr(), xd.data_ptr(), so, do, nb, s1.cuda_stream), d2h(s2))); print(f'copy_runs batch 1 stream || D2H: {t:.2f} ms = {(mb_in + mb_out) / t:.1f} GB/s aggregate')
# merged runs (bridge gaps <= G rows): fewer, longer copies, a few more bytes
for G in (64, 256, 1024):
    st_, ln_ = cov['run_start'], cov['run_len']
    gaps = st_[1:] - (st_[:-1] + ln_[:-1])
    brk = np.nonzero(gaps > G)[0]
    ms_ = np.concatenate([[st_[0]], st_[brk + 1]]); me_ = np.concatenate([st_[brk] + ln_[brk], [st_[-1] + ln_[-1]]])
    ml = me_ - ms_; mp = np.concatenate([[0], np.cumsum(ml)[:-1]])
    big_x = torch.empty((int(ml.sum()), L), dtype=torch.float64, device='cuda')
    t = timed(lambda: _cabi.copy_runs(host_t[0].data_ptr(), big_x.data_ptr(), ms_ * rb, mp * rb, ml * rb, s1.cuda_stream))
    print(f'merged gaps<={G}: {ms_.size} runs, {ml.sum() * rb / 1e6:.0f} MB: {t:.2f} ms')
    del big_x
for h2d in ('gather', 'dma'):
    os.environ['B200REMAP_H2D'] = h2d
    for n in (1, 2, 4, T):
        r.remap_array(host[:n], [1], 0.01); r.remap_array(host[:n], [1], 0.01)
        t = timed(lambda: r.remap_array(host[:n], [1], 0.01), n=3)
        print(f'{h2d}: remap_array(T={n}): {t:.1f} ms = {t / n:.2f} ms/slice')
os.environ['B200REMAP_TRACE'] = '1'
r.remap_array(host, [1], 0.01)
