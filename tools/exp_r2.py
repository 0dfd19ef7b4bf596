#!/usr/bin/env python
"""Round-2 kernel experiments on one B200 (development tool, not part of the product).

C3 masked, (Time=nb, nCells, nVertLevels=80) launches: dynamic vs static item claiming,
resident warps per SM, binning segment length, slices per launch; short-run medians plus a
sustained loop (the bench's regime: power-capped clocks).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, mapfile, synthetic as syn  # noqa: E402
from tools.sweep import PEAK, alg_bytes, make_ring, time_launch  # noqa: E402


def launch_bytes(csr, K, nb):
    """weights once per launch, fields per slice (SURVEY 8d with the weights counted once)"""
    w = csr.nnz * 12 + (csr.n_row + 1) * 4
    return w + nb * (csr.n_touched * K * 8 + csr.n_row * K * 8)


def fresh_csr(m, ip, ix, d, seg_units):
    _cabi.set_tunable(4, seg_units)
    csr = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b).on_device(0)
    _cabi.set_tunable(4, 0)
    return csr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--order', default='rings')
    ap.add_argument('--segs', default='0')
    ap.add_argument('--grids', default='0')
    ap.add_argument('--nbs', default='8')
    ap.add_argument('--sustain', type=int, default=1)
    ap.add_argument('--variants', default='0')
    ap.add_argument('--pfs', default='0')
    ap.add_argument('--wpcs', default='0')
    ap.add_argument('--carves', default='0')
    ap.add_argument('--dump', type=int, default=0)
    ap.add_argument('--group', type=int, default=0)
    ap.add_argument('--dyns', default='1,0')
    ap.add_argument('--kernels', default='7')
    ap.add_argument('--mode', default='masked')
    a = ap.parse_args()
    cache = f'/tmp/c3_csr_{a.order}.npz'
    if os.path.exists(cache):
        z = np.load(cache)
        ip, ix, d = z['ip'], z['ix'], z['d']

        class M:
            pass
        m = M()
        m.n_a, m.n_b, m.frac_b = int(z['n_a']), int(z['n_b']), z['frac_b']
    else:
        m = syn.make_c3(order=a.order)
        ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1,
                                       m.n_b, m.n_a)
        np.savez(cache, ip=ip, ix=ix, d=d, frac_b=m.frac_b, n_a=m.n_a, n_b=m.n_b)
    K = 80
    nbs = [int(v) for v in a.nbs.split(',')]
    ring = make_ring(m.n_a, K, max(8, max(nbs)), a.mode == 'masked')
    mode = _cabi.MODE_MASKED if a.mode == 'masked' else _cabi.MODE_FRACB
    y = torch.empty((max(nbs), m.n_b, K), dtype=torch.float64, device='cuda')
    st = torch.cuda.current_stream().cuda_stream

    def run(csr, nb, i, kern=7):
        n = ring.shape[0]
        base = ((i * nb) % n) if nb < n else 0
        csr.spmm(ring[base].data_ptr(), _cabi.F64, K, K, nb, m.n_a * K, y.data_ptr(), K,
                 m.n_b * K, mode, 0.01, kernel=kern, stream=st)

    for seg in [int(v) for v in a.segs.split(',')]:
        csr = fresh_csr(m, ip, ix, d, seg)
        for nb in nbs:
            nbytes = launch_bytes(csr, K, nb)
            for dyn, var, kern, pf, wpc, carve in [
                    (int(dy), int(v), int(k), int(f), int(w), int(cv)) for k in a.kernels.split(',')
                    for dy in a.dyns.split(',') for v in a.variants.split(',')
                    for f in a.pfs.split(',') for w in a.wpcs.split(',') for cv in a.carves.split(',')]:
                for grid in [int(v) for v in a.grids.split(',')]:
                    _cabi.set_tunable(8, 0 if dyn else 1)
                    _cabi.set_tunable(9, var)
                    _cabi.set_tunable(13, pf)
                    _cabi.set_tunable(14, wpc)
                    _cabi.set_tunable(15, carve)
                    _cabi.set_tunable(7, grid)
                    _cabi.set_tunable(12, a.group)
                    ms, best = time_launch(lambda i: run(csr, nb, i, kern), reps=12)
                    if a.dump:
                        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                               for _ in range(24)]
                        for i, (e0, e1) in enumerate(evs):
                            e0.record()
                            run(csr, nb, i, kern)
                            e1.record()
                        torch.cuda.synchronize()
                        print('   per-launch us:', ' '.join(f'{e0.elapsed_time(e1) * 1e3:.0f}' for e0, e1 in evs),
                              flush=True)
                    gbs = nbytes / (ms * 1e-3) / 1e9
                    line = (f'{a.mode} order={a.order} seg={seg or 256} group={a.group or 8} nb={nb} kernel={kern} dyn={dyn} pf={pf} wpc={wpc} carve={carve} '
                            f'warps/SM={grid or "occ"}  '
                            f'median {ms * 1e3:8.1f} us best {best * 1e3:8.1f} us  {gbs:7.1f} GB/s '
                            f'{gbs / PEAK * 100:5.1f}%')
                    if a.sustain and grid == 0:
                        reps = 46 * 8 // nb * 6
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for i in range(reps):
                            run(csr, nb, i, kern)
                        e1.record()
                        torch.cuda.synchronize()
                        sms = e0.elapsed_time(e1) / reps
                        g2 = nbytes / (sms * 1e-3) / 1e9
                        line += f'   sustained {sms * 1e3:8.1f} us {g2:7.1f} GB/s {g2 / PEAK * 100:5.1f}%'
                    print(line, flush=True)
        for t in (7, 8, 9, 12, 13, 14, 15):
            _cabi.set_tunable(t, 0)
        del csr


if __name__ == '__main__':
    main()
