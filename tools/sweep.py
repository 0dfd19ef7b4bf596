#!/usr/bin/env python
"""Kernel-variant sweep on one B200 (development tool, not part of the product).

Times b200remap_spmm with CUDA events for the BASELINE configs across the
library's tunables; prints one table row per variant:  config, variant, ms,
algorithmic GB/s, fraction of the measured HBM peak.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, mapfile, synthetic as syn  # noqa: E402

try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    PEAK = 6650.0


def device_csr(m):
    ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1,
                                   m.n_b, m.n_a)
    return mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b).on_device(0)


def alg_bytes(csr, K, w=8, w_out=8):
    return csr.nnz * 12 + (csr.n_row + 1) * 4 + csr.n_touched * K * w + csr.n_row * K * w_out


def time_launch(fn, reps=10, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(reps)]
    for i, (a, b) in enumerate(evs):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def report(tag, variant, ms, best, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(f'{tag:28s} {variant:34s} median {ms*1e3:9.1f} us  best {best*1e3:9.1f} us  '
          f'{gbs:8.1f} GB/s  {gbs / PEAK * 100:5.1f}% of measured peak', flush=True)


def make_ring(n_a, K, n, masked, dtype=torch.float64):
    ring = torch.empty((n, n_a, K), dtype=dtype, device='cuda')
    ring.uniform_(-2.0, 30.0)
    if masked:
        lv = torch.from_numpy(syn.bathymetry_levels(n_a, K, seed=5)).cuda()
        ring.masked_fill_((torch.arange(K, device='cuda')[None, :] >= lv[:, None])[None], float('nan'))
    return ring


def run_spmm(csr, ring, y, K, nb, mode, i, kernel=0):
    st = torch.cuda.current_stream().cuda_stream
    n = ring.shape[0]
    base = ((i * nb) % n) if nb < n else 0
    code = _cabi.F64 if ring.dtype == torch.float64 else _cabi.F32
    csr.spmm(ring[base].data_ptr(), code, K, K, nb, ring.shape[1] * K, y.data_ptr(), K,
             csr.n_row * K, mode, 0.01, kernel=kernel, stream=st,
             y_f32=y.dtype == torch.float32)


def sweep_c3(args):
    m = syn.make_c3(scale=args.scale)
    csr = device_csr(m)
    K = 80
    print(f'# C3 n_a={m.n_a} n_b={m.n_b} nnz={csr.nnz} touched={csr.n_touched} '
          f'B/slice={alg_bytes(csr, K) / 1e6:.1f} MB', flush=True)
    y = torch.empty((8, m.n_b, K), dtype=torch.float64, device='cuda')
    for masked in (True, False):
        ring = make_ring(m.n_a, K, 8, masked)
        mode = _cabi.MODE_MASKED if masked else _cabi.MODE_FRACB
        tag = 'C3 masked' if masked else 'C3 unmasked'
        for nb in (1, 8):
            nbytes = alg_bytes(csr, K) * nb
            # (kernel, target threads, gather policy, max straight-line class)
            variants = [(7, 0, 0, 0)]
            if args.full:
                variants += [(1, 160, 0, 0)]
            for kern, threads, pol, maxn in variants:
                for which, v in ((0, threads), (1, pol), (5, maxn)):
                    _cabi.set_tunable(which, v)
                ms, best = time_launch(lambda i: run_spmm(csr, ring, y, K, nb, mode, i, kern))
                report(f'{tag} x{nb}', f'kernel={kern} thr={threads} pol={pol} maxn={maxn}',
                       ms, best, nbytes)
            for which in range(7):
                _cabi.set_tunable(which, 0)
        if masked:
            # fp32 input
            ring32 = make_ring(m.n_a, K, 4, True, torch.float32)
            nbytes = alg_bytes(csr, K, 4) * 4
            ms, best = time_launch(lambda i: run_spmm(csr, ring32, y, K, 4, mode, i, 0))
            report('C3 masked f32-in x4', 'default', ms, best, nbytes)
            y32 = torch.empty((8, m.n_b, K), dtype=torch.float32, device='cuda')
            ms, best = time_launch(lambda i: run_spmm(csr, ring32, y32, K, 4, mode, i, 0))
            report('C3 masked f32-in f32-out x4', 'default', ms, best, alg_bytes(csr, K, 4, 4) * 4)
            ms, best = time_launch(lambda i: run_spmm(csr, ring, y32, K, 8, mode, i, 0))
            report('C3 masked f64-in f32-out x8', 'default', ms, best, alg_bytes(csr, K, 8, 4) * 8)
            del y32
            del ring32
        del ring
    # any-NaN scan over a NaN-free slice (worst case: no early exit) and a masked one
    from pyremap_b200.engine import device_any_nan
    x = make_ring(m.n_a, K, 1, False)
    flag = torch.empty(1, dtype=torch.int32, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    ms, best = time_launch(lambda i: _cabi.any_nan(x.data_ptr(), 0, x.numel(), flag.data_ptr(), st))
    report('any_nan NaN-free slice', 'full scan', ms, best, x.numel() * 8)
    x[0, 100, 3] = float('nan')
    ms, best = time_launch(lambda i: _cabi.any_nan(x.data_ptr(), 0, x.numel(), flag.data_ptr(), st))
    report('any_nan early NaN', 'early exit', ms, best, x.numel() * 8)
    assert device_any_nan(x)


def sweep_c2(args):
    m = syn.make_c2(scale=args.scale)
    csr = device_csr(m)
    print(f'# C2 n_a={m.n_a} n_b={m.n_b} nnz={csr.nnz} touched={csr.n_touched}', flush=True)
    # native (Time=12, nCells, nVertLevels=60) batched, and flat K=720
    ring = make_ring(m.n_a, 60, 12, True)
    y = torch.empty((12, m.n_b, 60), dtype=torch.float64, device='cuda')
    nbytes = alg_bytes(csr, 720)
    for kern, lw in ((7, 0), (7, 3), (7, 4), (7, 5), (1, 0)):
        _cabi.set_tunable(0, lw)
        ms, best = time_launch(lambda i: run_spmm(csr, ring, y, 60, 12, _cabi.MODE_MASKED, i, kern))
        report('C2 masked (12,nCells,60)', f'batched x12 K=60 kernel={kern} lw={lw}', ms, best, nbytes)
    _cabi.set_tunable(0, 0)
    flat = ring.permute(1, 0, 2).reshape(1, m.n_a, 720).contiguous()
    y2 = torch.empty((1, m.n_b, 720), dtype=torch.float64, device='cuda')
    for kern, lw in ((0, 0), (7, 4), (7, 5), (7, 6), (1, 0)):
        _cabi.set_tunable(0, lw)         # WROW: lanes per row = 1 << (lw - 1)
        ms, best = time_launch(lambda i: run_spmm(csr, flat, y2, 720, 1, _cabi.MODE_MASKED, i, kern))
        report('C2 masked [nCells,720]', f'flat K=720 kernel={kern} lw={lw}', ms, best, nbytes)
    _cabi.set_tunable(0, 0)


def sweep_c1(args):
    m = syn.make_c1()
    csr = device_csr(m)
    x = torch.randn((1, m.n_a, 10), dtype=torch.float64, device='cuda')
    y = torch.empty((1, m.n_b, 10), dtype=torch.float64, device='cuda')
    for kern in (7, 1):
        ms, best = time_launch(lambda i: run_spmm(csr, x, y, 10, 1, _cabi.MODE_FRACB, i, kern), reps=50)
        report('C1 unmasked K=10', f'kernel={kern} (latency)', ms, best, alg_bytes(csr, 10))


def sweep_c4(args):
    m = syn.make_c4(scale=args.scale)
    csr = device_csr(m)
    print(f'# C4 n_a={m.n_a} n_b={m.n_b} nnz={csr.nnz}', flush=True)
    for K in (1, 4):
        x = make_ring(m.n_a, K, 2, False)
        x[:, :: 97, :] = float('nan')
        y = torch.empty((1, m.n_b, K), dtype=torch.float64, device='cuda')
        for kernel, name, ts in ((1, 'lanes_k', 0),):
            _cabi.set_tunable(5, ts)
            for mode, mname in ((_cabi.MODE_MASKED, 'masked'), (_cabi.MODE_FRACB, 'fracb')):
                xin = x if mode == _cabi.MODE_MASKED else torch.nan_to_num(x, nan=1.0)
                ms, best = time_launch(lambda i: run_spmm(csr, xin, y, K, 1, mode, i, kernel))
                report(f'C4 {mname} K={K}', name, ms, best, alg_bytes(csr, K))
        _cabi.set_tunable(5, 0)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='c3,c2,c1,c4')
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--full', type=int, default=0)
    a = ap.parse_args()
    for c in a.configs.split(','):
        {'c3': sweep_c3, 'c2': sweep_c2, 'c1': sweep_c1, 'c4': sweep_c4}[c](a)
        torch.cuda.empty_cache()
