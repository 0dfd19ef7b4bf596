#!/usr/bin/env python
"""First-call cost of a map (development probe): CSR build, touched-row cover, device upload
(`b200remap_csr_create`: binning + ELL build), first product.  Usage: first_call_probe.py c3|c4"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, mapfile, synthetic as syn  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else 'c4'
    torch.cuda.init()
    torch.zeros(1, device='cuda')
    t = time.perf_counter()
    m = syn.make_c4() if which == 'c4' else syn.make_c3()
    print(f'{which}: synthetic map {time.perf_counter() - t:.2f} s  n_a={m.n_a} n_b={m.n_b} n_s={m.n_s}')
    row, col = m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1
    t = time.perf_counter()
    ip, ix, d = mapfile.coo_to_csr_gpu(m.S, row, col, m.n_b, m.n_a, 0)
    print(f'coo_to_csr_gpu          {time.perf_counter() - t:.3f} s')
    W = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
    t = time.perf_counter()
    W._touched_rows()
    print(f'touched rows            {time.perf_counter() - t:.3f} s')
    t = time.perf_counter()
    cov = W.cover_exact()
    print(f'cover_exact             {time.perf_counter() - t:.3f} s  ({"none" if cov is None else cov["n_cover"]})')
    t = time.perf_counter()
    csr = W.on_device(0)
    torch.cuda.synchronize()
    print(f'csr_create (full)       {time.perf_counter() - t:.3f} s')
    if cov is not None:
        t = time.perf_counter()
        W.on_device_cover(0, exact=True)
        torch.cuda.synchronize()
        print(f'csr_create (cover)      {time.perf_counter() - t:.3f} s')
    K = 1 if which == 'c4' else 80
    x = torch.rand((m.n_a, K), dtype=torch.float64, device='cuda')
    y = torch.empty((m.n_b, K), dtype=torch.float64, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for i in range(2):
        t = time.perf_counter()
        csr.spmm(x.data_ptr(), _cabi.F64, K, K, 1, m.n_a * K, y.data_ptr(), K, m.n_b * K,
                 _cabi.MODE_FRACB, 0.0, stream=st)
        torch.cuda.synchronize()
        print(f'product #{i}              {time.perf_counter() - t:.4f} s')


if __name__ == '__main__':
    main()
