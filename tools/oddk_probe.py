#!/usr/bin/env python
"""Odd K (vector width 1) against K padded to a multiple of 4 (256-bit lanes): development probe."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, mapfile, synthetic as syn  # noqa: E402

m = syn.make_c2()
ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1, m.n_b, m.n_a)
csr = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b).on_device(0)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for K, ld in ((365, 365), (365, 368), (368, 368), (31, 31), (31, 32), (32, 32), (30, 30), (30, 32)):
    x = torch.rand((m.n_a, ld), dtype=torch.float64, device='cuda')
    x[torch.rand(m.n_a, device='cuda') < 0.2] = float('nan')
    y = torch.empty((m.n_b, ld), dtype=torch.float64, device='cuda')
    ts = []
    for i in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        csr.spmm(x.data_ptr(), _cabi.F64, K, ld, 1, 0, y.data_ptr(), ld, 0, _cabi.MODE_MASKED, 0.01, stream=st)
        b.record()
        ts.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ts)
    print(f'C2 map K={K:4d} ld={ld:4d}: {t[len(t) // 2] * 1e3:8.1f} us', flush=True)
