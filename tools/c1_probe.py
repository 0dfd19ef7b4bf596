#!/usr/bin/env python
"""C1 (2 deg -> 1 deg bilinear, K = 10, frac_b branch): latency of the kernels (development probe)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, mapfile, synthetic as syn  # noqa: E402

_cabi.set_tunable(2, 1)
m = syn.make_c1()
ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1, m.n_b, m.n_a)
csr = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b).on_device(0)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for K in (10, 12, 16, 1):
    x = torch.rand((m.n_a, K), dtype=torch.float64, device='cuda')
    y = torch.empty((m.n_b, K), dtype=torch.float64, device='cuda')
    ref = None
    for kern in (7, 1, 8):
        ts = []
        for i in range(30):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            csr.spmm(x.data_ptr(), _cabi.F64, K, K, 1, 0, y.data_ptr(), K, 0, _cabi.MODE_FRACB, 0.0,
                     kernel=kern, stream=st)
            b.record()
            ts.append((a, b))
        torch.cuda.synchronize()
        if ref is None:
            ref = y.clone()
        assert torch.equal(ref.view(torch.int64), y.view(torch.int64))
        t = sorted(a.elapsed_time(b) for a, b in ts)
        print(f'C1 K={K:3d} kernel={kern}: median {t[len(t) // 2] * 1e3:6.1f} us  best {t[0] * 1e3:6.1f} us', flush=True)
