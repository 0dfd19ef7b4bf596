#!/usr/bin/env python
"""Segment size of the binned view x slices per launch (development tool, one B200)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, synthetic as syn  # noqa: E402
from tools import sweep  # noqa: E402

m = syn.make_c3()
K = 80
ring = sweep.make_ring(m.n_a, K, 8, True)
y = torch.empty((8, m.n_b, K), dtype=torch.float64, device='cuda')
for seg in (2048, 3072, 4096, 6144, 8192, 16384):
    _cabi.set_tunable(4, seg // 8)
    csr = sweep.device_csr(m)
    for kern in (7, 6):
        for nb in (2, 4, 8):
            ms, best = sweep.time_launch(lambda i: sweep.run_spmm(csr, ring, y, K, nb, _cabi.MODE_MASKED, i, kern))
            sweep.report(f'seg={seg} kernel={kern}', f'x{nb}', ms, best, sweep.alg_bytes(csr, K) * nb)
    csr.close()
_cabi.set_tunable(4, 0)
