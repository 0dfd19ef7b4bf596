#!/usr/bin/env python
"""Launch one kernel configuration on C3 a few times (to be run under ncu; development tool).
usage: ncu_probe.py <kernel> <t5> <order-mode> [masked=1] [nb=8]"""
import sys
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, synthetic as syn  # noqa: E402
from tools import sweep  # noqa: E402


def main():
    kern, t5, order = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    masked = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    nb = int(sys.argv[5]) if len(sys.argv) > 5 else 8
    _cabi.set_tunable(9, order)
    m = syn.make_c3()
    csr = sweep.device_csr(m)
    K = 80
    ring = sweep.make_ring(m.n_a, K, 8, bool(masked))
    y = torch.empty((8, m.n_b, K), dtype=torch.float64, device='cuda')
    _cabi.set_tunable(5, t5)
    mode = _cabi.MODE_MASKED if masked else _cabi.MODE_FRACB
    ms, best = sweep.time_launch(lambda i: sweep.run_spmm(csr, ring, y, K, nb, mode, i, kern), reps=5, warm=2)
    sweep.report(f'C3 x{nb} masked={masked}', f'kernel={kern} t5={t5} order={order}', ms, best,
                 sweep.alg_bytes(csr, K) * nb)


if __name__ == '__main__':
    main()
