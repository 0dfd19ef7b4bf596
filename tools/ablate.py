#!/usr/bin/env python
"""Ablation timings of the WROW kernel on C3 (development tool; needs the library built with
-DB200REMAP_ABLATE, see B200REMAP_LIB; results of ablated kernels are wrong by design).
usage: ablate.py [order-modes, e.g. 0,1,2] [t5 values, e.g. 0,6]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, synthetic as syn  # noqa: E402
from tools import sweep  # noqa: E402

NAMES = {0: 'full kernel', 1: 'no division', 2: 'one add per element (no recurrence)',
         3: 'no division, no recurrence', 7: 'gathers only (no stores)', 8: 'no gathers',
         11: 'stores only (no gathers, no arithmetic)'}


def main():
    orders = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '0').split(',')]
    t5s = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else '0').split(',')]
    m = syn.make_c3()
    K = 80
    ring = sweep.make_ring(m.n_a, K, 8, True)
    y = torch.empty((8, m.n_b, K), dtype=torch.float64, device='cuda')
    for order in orders:
        _cabi.set_tunable(9, order)
        csr = sweep.device_csr(m)
        nbytes = sweep.alg_bytes(csr, K) * 8
        for t5 in t5s:
            _cabi.set_tunable(5, t5)
            for abl in (0, 3, 7, 11):
                _cabi.set_tunable(8, abl)
                ms, best = sweep.time_launch(
                    lambda i: sweep.run_spmm(csr, ring, y, K, 8, _cabi.MODE_MASKED, i, 7))
                sweep.report(f'order={order} t5={t5}', f'ABL={abl} {NAMES[abl]}', ms, best, nbytes)
        csr.close()
    _cabi.set_tunable(8, 0)


if __name__ == '__main__':
    main()
