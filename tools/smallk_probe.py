#!/usr/bin/env python
"""Thin fields (K = 1..8) on a short-row map (C3): warp tiles of the binned view (WROW) against
the lane-per-row walk on the sliced-ELL view (SELL).  Development probe."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, mapfile, synthetic as syn  # noqa: E402
from tools.sweep import time_launch  # noqa: E402


def main():
    _cabi.set_tunable(2, 1)            # build the sliced-ELL view for this short-row map too
    for name, m in (('C3', syn.make_c3()), ('C2', syn.make_c2())):
        ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1,
                                       m.n_b, m.n_a)
        csr = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b).on_device(0)
        st = torch.cuda.current_stream().cuda_stream
        for K in (1, 2, 4, 8, 16):
            for dtype, code in ((torch.float64, _cabi.F64), (torch.float32, _cabi.F32)):
                x = torch.rand((m.n_a, K), dtype=dtype, device='cuda')
                x[torch.rand(m.n_a, device='cuda') < 0.2] = float('nan')
                y = torch.empty((m.n_b, K), dtype=torch.float64, device='cuda')
                flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
                res = {}
                for kern in (7, 8, 1):
                    def run(i):
                        flush.zero_()
                        csr.spmm(x.data_ptr(), code, K, K, 1, 0, y.data_ptr(), K, 0,
                                 _cabi.MODE_MASKED, 0.01, kernel=kern, stream=st)
                    run(0)
                    torch.cuda.synchronize()
                    ref = res.setdefault('ref', y.clone())
                    same = torch.equal(ref.view(torch.int64), y.view(torch.int64))
                    # time without the flush in the bracket
                    evs = []
                    for i in range(12):
                        flush.zero_()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        csr.spmm(x.data_ptr(), code, K, K, 1, 0, y.data_ptr(), K, 0,
                                 _cabi.MODE_MASKED, 0.01, kernel=kern, stream=st)
                        b.record()
                        evs.append((a, b))
                    torch.cuda.synchronize()
                    ts = sorted(a.elapsed_time(b) for a, b in evs)
                    res[kern] = ts[len(ts) // 2] * 1e3
                    assert same, (name, K, kern)
                print(f'{name} K={K:3d} {str(dtype)[6:]:8s} wrow {res[7]:7.1f} us   sell {res[8]:7.1f} us   '
                      f'lanes_k {res[1]:7.1f} us', flush=True)
        csr.close()


if __name__ == '__main__':
    main()
