#!/usr/bin/env python
"""Map loader timing: NumPy COO->CSR vs b200remap_coo_to_csr (development tool, one B200)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import mapfile, synthetic as syn
for name, mk in (('C3', syn.make_c3), ('C4', syn.make_c4)):
    m = mk()
    row, col = m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1
    rng = np.random.default_rng(0)
    perm = rng.permutation(m.S.size)            # file order of real maps is not sorted by (row, col)
    for tag, (S, r, c) in (('sorted file order', (m.S, row, col)), ('shuffled file order', (m.S[perm], row[perm], col[perm]))):
        t = time.perf_counter(); a = mapfile.coo_to_csr(S, r, c, m.n_b, m.n_a); t_cpu = time.perf_counter() - t
        mapfile.coo_to_csr_gpu(S, r, c, m.n_b, m.n_a)
        torch.cuda.synchronize(); t = time.perf_counter(); b = mapfile.coo_to_csr_gpu(S, r, c, m.n_b, m.n_a); t_gpu = time.perf_counter() - t
        ok = all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))
        print(f'{name} {tag}: n_s={S.size}  NumPy {t_cpu*1e3:.0f} ms   GPU (H2D + kernels + D2H) {t_gpu*1e3:.0f} ms   identical={ok}', flush=True)
