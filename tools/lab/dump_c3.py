#!/usr/bin/env python
"""Dump the C3 weight matrix (canonical CSR) and the bathymetry levels to a flat binary file for
tools/lab/kernel_lab.cu (development tool).  usage: dump_c3.py out.bin [scale]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyremap_b200 import mapfile, synthetic as syn  # noqa: E402

scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
m = syn.make_c3(scale=scale)
ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1,
                               m.n_b, m.n_a)
lv = syn.bathymetry_levels(m.n_a, 80, seed=5).astype(np.int32)
nx = int(m.dst_grid_dims[0])
with open(sys.argv[1], 'wb') as f:
    np.array([m.n_b, m.n_a, ix.size, nx], dtype=np.int64).tofile(f)
    ip.astype(np.int32).tofile(f)
    ix.astype(np.int32).tofile(f)
    d.astype(np.float64).tofile(f)
    lv.tofile(f)
print('wrote', sys.argv[1], 'n_b', m.n_b, 'n_a', m.n_a, 'nnz', ix.size, 'nx', nx)
