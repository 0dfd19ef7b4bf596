"""Small C3 batched launches of every batch size 1..13 (and > 8 in one launch) vs the C oracle."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyremap_b200 import _cabi, mapfile, synthetic as syn
from oracle import c_oracle, remap_oracle
m = syn.make_c3(scale=0.01)
ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1, m.n_b, m.n_a)
csr = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b).on_device(0)
A = remap_oracle.build_matrix(m.S, m.row, m.col, m.n_b, m.n_a)
K = 80
lv = syn.bathymetry_levels(m.n_a, K, seed=5)
X = np.stack([syn.ocean_field(m.n_a, K, seed=9 + t, max_level=lv) for t in range(13)])
Xd = torch.from_numpy(X).cuda()
st = torch.cuda.current_stream().cuda_stream
for per in (0, 12, 16):
    _cabi.set_tunable(12, per)
    for nb in (1, 2, 3, 8, 9, 12, 13):
        for mode, thr in ((2, 0.01), (1, 0.0)):
            Y = torch.full((nb, m.n_b, K), 7.0, dtype=torch.float64, device='cuda')
            xin = Xd if mode == 2 else torch.nan_to_num(Xd, nan=1.0)
            csr.spmm(xin.data_ptr(), _cabi.F64, K, K, nb, m.n_a * K, Y.data_ptr(), K, m.n_b * K, mode, thr,
                     kernel=7, stream=st)
            torch.cuda.synchronize()
            y = Y.cpu().numpy()
            for b in (0, nb - 1):
                xb = X[b] if mode == 2 else np.nan_to_num(X[b], nan=1.0)
                ref, keep = c_oracle.remap_fused(A, m.frac_b, xb, mode, thr, want_keep=True)
                assert np.array_equal(np.isnan(y[b]), ~keep), (per, nb, mode, b)
                assert np.array_equal(y[b][keep].view(np.uint64), ref[keep].view(np.uint64)), (per, nb, mode, b)
print('sanity_nb ok')
