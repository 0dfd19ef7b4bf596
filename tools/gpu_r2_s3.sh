#!/bin/bash
# round 2, session 3: L2 prefetch variants of the dynamically scheduled WROW kernel
mkdir -p gpurun_out
timeout 900 python tools/exp_r2.py --segs 256 --pfs 0,1,2,3 --dyns 1 --grids 0,20 > gpurun_out/r2s3_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 256 --pfs 0,2,3 --dyns 1 --nbs 1,8 --mode unmasked >> gpurun_out/r2s3_exp.log 2>&1
for pf in 1 2 3; do
B200REMAP_TUNABLES="13=$pf" timeout 900 python -m pytest tests -m gpu -x -q -k "wrow or kernels_bitwise or golden or full_size" 2>&1 | tail -2 >> gpurun_out/r2s3_exp.log
done
cat gpurun_out/r2s3_exp.log
