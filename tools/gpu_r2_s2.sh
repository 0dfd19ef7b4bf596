#!/bin/bash
# round 2, session 2: occupancy / register variants of the dynamically scheduled WROW kernel
mkdir -p gpurun_out
timeout 900 python tools/exp_r2.py --segs 256,512 --variants 0,1,2,3,4,5 --dyns 1 > gpurun_out/r2s2_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 256 --variants 0,4 --dyns 1 --nbs 1,8 --kernels 7,6 --mode unmasked >> gpurun_out/r2s2_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 256 --variants 0,4 --dyns 1 --nbs 1 --kernels 7,6 --mode masked >> gpurun_out/r2s2_exp.log 2>&1
cat gpurun_out/r2s2_exp.log
