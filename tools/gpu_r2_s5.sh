#!/bin/bash
# round 2, session 5: frac_b in the slot record; full GPU suite; other configs
mkdir -p gpurun_out
timeout 900 python tools/exp_r2.py --segs 256 --dyns 1 > gpurun_out/r2s5_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --nbs 1,8 --kernels 7,6 --mode unmasked >> gpurun_out/r2s5_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --nbs 1 --kernels 7,6 --mode masked >> gpurun_out/r2s5_exp.log 2>&1
timeout 900 python tools/sweep.py --configs c2,c1,c4 >> gpurun_out/r2s5_exp.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r2s5_exp.log
cat gpurun_out/r2s5_exp.log
