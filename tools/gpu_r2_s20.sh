#!/bin/bash
# round 2, session 19: vectorised entry loads in lanes_k
mkdir -p gpurun_out
: > gpurun_out/r2s20_exp.log
timeout 900 python -m pytest tests -m gpu -q -x -k "kernels_bitwise or c4_full or batched_strided or float32_result or non_finite" 2>&1 | tail -5 >> gpurun_out/r2s20_exp.log
timeout 900 python tools/sweep.py --configs c4 >> gpurun_out/r2s20_exp.log 2>&1
grep -v CUDAEvent gpurun_out/r2s20_exp.log
