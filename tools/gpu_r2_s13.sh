#!/bin/bash
# round 2, session 13: one launch per call (groups of 8 slices inside), self-resetting counters
mkdir -p gpurun_out
: > gpurun_out/r2s13_exp.log
timeout 600 python tools/sanity_nb.py >> gpurun_out/r2s13_exp.log 2>&1
python tools/exp_r2.py --segs 0 --dyns 1 --mode masked --sustain 0 > /dev/null 2>&1   # builds the map cache
timeout 600 python tools/exp_r2.py --segs 0 --dyns 1 --mode masked --nbs 8,16,32 >> gpurun_out/r2s13_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 0 --dyns 1 --mode masked --nbs 32 --group 4 >> gpurun_out/r2s13_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 0 --dyns 1 --mode masked --nbs 32 --group 16 >> gpurun_out/r2s13_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 0 --dyns 1 --mode unmasked --nbs 8,32 >> gpurun_out/r2s13_exp.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r2s13_exp.log
grep -v CUDAEvent gpurun_out/r2s13_exp.log
