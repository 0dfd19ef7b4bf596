#!/bin/bash
# round 2, session 14: group size inside one launch
mkdir -p gpurun_out
: > gpurun_out/r2s14_exp.log
python tools/exp_r2.py --segs 0 --dyns 1 --mode masked --sustain 0 > /dev/null 2>&1   # builds the map cache
for g in 1 2 3 4 6 8; do
timeout 600 python tools/exp_r2.py --segs 0 --dyns 1 --mode masked --nbs 8,24 --group $g >> gpurun_out/r2s14_exp.log 2>&1
done
for g in 2 4 8; do
timeout 600 python tools/exp_r2.py --segs 0 --dyns 1 --mode unmasked --nbs 8,24 --group $g >> gpurun_out/r2s14_exp.log 2>&1
done
grep -v CUDAEvent gpurun_out/r2s14_exp.log
