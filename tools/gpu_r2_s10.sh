#!/bin/bash
# round 2, session 10: segment length and slices per launch with dynamic claiming (wpc=4, carve 28)
mkdir -p gpurun_out
: > gpurun_out/r2s10_exp.log
python tools/exp_r2.py --segs 256 --dyns 1 --mode masked --sustain 0 > /dev/null 2>&1   # builds the map cache
export B200REMAP_TUNABLES="14=4,15=28"
timeout 600 python tools/exp_r2.py --segs 64,128,192,256,384 --dyns 1 --mode masked --wpcs 4 --carves 28 >> gpurun_out/r2s10_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --mode masked --wpcs 4 --carves 28 --nbs 2,4,6,8,12 >> gpurun_out/r2s10_exp.log 2>&1
cat gpurun_out/r2s10_exp.log
