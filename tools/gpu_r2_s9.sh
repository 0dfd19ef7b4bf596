#!/bin/bash
# round 2, session 9: carveout / warps per CTA, one process per configuration, per-launch times
mkdir -p gpurun_out
: > gpurun_out/r2s9_exp.log
python tools/exp_r2.py --segs 256 --dyns 1 --mode masked --sustain 0 > /dev/null 2>&1   # builds the map cache
for mode in unmasked masked; do
for cfg in "1 44" "1 28" "2 28" "4 28" "4 44" "4 58"; do
set -- $cfg
timeout 300 python tools/exp_r2.py --segs 256 --dyns 1 --mode $mode --wpcs $1 --carves $2 --dump 1 >> gpurun_out/r2s9_exp.log 2>&1
done; done
cat gpurun_out/r2s9_exp.log
