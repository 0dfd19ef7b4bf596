#!/bin/bash
# round 2, session 8: shared-memory carveout / L1 size and warps per CTA of the WROW kernel
mkdir -p gpurun_out
: > gpurun_out/r2s8_exp.log
for mode in unmasked masked; do
timeout 900 python tools/exp_r2.py --segs 256 --dyns 1 --mode $mode --wpcs 1,4 --carves 28,44,58,0 >> gpurun_out/r2s8_exp.log 2>&1
done
timeout 900 python tools/exp_r2.py --segs 256 --dyns 1 --mode masked --wpcs 2,4 --carves 14,28 --grids 24,28,32 >> gpurun_out/r2s8_exp.log 2>&1
cat gpurun_out/r2s8_exp.log
