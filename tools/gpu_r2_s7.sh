#!/bin/bash
# round 2, session 7: bisect the unmasked regression (loop form / plain-output store)
mkdir -p gpurun_out
: > gpurun_out/r2s7_exp.log
for v in old new OLD_LOOP NO_PLAIN BOTH; do
echo "== lib_$v" >> gpurun_out/r2s7_exp.log
B200REMAP_LIB=$PWD/tools/ab/lib_$v.so timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --mode unmasked >> gpurun_out/r2s7_exp.log 2>&1
B200REMAP_LIB=$PWD/tools/ab/lib_$v.so timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --mode masked >> gpurun_out/r2s7_exp.log 2>&1
done
cat gpurun_out/r2s7_exp.log
