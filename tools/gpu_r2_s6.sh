#!/bin/bash
# round 2, session 6: why did the unmasked WROW launch regress?  old vs new library under ncu
mkdir -p gpurun_out
for v in old new; do
B200REMAP_LIB=$PWD/tools/ab/lib_$v.so timeout 600 ncu --set full --clock-control none -k regex:wrow_kernel -s 6 -c 1 -o gpurun_out/r2s6_$v python tools/exp_r2.py --sustain 0 --segs 256 --dyns 1 --mode unmasked > gpurun_out/r2s6_ncu_$v.log 2>&1
ncu -i gpurun_out/r2s6_$v.ncu-rep --page raw --csv > gpurun_out/r2s6_${v}_raw.csv 2>/dev/null
done
B200REMAP_LIB=$PWD/tools/ab/lib_old.so timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --mode unmasked > gpurun_out/r2s6_exp.log 2>&1
B200REMAP_LIB=$PWD/tools/ab/lib_new.so timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --mode unmasked >> gpurun_out/r2s6_exp.log 2>&1
cat gpurun_out/r2s6_exp.log
