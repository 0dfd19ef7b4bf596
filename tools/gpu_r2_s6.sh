#!/bin/bash
# round 2, session 6: why did the unmasked WROW launch regress?  old vs new library under ncu
mkdir -p gpurun_out
for v in old new; do
B200REMAP_LIB=$PWD/tools/ab/lib_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:wrow_kernel -s 6 -c 1 -o /tmp/r2s6_$v python tools/exp_r2.py --sustain 0 --segs 256 --dyns 1 --mode unmasked > gpurun_out/r2s6_ncu_$v.log 2>&1
ncu -i /tmp/r2s6_$v.ncu-rep --page raw --csv > gpurun_out/r2s6_${v}_raw.csv 2>/dev/null
ncu -i /tmp/r2s6_$v.ncu-rep --page source --csv --print-source sass > gpurun_out/r2s6_${v}_src.csv 2>/dev/null
done
