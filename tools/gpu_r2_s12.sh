#!/bin/bash
# round 2, session 12: ncu of the current WROW kernel (masked x8) + bench without e2e
mkdir -p gpurun_out
python tools/exp_r2.py --segs 256 --dyns 1 --mode masked --sustain 0 > /dev/null 2>&1   # builds the map cache
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wrow_kernel -s 6 -c 1 -o /tmp/r2s12 python tools/exp_r2.py --sustain 0 --segs 256 --dyns 1 --mode masked > gpurun_out/r2s12_ncu.log 2>&1
ncu -i /tmp/r2s12.ncu-rep --page raw --csv > gpurun_out/r2s12_raw.csv 2>/dev/null
ncu -i /tmp/r2s12.ncu-rep --page source --csv --print-source sass > gpurun_out/r2s12_src.csv 2>/dev/null
ncu -i /tmp/r2s12.ncu-rep --page details > gpurun_out/r2s12_details.txt 2>/dev/null
timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2s12_bench.json 2> gpurun_out/r2s12_bench.err
cat gpurun_out/r2s12_bench.json
