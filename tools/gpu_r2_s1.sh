#!/bin/bash
# round 2, session 1: dynamic claiming A/B + parity subset + ncu traffic of the DYN kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "wrow or kernels_bitwise or golden or full_size" 2>&1 | tail -3 > gpurun_out/r2s1_pytest.log
timeout 900 python tools/exp_r2.py --segs 0,256,1024 --grids 0,16,20 --nbs 8,16 > gpurun_out/r2s1_exp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wrow_kernel -s 6 -c 1 -o gpurun_out/r2s1_wrow_dyn python tools/exp_r2.py --sustain 0 > gpurun_out/r2s1_ncu.log 2>&1
ncu -i gpurun_out/r2s1_wrow_dyn.ncu-rep --page raw --csv > gpurun_out/r2s1_wrow_dyn_raw.csv 2>/dev/null
tail -30 gpurun_out/r2s1_exp.log
cat gpurun_out/r2s1_pytest.log
