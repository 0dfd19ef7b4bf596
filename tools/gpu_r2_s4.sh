#!/bin/bash
# round 2, session 4: instruction diet of the WROW pass loop
mkdir -p gpurun_out
timeout 900 python tools/exp_r2.py --segs 256 --dyns 1 > gpurun_out/r2s4_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --nbs 1,8 --mode unmasked >> gpurun_out/r2s4_exp.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "golden or wrow or full_size" 2>&1 | tail -5 >> gpurun_out/r2s4_exp.log
cat gpurun_out/r2s4_exp.log
