#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/exp_r2.py --dyns 1 --pfs 0,1
timeout 300 python tools/exp_r2.py --dyns 1 --pfs 1 --mode unmasked
} > gpurun_out/tmp_exp.log 2>&1
cat gpurun_out/tmp_exp.log
