#!/bin/bash
# GPU session 11: ncu --set full of one launch of kernel $1 (C3 masked x8)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wrow_kernel|patch_kernel|pbin_kernel' -s 4 -c 1 -o gpurun_out/prof_k$1 -f python tools/ncu_probe.py $1 0 0 > gpurun_out/ncu_k$1.log 2>&1
tail -3 gpurun_out/ncu_k$1.log
