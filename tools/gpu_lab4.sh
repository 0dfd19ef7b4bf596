#!/bin/bash
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > gpurun_out/lab4.log 2>&1
for cfg in "0 32" "0 8" "64 8" "32 8" "16 8" "128 8"; do
  set -- $cfg
  LAB_BLOCK=$1 LAB_PAD=$2 timeout 600 tools/lab/kernel_lab /tmp/c3.bin "wrow  maxn6 24/SM epi2" >> gpurun_out/lab4.log 2>&1
done
cat gpurun_out/lab4.log
