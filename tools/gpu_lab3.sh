#!/bin/bash
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > gpurun_out/lab3.log 2>&1
for seg in 4096 1024 512; do
  LAB_SEG=$seg timeout 600 tools/lab/kernel_lab /tmp/c3.bin "$@" >> gpurun_out/lab3.log 2>&1
done
cat gpurun_out/lab3.log
