#!/bin/bash
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > gpurun_out/lab2.log 2>&1
LAB_RENUMBER=1 timeout 600 tools/lab/kernel_lab /tmp/c3.bin "$@" >> gpurun_out/lab2.log 2>&1
cat gpurun_out/lab2.log
