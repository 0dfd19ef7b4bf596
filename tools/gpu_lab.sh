#!/bin/bash
# GPU lab session: dump C3, run the kernel lab (args: optional variant filter)
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > gpurun_out/lab.log 2>&1
timeout 600 tools/lab/kernel_lab /tmp/c3.bin "$@" >> gpurun_out/lab.log 2>&1
cat gpurun_out/lab.log
