#!/bin/bash
# GPU session 16: slot padding 8: parity suite, C3 sweep, lab segment-size check
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep.py --configs c3,c2 > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
python tools/lab/dump_c3.py /tmp/c3.bin > gpurun_out/lab5.log 2>&1
for seg in 2048 4096 8192 16384 65536; do
  LAB_SEG=$seg LAB_PAD=8 timeout 600 tools/lab/kernel_lab /tmp/c3.bin "wrow  maxn6 24/SM epi2" >> gpurun_out/lab5.log 2>&1
done
grep -E "^# seg|wrow" gpurun_out/lab5.log
