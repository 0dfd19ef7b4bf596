#!/bin/bash
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > /dev/null 2>&1
export LAB_PAD=8
for rep in 1 2; do
for d in x 1; do
  if [ "$d" = "1" ]; then export LAB_ZERO_LAST=1; else unset LAB_ZERO_LAST; fi
  echo "== zero_last=$d"
  timeout 300 tools/lab/kernel_lab /tmp/c3.bin "wrow  maxn6 24/SM epi2" 2>&1 | grep -E "wrow"
done; done
