#!/bin/bash
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > /dev/null 2>&1
export LAB_PAD=8
for cfg in "x 4096" "1 4096" "1 1024" "1 16384" "1 65536" "1 400000"; do
  set -- $cfg
  if [ "$1" = "1" ]; then export LAB_INTERLEAVE=1; else unset LAB_INTERLEAVE; fi
  echo "== interleave=$1 seg=$2"
  LAB_SEG=$2 timeout 300 tools/lab/kernel_lab /tmp/c3.bin "wrow  maxn6 24/SM epi2" 2>&1 | grep -E "wrow|slots"
done
