#!/bin/bash
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > /dev/null 2>&1
export LAB_PAD=8
for n in "" _p128 _p256 _el_p256; do
  [ -x tools/lab/kernel_lab$n ] || continue
  echo "== gathers policy: ${n:-default}"
  timeout 300 tools/lab/kernel_lab$n /tmp/c3.bin "wrow  " 2>&1 | grep -E "wrow  maxn6 24/SM epi2|wrow  ABL4|wrow  ABL7"
done
