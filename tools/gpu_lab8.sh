#!/bin/bash
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > /dev/null 2>&1
export LAB_PAD=8
for n in "" _el2 _el1el2 _na_el2 _ef2; do
  echo "== gathers policy: ${n:-default}"
  timeout 300 tools/lab/kernel_lab$n /tmp/c3.bin "w" 2>&1 | grep -E "wrow  maxn6 24/SM epi2|wrow  ABL4|wrow  ABL7|wpatch 8x4 maxn6 20"
done
