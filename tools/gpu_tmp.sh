#!/bin/bash
mkdir -p gpurun_out
{
for sh in 0 1 2 3 4 5 6; do
echo "== vec=2 shape=$sh"
B200REMAP_TUNABLES="3=2,5=$sh" timeout 300 python tools/exp_r2.py --dyns 1
done
echo "== vec=4 default"
timeout 300 python tools/exp_r2.py --dyns 1
} > gpurun_out/tmp_exp.log 2>&1
cat gpurun_out/tmp_exp.log
