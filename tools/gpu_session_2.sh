#!/bin/bash
# GPU session 2: parity of the binned kernel, variant sweep, bench, ncu capture of the default
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep.py --full > gpurun_out/sweep.log 2>&1; tail -70 gpurun_out/sweep.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:binned -s 4 -c 1 -o gpurun_out/prof_binned -f python bench.py --steps 1 --warmup 1 --slices 16 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
