#!/bin/bash
# GPU session 12: re-baseline after the container was re-created: parity suite, memory-pattern
# probe, C3 sweep (PBIN vs WROW), default bench line
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 tools/probes/membw_probe > gpurun_out/membw.log 2>&1; cat gpurun_out/membw.log
timeout 900 python tools/sweep.py --configs c3 > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2>&1; tail -5 gpurun_out/bench.log
