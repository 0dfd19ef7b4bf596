#!/bin/bash
# GPU session 9: parity suite incl. the WROW kernel, C3 sweep
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep.py --configs ${1:-c3} > gpurun_out/sweep.log 2>&1; cat gpurun_out/sweep.log
