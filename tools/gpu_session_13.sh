#!/bin/bash
# GPU session 13: PCIe overlap probe + GPU parity suite after removing the PATCH selector
mkdir -p gpurun_out
timeout 600 python tools/e2e_probe2.py 8 > gpurun_out/e2e_probe2.log 2>&1; cat gpurun_out/e2e_probe2.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
