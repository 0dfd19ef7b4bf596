#!/bin/bash
# GPU session 14: DMA-runs H2D path: new tests + e2e probe (gather vs dma)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "copy_runs or streamed or host" ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/e2e_probe2.py 8 > gpurun_out/e2e_probe2.log 2>&1; grep -v "blocks=" gpurun_out/e2e_probe2.log
