#!/bin/bash
# GPU session 3: TMA kernel bring-up (guarded by short timeouts), then tests, sweep, bench, ncu
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_staged_pipeline or test_batched_strided" > gpurun_out/pytest_tma.log 2>&1; rc=$?; echo "tma pytest rc=$rc" >> gpurun_out/pytest_tma.log; tail -15 gpurun_out/pytest_tma.log
if [ $rc -ne 0 ]; then echo "TMA bring-up failed; stopping"; nvidia-smi > gpurun_out/nvsmi.txt 2>&1; exit 0; fi
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep.py --full 1 > gpurun_out/sweep.log 2>&1; tail -60 gpurun_out/sweep.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pbin_kernel -s 20 -c 1 -o gpurun_out/prof_pbin -f python tools/sweep.py --configs c3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
