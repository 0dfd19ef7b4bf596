#!/bin/bash
# compute-sanitizer over a slice of the parity suite (memcheck + racecheck on the smem kernels)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "kernels_bitwise_vs_oracle_all_modes and (f64-80 or f64-1- or f32-12 or f64-81) or wrow_large or float32_result" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck.log
tail -6 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "kernels_bitwise_vs_oracle_all_modes and f64-80 and (6 or 7 or 5 or 4) or wrow_large" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck.log
tail -6 gpurun_out/racecheck.log
