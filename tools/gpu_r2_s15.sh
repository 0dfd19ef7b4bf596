#!/bin/bash
# round 2, session 15: full GPU suite, sweeps of the other configs, full bench
mkdir -p gpurun_out
: > gpurun_out/r2s15_exp.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 >> gpurun_out/r2s15_exp.log
timeout 900 python tools/sweep.py --configs c2,c1,c4 >> gpurun_out/r2s15_exp.log 2>&1
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s15_bench.json 2> gpurun_out/r2s15_bench.err
tail -5 gpurun_out/r2s15_bench.err >> gpurun_out/r2s15_exp.log
grep -v CUDAEvent gpurun_out/r2s15_exp.log
cat gpurun_out/r2s15_bench.json
