#!/bin/bash
# round 2, session 11: real-tile items + static fills; sanitizer on large batches; full GPU suite
mkdir -p gpurun_out
: > gpurun_out/r2s11_exp.log
timeout 600 python tools/sanity_nb.py >> gpurun_out/r2s11_exp.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanity_nb.py 2>&1 | tail -15 >> gpurun_out/r2s11_exp.log
python tools/exp_r2.py --segs 256 --dyns 1 --mode masked --sustain 0 > /dev/null 2>&1   # builds the map cache
for mode in masked unmasked; do
timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --mode $mode >> gpurun_out/r2s11_exp.log 2>&1
done
B200REMAP_LIB=$PWD/tools/ab/lib_old.so timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --mode masked >> gpurun_out/r2s11_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --mode masked --nbs 1 --kernels 7,6 >> gpurun_out/r2s11_exp.log 2>&1
timeout 600 python tools/exp_r2.py --segs 256 --dyns 1 --mode unmasked --nbs 1 --kernels 7,6 >> gpurun_out/r2s11_exp.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 >> gpurun_out/r2s11_exp.log
grep -v CUDAEvent gpurun_out/r2s11_exp.log
