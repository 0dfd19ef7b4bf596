#!/bin/bash
# round 2, session 17: ROWTILE kernel, lane rule, pinned result pool, host bandwidth probe (1 GPU)
mkdir -p gpurun_out
: > gpurun_out/r2s17_exp.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 >> gpurun_out/r2s17_exp.log
timeout 900 python tools/sweep.py --configs c4 >> gpurun_out/r2s17_exp.log 2>&1
timeout 300 python tools/host_bw_probe.py > gpurun_out/r2s17_hostbw_n1.json 2>> gpurun_out/r2s17_exp.log
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s17_bench.json 2> gpurun_out/r2s17_bench.err
tail -5 gpurun_out/r2s17_bench.err >> gpurun_out/r2s17_exp.log
grep -v CUDAEvent gpurun_out/r2s17_exp.log
cat gpurun_out/r2s17_hostbw_n1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s17_bench.json'))
print('frac',d['roofline']['frac'],'launch_ms',d['roofline']['launch_ms'])
print('e2e',d['e2e']['ms_per_slice'],'dropin',d['e2e_dropin']['ms_per_slice'])
for k,v in d['configs'].items(): print(f"{k:70s} {v['ms']*1e3:9.1f} us {v['frac']*100:5.1f}%  {v['kernel']}")
PY
