#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kernels_bitwise or float32 or large_batches or persistent or tunables or batched_strided or midsize" 2>&1 | tail -5 > gpurun_out/tmp_pytest.log
cat gpurun_out/tmp_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --slices 64 --no-e2e --no-cpu-baseline --no-sharded > gpurun_out/tmp_bench.json 2> gpurun_out/tmp_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/tmp_bench.json'))
print('value',round(d['value']),'frac',round(d['roofline']['frac'],4))
for k,v in d.get('configs',{}).items(): print(f"{k:70s} {v['ms']*1e3:9.1f} us {v['frac']*100:5.1f}%  {v['kernel']}")
PY
