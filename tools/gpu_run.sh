#!/bin/bash
# generic GPU session: full GPU test-suite, then the bench (what the driver runs at round end)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/run_pytest.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/run_bench.json 2> gpurun_out/run_bench.err
cat gpurun_out/run_pytest.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/run_bench.json'))
print('value',round(d['value']),'frac',round(d['roofline']['frac'],4),'launch_ms',round(d['roofline']['launch_ms'],4),'clocks',d['clocks'])
print('sharded',d['sharded']['value'],d['sharded']['sharded_parity'])
print('e2e',d['e2e']['ms_per_slice'],'dropin',d['e2e_dropin']['ms_per_slice'],d['e2e_dropin']['ms_per_slice_previous_result_still_held'])
for k,v in d.get('configs',{}).items(): print(f"{k:70s} {v['ms']*1e3:9.1f} us {v['frac']*100:5.1f}%  {v['kernel']}")
print('cpu',d.get('cpu_baseline',{}).get('value'))
PY
