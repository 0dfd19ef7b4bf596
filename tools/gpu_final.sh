#!/bin/bash
# final evidence: full GPU suite, bench, launch list + full ncu set of the dominant kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/final_pytest.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --slices 64 --no-e2e --no-cpu-baseline --no-configs --no-sharded > gpurun_out/final_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wrow_kernel -s 12 -c 1 -o /tmp/final python bench.py --steps 1 --warmup 1 --slices 64 --no-e2e --no-cpu-baseline --no-configs --no-sharded > gpurun_out/final_ncu_full.log 2>&1
ncu -i /tmp/final.ncu-rep --page raw --csv > gpurun_out/final_ncu_raw.csv 2>/dev/null
ncu -i /tmp/final.ncu-rep --page details > gpurun_out/final_ncu_details.txt 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:wrow_kernel -s 20 -c 1 -o /tmp/c2 python tools/sweep.py --configs c2 > gpurun_out/final_ncu_c2.log 2>&1
ncu -i /tmp/c2.ncu-rep --page raw --csv > gpurun_out/final_ncu_raw_c2.csv 2>/dev/null
cat gpurun_out/final_pytest.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench.json'))
print('value',round(d['value']),'frac',round(d['roofline']['frac'],4),'launch_ms',round(d['roofline']['launch_ms'],4),'clocks',d['clocks'])
print('sharded',d['sharded']['value'],d['sharded']['sharded_parity'])
print('e2e',d['e2e']['ms_per_slice'],'dropin',d['e2e_dropin']['ms_per_slice'],d['e2e_dropin']['ms_per_slice_previous_result_still_held'])
for k,v in d.get('configs',{}).items(): print(f"{k:70s} {v['ms']*1e3:9.1f} us {v['frac']*100:5.1f}%  {v['kernel']}")
PY
