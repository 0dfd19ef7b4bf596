#!/bin/bash
# sustained A/B of two library builds (bench without e2e / cpu baseline), alternating
for rep in 1 2; do
for v in base intnan; do
  B200REMAP_LIB=$PWD/tools/ab/lib_$v.so python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline | python -c "
import json,sys; b=json.loads(sys.stdin.read()); print('$v', 'value', round(b['value']), 'frac', round(b['roofline']['frac'],4), 'launch_ms', round(b['roofline']['launch_ms'],4), 'sm_mhz', b['clocks']['sm_mhz'], 'W', b['clocks']['power_w_max'])"
done; done
B200REMAP_LIB=$PWD/tools/ab/lib_intnan.so timeout 600 python -m pytest tests -m gpu -x -q -k "kernels_bitwise or golden or wrow" 2>&1 | tail -2
