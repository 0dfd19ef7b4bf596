#!/bin/bash
# multi-GPU session: N = $1 GPUs of one box
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/multi_pytest_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/host_bw_probe.py > gpurun_out/hostbw_n$N.json 2> gpurun_out/hostbw_n$N.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/multi_pytest_n$N.log
cat gpurun_out/hostbw_n$N.json
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('N',d['n_gpus'],'value',round(d['value']),'frac',round(d['roofline']['frac'],4))
print('sharded',d['sharded'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_slice'],'dropin',d['e2e_dropin']['value'],d['e2e_dropin']['ms_per_slice'])
print('binding',d['host_binding'])
PY
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
lscpu | head -25 >> gpurun_out/topo_n$N.txt
free -g >> gpurun_out/topo_n$N.txt
