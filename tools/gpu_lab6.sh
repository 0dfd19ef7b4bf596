#!/bin/bash
# ncu DRAM bytes: distinct-row fill probe vs WROW gathers-only
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > /dev/null 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum
run() {  # name, kernel regex, filter
  timeout 300 ncu --metrics $M --clock-control none -k regex:$2 -s 3 -c 1 --csv --log-file gpurun_out/ncu_$1.csv tools/lab/kernel_lab /tmp/c3.bin "$3" > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_$1.csv')) if len(r)>5]
h=rows[0]
print('$1', ' '.join(f"{dict(zip(h,r))['Metric Name'][:60]}={dict(zip(h,r))['Metric Value']}" for r in rows[1:]))
PY
}
export LAB_PAD=8
run fill_distinct fill_probe "fillprobe ldg256 sum 160thr 8/SM"
run wrow_gathers_only wrow_kernel "wrow  ABL7 gathers only"
run wrow_full wrow_kernel "wrow  maxn6 24/SM epi2"
