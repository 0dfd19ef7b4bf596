#!/bin/bash
# ncu DRAM/L2 counters for WROW variants in the lab (slot order experiments)
mkdir -p gpurun_out
python tools/lab/dump_c3.py /tmp/c3.bin > /dev/null 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,l1tex__t_sector_hit_rate.pct
run() {  # name, env..., filter
  name=$1; shift
  env "$@" LAB_PAD=8 timeout 300 ncu --metrics $M --clock-control none -k regex:wrow_kernel -s 3 -c 1 --csv --log-file gpurun_out/ncu_$name.csv tools/lab/kernel_lab /tmp/c3.bin "$FILTER" > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_$name.csv')) if len(r)>5]
h=rows[0]
out=[]
for r in rows[1:]:
    d=dict(zip(h,r)); out.append(f"{d['Metric Name'][:60]}={d['Metric Value']}")
print('$name', ' '.join(out))
PY
}
FILTER="wrow  maxn6 24/SM epi2" run base LAB_SEG=4096
#FILTER="wrow  maxn6 24/SM epi2" run seg2048 LAB_SEG=2048
#FILTER="wrow  maxn6 24/SM epi2" run seg16384 LAB_SEG=16384
#FILTER="wrow  slice-major (b slowest)" run bslow LAB_SEG=4096
#FILTER="wrow  ABL4 no stores" run nostore LAB_SEG=4096
FILTER="wrow  maxn6 24/SM epi2" run block32 LAB_BLOCK=32
FILTER="wrow  maxn6 24/SM epi2" run block64 LAB_BLOCK=64
FILTER="wrow  maxn6 24/SM epi2" run block16 LAB_BLOCK=16
