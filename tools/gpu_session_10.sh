#!/bin/bash
# GPU session 10: ncu counters for PBIN and the WROW variants (slot orders 0 = BFS patches, 2 = binned)
mkdir -p gpurun_out
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_op_read_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for cfg in "6 0 0" "7 6 2" "7 0 2" "7 6 0" "7 0 0"; do
  tag=$(echo $cfg | tr ' ' '_')
  python tools/ncu_probe.py $cfg 2>&1 | tail -1
  timeout 300 ncu --metrics $M --clock-control none -k regex:'pbin_kernel|wrow_kernel' -s 4 -c 1 --csv --log-file gpurun_out/probe_$tag.csv python tools/ncu_probe.py $cfg > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/probe_$tag.csv')) if len(r)>5]
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r)); print('   ', d['Metric Name'], d['Metric Unit'], d['Metric Value'])
PY
done
