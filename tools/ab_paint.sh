#!/bin/bash
mkdir -p gpurun_out
for cfg in "5 1" "13 1" "5 0" "13 0"; do
  set -- $cfg
  PMB_DBG=$1 PMB_SCHED=$2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed.sum --clock-control none -k regex:"pmb_k_paint_sched" -s 1 -c 1 --csv --log-file gpurun_out/ab_$1_$2.csv python bench.py --nmesh 1024 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
  echo "DBG=$1 SCHED=$2: $(grep -E 'pmb_k_' gpurun_out/ab_$1_$2.csv | awk -F'","' '{printf "%s=%s ", $(NF-2), $NF}' | tr -d '"')"
done
