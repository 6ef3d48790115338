#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 8" "128 8" "128 4" "32 8" "512 8" "4096 8"; do
  set -- $cfg
  PMB_CARRY_UNIT=$1 PMB_GRID_MULT=$2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed.sum --clock-control none -k regex:"pmb_k_paint_(sched|cic_carry)" -s 1 -c 1 --csv --log-file gpurun_out/ab_$1_$2.csv python bench.py --nmesh 1024 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
  echo "UNIT=$1 GRID=$2: $(grep -E 'pmb_k_' gpurun_out/ab_$1_$2.csv | awk -F'","' '{printf "%s=%s ", $(NF-2), $NF}' | tr -d '"')"
done
