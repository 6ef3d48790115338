set -x
for v in "PMB_BIN_VARIANT=0" "PMB_BIN_VARIANT=1" "PMB_BIN_VARIANT=2" "PMB_BIN_VARIANT=3" "PMB_BIN_VARIANT=0 PMB_BIN_CSTRIDE=32" "PMB_BIN_VARIANT=0 PMB_BIN_SCATTER_CTAS=8"; do
tag=$(echo "$v" | tr ' =' '__')
env $v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests.sum,lts__t_sectors.sum --clock-control none -k regex:"pmb_k_bin|pmb_k_paint|pmb_k_readout" -c 14 --csv --log-file gpurun_out/r2b_binvar_$tag.csv python tools/bench_bin.py --nmesh 512 --reps 1 > gpurun_out/r2b_binvar_$tag.log 2>&1
done
