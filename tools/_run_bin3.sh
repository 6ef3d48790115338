set -x
timeout 900 python -m pytest tests/test_gpu_bin.py tests/test_gpu_window.py -x -q -m gpu > gpurun_out/r2b_bin_tests2.log 2>&1; tail -5 gpurun_out/r2b_bin_tests2.log
i=0
for v in "PMB_BIN=1" "PMB_BIN_PLAIN=2" "PMB_BIN_TILES=262144 PMB_BIN_TZ=6" "PMB_BIN_TILES=262144 PMB_BIN_TZ=6 PMB_BIN_PLAIN=2" "PMB_BIN_TILES=1048576 PMB_BIN_TZ=5" "PMB_BIN_TILES=1048576 PMB_BIN_TZ=5 PMB_BIN_PLAIN=2"; do
i=$((i+1))
env $v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests.sum,lts__t_sectors.sum --clock-control none -k regex:"pmb_k_bin|pmb_k_paint|pmb_k_readout" -c 22 --csv --log-file gpurun_out/r2b_bin3_$i.csv python tools/bench_bin.py --nmesh 1024 --reps 1 > gpurun_out/r2b_bin3_$i.log 2>&1
echo "$v" > gpurun_out/r2b_bin3_$i.env
done
