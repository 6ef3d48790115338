set -x
timeout 900 python -m pytest tests/test_gpu_bin.py tests/test_gpu_window.py tests/test_gpu_pm.py -x -q -m gpu > gpurun_out/r2k_tests.log 2>&1; tail -4 gpurun_out/r2k_tests.log
timeout 900 python tools/bench_bin.py --nmesh 1024 --env "" --env "PMB_BIN_TZ=5" --env "PMB_BIN_TILE_READOUT=0" > gpurun_out/r2k_bin_1024.jsonl 2> gpurun_out/r2k_bin_1024.err; cat gpurun_out/r2k_bin_1024.jsonl; tail -3 gpurun_out/r2k_bin_1024.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests.sum --clock-control none -k regex:"pmb_k_readout_cic_tile|pmb_k_bin_unsort" -c 12 --csv --log-file gpurun_out/r2k_tile_launches.csv python tools/bench_bin.py --nmesh 1024 --reps 1 > gpurun_out/r2k_tile_ncu.log 2>&1
