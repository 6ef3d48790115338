set -x
timeout 900 python -m pytest tests/test_gpu_bin.py tests/test_gpu_window.py tests/test_gpu_pm.py -x -q -m gpu > gpurun_out/r2p_tests.log 2>&1; tail -3 gpurun_out/r2p_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests.sum --clock-control none -k regex:"pmb_k_paint_cic_tile" -c 4 --csv --log-file gpurun_out/r2p_tile_launches.csv python tools/bench_bin.py --nmesh 1024 --reps 1 > gpurun_out/r2p_tile_ncu.log 2>&1
timeout 900 python tools/bench_bin.py --nmesh 1024 > gpurun_out/r2p_bin_1024.jsonl 2> gpurun_out/r2p_bin_1024.err; cut -c1-600 gpurun_out/r2p_bin_1024.jsonl
