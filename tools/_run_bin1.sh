set -x
timeout 900 python -m pytest tests/test_gpu_bin.py tests/test_gpu_window.py -x -q -m gpu > gpurun_out/r2b_bin_tests.log 2>&1; tail -15 gpurun_out/r2b_bin_tests.log
timeout 600 python tools/bench_bin.py --nmesh 512 --env "" --env "PMB_BIN_PLAIN=1" > gpurun_out/r2b_bin_512.jsonl 2> gpurun_out/r2b_bin_512.err; cat gpurun_out/r2b_bin_512.jsonl; tail -3 gpurun_out/r2b_bin_512.err
timeout 900 python tools/bench_bin.py --nmesh 1024 --env "" --env "PMB_BIN_PLAIN=1" --perm > gpurun_out/r2b_bin_1024.jsonl 2> gpurun_out/r2b_bin_1024.err; cat gpurun_out/r2b_bin_1024.jsonl; tail -3 gpurun_out/r2b_bin_1024.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_bin_launches_1024.csv python tools/bench_bin.py --nmesh 1024 --reps 1 > gpurun_out/r2b_bin_ncu.log 2>&1; tail -3 gpurun_out/r2b_bin_ncu.log
