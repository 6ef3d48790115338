set -x
timeout 900 python tools/bench_kernels.py --nmesh 1024 --inputs zeldovich,lattice --env "PMB_RING_SCHED=0" --env "PMB_RING_SCHED=1" > gpurun_out/r2q_kernels_ringsched.jsonl 2> gpurun_out/r2q_kernels_ringsched.err; cat gpurun_out/r2q_kernels_ringsched.jsonl; tail -2 gpurun_out/r2q_kernels_ringsched.err
timeout 600 python -m pytest tests/test_gpu_window.py tests/test_gpu_pm.py -x -q -m gpu > gpurun_out/r2q_tests.log 2>&1; tail -2 gpurun_out/r2q_tests.log
