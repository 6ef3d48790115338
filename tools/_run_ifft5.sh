set -x
timeout 600 python -m pytest tests/test_gpu_pm.py tests/test_gpu_nbody.py -x -q -m gpu > gpurun_out/r2s_pytest_pm.log 2>&1; tail -3 gpurun_out/r2s_pytest_pm.log
python tools/bench_ifft.py --no-unfused > gpurun_out/r2s_ifft_whole.json 2> gpurun_out/r2s_ifft_whole.err; tail -c 300 gpurun_out/r2s_ifft_whole.err; cat gpurun_out/r2s_ifft_whole.json
