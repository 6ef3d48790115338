set -x
timeout 600 python -m pytest tests/test_gpu_domain.py -x -q -m gpu > gpurun_out/r2s_pytest_domain.log 2>&1; tail -3 gpurun_out/r2s_pytest_domain.log
python tools/bench_route.py > gpurun_out/r2s_route_staged.json 2> gpurun_out/r2s_route_staged.err; tail -c 300 gpurun_out/r2s_route_staged.err; cat gpurun_out/r2s_route_staged.json
PMB_ROUTE_STAGED=0 python tools/bench_route.py > gpurun_out/r2s_route_strided.json 2>/dev/null; cat gpurun_out/r2s_route_strided.json
