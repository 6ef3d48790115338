set -x
timeout 900 python -m pytest tests/test_gpu_bin.py tests/test_gpu_window.py tests/test_gpu_pm.py tests/test_gpu_nbody.py -x -q -m gpu > gpurun_out/r2n_tests.log 2>&1; tail -3 gpurun_out/r2n_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --particles uniform --inputs uniform --breakdown > gpurun_out/r2n_bench1_step_uniform.json 2> gpurun_out/r2n_bench1_step_uniform.err; tail -c 300 gpurun_out/r2n_bench1_step_uniform.err; python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench1_step_uniform.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['verify'].get('parity_rel_err')); print(d['inputs'])
"
timeout 900 python tools/bench_bin.py --nmesh 1024 > gpurun_out/r2n_bin_1024.jsonl 2> gpurun_out/r2n_bin_1024.err; cut -c1-600 gpurun_out/r2n_bin_1024.jsonl
