set -x
timeout 900 python -m pytest tests/test_gpu_pull.py tests/test_gpu_bin.py -x -q -m gpu > gpurun_out/r2d_tests.log 2>&1; tail -4 gpurun_out/r2d_tests.log
timeout 600 python tools/bench_windows.py --n 512 --windows cic,tsc,pcs > gpurun_out/r2d_windows_512.json 2> gpurun_out/r2d_windows_512.err; python -c "
import json
d=json.load(open('gpurun_out/r2d_windows_512.json'))
for k,v in d['windows'].items(): print(k, v['paint_deterministic'])
"
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-verify --particles zeldovich --inputs uniform > gpurun_out/r2d_bench1_uniform.json 2> gpurun_out/r2d_bench1_uniform.err; tail -c 400 gpurun_out/r2d_bench1_uniform.err; python -c "
import json
d=json.loads(open('gpurun_out/r2d_bench1_uniform.json').read().strip().splitlines()[-1])
print(d['value']); print(json.dumps(d['inputs'],indent=0))
"
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --particles uniform --inputs uniform --breakdown > gpurun_out/r2d_bench1_step_uniform.json 2> gpurun_out/r2d_bench1_step_uniform.err; tail -c 400 gpurun_out/r2d_bench1_step_uniform.err; python -c "
import json
d=json.loads(open('gpurun_out/r2d_bench1_step_uniform.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['verify'])
"
PMB_BIN=0 timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-verify --particles uniform --inputs uniform --breakdown > gpurun_out/r2d_bench1_step_uniform_perm.json 2> gpurun_out/r2d_bench1_step_uniform_perm.err; python -c "
import json
d=json.loads(open('gpurun_out/r2d_bench1_step_uniform_perm.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'])
"
