set -x
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_pytest_all_2gpu.log 2>&1; tail -4 gpurun_out/r2u_pytest_all_2gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --breakdown --no-cpu --no-e2e > gpurun_out/r2u_bench1.json 2> gpurun_out/r2u_bench1.err; tail -c 300 gpurun_out/r2u_bench1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2u_bench1.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['verify'].get('parity_rel_err'))
for k,v in d['inputs'].items(): print(k, v)
"
