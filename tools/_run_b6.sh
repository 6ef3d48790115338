set -x
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --particles uniform --inputs uniform --breakdown > gpurun_out/r2f_bench1_step_uniform.json 2> gpurun_out/r2f_bench1_step_uniform.err; tail -c 300 gpurun_out/r2f_bench1_step_uniform.err; python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench1_step_uniform.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['verify'].get('parity_rel_err'), d['inputs'])
"
timeout 600 python tools/bench_windows.py --n 512 --windows cic,tsc,pcs > gpurun_out/r2f_windows_512.json 2> gpurun_out/r2f_windows_512.err; python -c "
import json
d=json.load(open('gpurun_out/r2f_windows_512.json'))
for k,v in d['windows'].items(): print(k, v['paint_deterministic'])
"
