set -x
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --particles uniform --inputs uniform --breakdown > gpurun_out/r2e_bench1_step_uniform.json 2> gpurun_out/r2e_bench1_step_uniform.err; tail -c 600 gpurun_out/r2e_bench1_step_uniform.err; python -c "
import json
d=json.loads(open('gpurun_out/r2e_bench1_step_uniform.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['verify'], d['inputs'])
"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^pmb_k_pull$' -c 2 -o gpurun_out/r2e_pull_full python tools/bench_windows.py --n 384 --windows cic,tsc > gpurun_out/r2e_pull_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
