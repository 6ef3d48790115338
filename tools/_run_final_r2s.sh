set -x
python -c 'import __graft_entry__ as g; g.build(); g.smoke()' > gpurun_out/r2t_smoke.log 2>&1; tail -2 gpurun_out/r2t_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_pytest_all.log 2>&1; tail -4 gpurun_out/r2t_pytest_all.log
timeout 1500 python bench.py --breakdown > gpurun_out/r2t_bench1.json 2> gpurun_out/r2t_bench1.err; tail -c 500 gpurun_out/r2t_bench1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2t_bench1.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['e2e'], d['roofline']['frac'], d['fused_transfer_ifft'], d['verify'].get('parity_rel_err'), d['cpu_baseline']['value'], d['clocks'])
for k,v in d['inputs'].items(): print(k, v)
"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2t_launches_1024_1gpu.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify --inputs zeldovich > gpurun_out/r2t_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pmb_k_ifft -c 1 -f -o gpurun_out/r2t_ifft_ncu python tools/bench_ifft.py --no-unfused --reps 1 > gpurun_out/r2t_ifft_ncu.log 2>&1
ls -la gpurun_out/r2t*
