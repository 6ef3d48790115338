set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; tail -2 gpurun_out/r2j_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_pytest_all.log 2>&1; tail -4 gpurun_out/r2j_pytest_all.log
timeout 1500 python bench.py > gpurun_out/r2j_bench1.json 2> gpurun_out/r2j_bench1.err; tail -c 500 gpurun_out/r2j_bench1.err; python -c "
import json
d=json.loads(open('gpurun_out/r2j_bench1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['traffic'], d['verify'].get('parity_rel_err'), d['cpu_baseline']['value'], d['clocks'])
for k,v in d['inputs'].items(): print(k, v)
"
