set -x
timeout 600 python -m pytest tests/test_gpu_pm.py tests/test_gpu_nbody.py -x -q -m gpu > gpurun_out/r2s_pytest_pm.log 2>&1; tail -5 gpurun_out/r2s_pytest_pm.log
for mode in "" "--unfused"; do
timeout 600 python bench.py --steps 5 --warmup 3 --breakdown --no-cpu --no-e2e --inputs zeldovich $mode > gpurun_out/r2s_bench1$mode.json 2> gpurun_out/r2s_bench1$mode.err; tail -c 400 gpurun_out/r2s_bench1$mode.err
python -c "
import json
d=json.loads(open('gpurun_out/r2s_bench1$mode.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['cufft_library_ms_per_step'], d.get('fused_transfer_ifft'), d['verify'].get('parity_rel_err'), d['gpu_launches'])
"
done
