set -x
timeout 300 python -m pytest tests/test_gpu_domain.py -x -q -m gpu > gpurun_out/r2v_pytest_domain.log 2>&1; tail -3 gpurun_out/r2v_pytest_domain.log
python tools/bench_route.py > gpurun_out/r2v_route_home.json 2>/dev/null; cat gpurun_out/r2v_route_home.json
PMB_ROUTE_HOME=0 python tools/bench_route.py > gpurun_out/r2v_route_nohome.json 2>/dev/null; cat gpurun_out/r2v_route_nohome.json
for mode in "" "--no-split"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --nmesh 256 --steps 20 --warmup 3 --breakdown --no-cpu --no-e2e --inputs zeldovich $mode > gpurun_out/r2v_bench2_256$mode.json 2> gpurun_out/r2v_bench2_256$mode.err
python -c "
import json
d=json.loads(open('gpurun_out/r2v_bench2_256$mode.json').read().strip().splitlines()[-1])
print('$mode', d['value'], d['gpu_launches'], d['stage_ms_per_step'], d['verify'].get('parity_rel_err'))
"
done
