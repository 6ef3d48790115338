set -x
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2r_mr2.log 2>&1; tail -3 gpurun_out/r2r_mr2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --breakdown --no-cpu > gpurun_out/r2r_bench2.json 2> gpurun_out/r2r_bench2.err; tail -c 300 gpurun_out/r2r_bench2.err; python -c "
import json
d=json.loads(open('gpurun_out/r2r_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['e2e'], d['verify'].get('parity_rel_err'))
for k,v in d['inputs'].items(): print(k, v)
"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --breakdown --no-cpu --no-e2e --particles uniform --inputs uniform > gpurun_out/r2r_bench2_uniform.json 2> gpurun_out/r2r_bench2_uniform.err; tail -c 300 gpurun_out/r2r_bench2_uniform.err; python -c "
import json
d=json.loads(open('gpurun_out/r2r_bench2_uniform.json').read().strip().splitlines()[-1])
print('uniform step', d['value'], d['stage_ms_per_step'], d['verify'].get('parity_rel_err'))
"
