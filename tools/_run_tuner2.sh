set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --breakdown --no-cpu --no-e2e --inputs zeldovich > gpurun_out/r2w_bench2.json 2> gpurun_out/r2w_bench2.err; tail -c 600 gpurun_out/r2w_bench2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2w_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['exchange'], d['stage_ms_per_step'], d['verify'].get('parity_rel_err'))
"
timeout 300 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "2" > gpurun_out/r2w_mr2.log 2>&1; tail -3 gpurun_out/r2w_mr2.log
