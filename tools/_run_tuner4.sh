set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 5 --warmup 3 --breakdown --no-cpu --no-e2e --inputs zeldovich > gpurun_out/r2w_bench4.json 2> gpurun_out/r2w_bench4.err
python -c "
import json
d=json.loads(open('gpurun_out/r2w_bench4.json').read().strip().splitlines()[-1])
print(d['value'], d['exchange']['mode'], d['exchange']['measured_ms'], d['stage_ms_per_step'], d['verify'].get('parity_rel_err'))
"
