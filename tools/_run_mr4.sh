set -x
N=$1
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2h_mr$N.log 2>&1; echo rc=$? >> gpurun_out/r2h_mr$N.log; tail -4 gpurun_out/r2h_mr$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --breakdown --no-cpu > gpurun_out/r2h_bench$N.json 2> gpurun_out/r2h_bench$N.err; tail -c 300 gpurun_out/r2h_bench$N.err; python -c "
import json
d=json.loads(open('gpurun_out/r2h_bench$N.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['e2e'], d['verify'].get('parity_rel_err'), d['fft_transpose'])
for k,v in d['inputs'].items(): print(k, v)
"
PMB_FFT_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --breakdown --no-cpu --no-e2e --no-verify --inputs zeldovich > gpurun_out/r2h_bench${N}_ov0.json 2> gpurun_out/r2h_bench${N}_ov0.err; python -c "
import json
d=json.loads(open('gpurun_out/r2h_bench${N}_ov0.json').read().strip().splitlines()[-1])
print('overlap off', d['value'], d['stage_ms_per_step'])
"
