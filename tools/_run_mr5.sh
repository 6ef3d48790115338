set -x
N=$1
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2i_mr$N.log 2>&1; echo rc=$? >> gpurun_out/r2i_mr$N.log; tail -4 gpurun_out/r2i_mr$N.log
for ch in 4 2 0; do
PMB_FFT_CHUNKS=$ch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$ch bench.py --gpus $N --steps 10 --warmup 3 --breakdown --no-cpu --no-e2e --inputs zeldovich > gpurun_out/r2i_bench${N}_ch$ch.json 2> gpurun_out/r2i_bench${N}_ch$ch.err; tail -c 200 gpurun_out/r2i_bench${N}_ch$ch.err; python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench${N}_ch$ch.json').read().strip().splitlines()[-1])
print('chunks $ch', d['value'], d['stage_ms_per_step'], d['verify'].get('parity_rel_err'), d['fft_transpose'])
"
done
