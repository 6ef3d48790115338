set -x
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu -k "test_multirank and 8" > gpurun_out/r2t_mr8.log 2>&1; tail -3 gpurun_out/r2t_mr8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 10 --warmup 3 --breakdown --no-cpu > gpurun_out/r2t_bench8.json 2> gpurun_out/r2t_bench8.err; tail -c 300 gpurun_out/r2t_bench8.err
python -c "
import json
d=json.loads(open('gpurun_out/r2t_bench8.json').read().strip().splitlines()[-1])
print(8, d['value'], d['stage_ms_per_step'], d['e2e']['value'], d['fused_transfer_ifft'], d['verify'].get('parity_rel_err'))
for k,v in d['inputs'].items(): print(k, v.get('paint_ms'), v.get('readout_ms'), v.get('paint_readout_frac'))
"
