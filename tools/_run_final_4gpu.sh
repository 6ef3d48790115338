set -x
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2t_mr4.log 2>&1; tail -3 gpurun_out/r2t_mr4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 10 --warmup 3 --breakdown --no-cpu > gpurun_out/r2t_bench4.json 2> gpurun_out/r2t_bench4.err; tail -c 300 gpurun_out/r2t_bench4.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --breakdown --no-cpu > gpurun_out/r2t_bench2.json 2> gpurun_out/r2t_bench2.err; tail -c 300 gpurun_out/r2t_bench2.err
for n in 4 2; do python -c "
import json
d=json.loads(open('gpurun_out/r2t_bench$n.json').read().strip().splitlines()[-1])
print($n, d['value'], d['stage_ms_per_step'], d['e2e']['value'], d['fused_transfer_ifft'], d['verify'].get('parity_rel_err'))
for k,v in d['inputs'].items(): print(k, v.get('paint_ms'), v.get('readout_ms'), v.get('paint_readout_frac'))
"; done
