set -x
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2s_mr2.log 2>&1; tail -5 gpurun_out/r2s_mr2.log
for mode in ""; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --breakdown --no-cpu --no-e2e --inputs zeldovich $mode > gpurun_out/r2s_bench2$mode.json 2> gpurun_out/r2s_bench2$mode.err; tail -c 300 gpurun_out/r2s_bench2$mode.err
python -c "
import json
d=json.loads(open('gpurun_out/r2s_bench2$mode.json').read().strip().splitlines()[-1])
print(d['value'], d['stage_ms_per_step'], d['cufft_library_ms_per_step'], d['fused_transfer_ifft'], d['verify'].get('parity_rel_err'))
"
done
