set -x
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2c_mr2.log 2>&1; echo rc=$? >> gpurun_out/r2c_mr2.log
tail -5 gpurun_out/r2c_mr2.log
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --breakdown --no-e2e --no-cpu --no-verify --inputs zeldovich > gpurun_out/r2c_bench2_$tag.json 2> gpurun_out/r2c_bench2_$tag.err; tail -c 300 gpurun_out/r2c_bench2_$tag.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2c_bench2_$tag.json').read().strip().splitlines()[-1])
print('$tag', d['value'], d['stage_ms_per_step'])
"; }
run ov0 PMB_FFT_OVERLAP=0
run ov1c1 PMB_FFT_OVERLAP=1 PMB_FFT_OVERLAP_CTAS=1
run ov1c2 PMB_FFT_OVERLAP=1 PMB_FFT_OVERLAP_CTAS=2
run ov1c4 PMB_FFT_OVERLAP=1 PMB_FFT_OVERLAP_CTAS=4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2c_bench2_full.json 2> gpurun_out/r2c_bench2_full.err; tail -c 600 gpurun_out/r2c_bench2_full.err; cut -c1-3000 gpurun_out/r2c_bench2_full.json
