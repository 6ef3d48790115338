set -x
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2b_mr2.log 2>&1; echo rc=$? >> gpurun_out/r2b_mr2.log
tail -5 gpurun_out/r2b_mr2.log
for ov in 1 0; do
PMB_FFT_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --breakdown --no-e2e --no-cpu --inputs zeldovich > gpurun_out/r2b_bench2_ov$ov.json 2> gpurun_out/r2b_bench2_ov$ov.err
tail -c 600 gpurun_out/r2b_bench2_ov$ov.err; head -c 400 gpurun_out/r2b_bench2_ov$ov.json
done
