set -x
timeout 900 ncu --set full --import-source on --clock-control none -k regex:pmb_k_ifft -c 1 -f -o gpurun_out/r2s_ifft_ncu python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify --inputs zeldovich > gpurun_out/r2s_ifft_ncu.log 2>&1
tail -3 gpurun_out/r2s_ifft_ncu.log
ls -la gpurun_out/*.ncu-rep
