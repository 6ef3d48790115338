set -x
python tools/bench_ifft.py > gpurun_out/r2s_ifft_wide.json 2> gpurun_out/r2s_ifft_wide.err; tail -c 300 gpurun_out/r2s_ifft_wide.err; cat gpurun_out/r2s_ifft_wide.json
PMB_IFFT_NARROW=1 python tools/bench_ifft.py --no-unfused > gpurun_out/r2s_ifft_narrow.json 2>&1; cat gpurun_out/r2s_ifft_narrow.json
PMB_IFFT_NARROW=1 PMB_IFFT_CTAS=1 python tools/bench_ifft.py --no-unfused > gpurun_out/r2s_ifft_narrow1.json 2>&1; cat gpurun_out/r2s_ifft_narrow1.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pmb_k_ifft -c 1 -f -o gpurun_out/r2s_ifft_ncu3 python tools/bench_ifft.py --no-unfused --reps 1 > gpurun_out/r2s_ifft_ncu3.log 2>&1
