set -x
timeout 900 python -m pytest tests/test_gpu_pull.py -x -q -m gpu > gpurun_out/r2b_tests5.log 2>&1; tail -5 gpurun_out/r2b_tests5.log
timeout 600 python tools/bench_windows.py --n 512 --windows cic,tsc,pcs > gpurun_out/r2b_windows_512_pull2.json 2> gpurun_out/r2b_windows_512_pull2.err; cat gpurun_out/r2b_windows_512_pull2.json | cut -c1-1800
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests.sum --clock-control none -k regex:"pmb_k_pull" -c 24 --csv --log-file gpurun_out/r2b_pull2_launches_512.csv python tools/bench_windows.py --n 512 --windows cic,tsc,pcs > gpurun_out/r2b_pull2_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pmb_k_bin_scatter|pmb_k_bin_unsort|pmb_k_paint_cic32_perm|pmb_k_bin_count" -c 5 -o gpurun_out/r2b_bin_full python tools/bench_bin.py --nmesh 768 --reps 1 > gpurun_out/r2b_bin_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
