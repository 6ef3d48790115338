set -x
timeout 1500 python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/r2c_bench1.json 2> gpurun_out/r2c_bench1.err; tail -c 800 gpurun_out/r2c_bench1.err; cut -c1-6000 gpurun_out/r2c_bench1.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 160 --csv --log-file gpurun_out/r2c_launches_1024_1gpu.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-verify --inputs zeldovich > gpurun_out/r2c_launches.log 2>&1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest_all.log 2>&1; tail -5 gpurun_out/r2c_pytest_all.log
