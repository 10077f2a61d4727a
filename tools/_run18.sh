mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_b.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu_b.log 2>&1
wc -l gpurun_out/r2e_launches.csv
