mkdir -p gpurun_out
for x in 0 1 2; do SDFA_TS_EXPERIMENT=$x timeout 100 python tools/solve_time.py 2>&1 | tail -1; done > gpurun_out/t8_times.txt
cat gpurun_out/t8_times.txt
