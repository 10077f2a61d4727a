mkdir -p gpurun_out
timeout 100 python tools/solve_time.py 2>&1 | tail -1 > gpurun_out/t7_times.txt
cat gpurun_out/t7_times.txt
timeout 120 python tools/tensor_timeline.py 75600 v > gpurun_out/t7_timeline.txt 2>&1; head -8 gpurun_out/t7_timeline.txt
