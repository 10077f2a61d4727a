mkdir -p gpurun_out
(for g in 2 3; do SDFA_ASM_GATHER=$g timeout 100 python tools/solve_time.py 2>&1 | tail -1; done) > gpurun_out/t13_times.txt
cat gpurun_out/t13_times.txt
timeout 300 python -m pytest tests/test_gpu_large_batches.py -m gpu -x -q -k "generations" 2>&1 | tail -5
