L=$PWD/sdfa-2019_b200
(for v in lib lib_h1 lib_h2; do SDFA_LIB=$L/$v/libsdfa_b200.so timeout 100 python tools/solve_time.py 2>&1 | tail -1; done) > gpurun_out/t15_times.txt; cat gpurun_out/t15_times.txt
