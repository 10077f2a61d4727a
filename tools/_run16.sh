mkdir -p gpurun_out
timeout 200 python tools/profile_target.py > gpurun_out/r2e_target.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_decode_tc16|k_assemble|k_solve_tc|k_output" --launch-skip 8 --launch-count 11 -f -o gpurun_out/r2e_full python tools/profile_target.py > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
