mkdir -p gpurun_out
timeout 200 python tools/profile_target.py > gpurun_out/r2d_target.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_tc --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2d_solve python tools/profile_target.py > gpurun_out/r2d_ncu.log 2>&1
timeout 120 python tools/tensor_timeline.py 75600 > gpurun_out/r2d_timeline.txt 2>&1
timeout 300 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -2 gpurun_out/r2d_ncu.log; head -12 gpurun_out/r2d_timeline.txt; python -c "
import json
d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('kernel_ms_per_step'), d['e2e']['value'], d.get('dgrad_resident',{}).get('value'), d['roofline'])"
