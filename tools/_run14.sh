mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t14_gputest.log 2>&1; tail -3 gpurun_out/t14_gputest.log
timeout 300 python bench.py > gpurun_out/t14_bench.json 2> gpurun_out/t14_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/t14_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('kernel_ms_per_step'), 'e2e', d['e2e']['value'], 'dgrad_resident', d.get('dgrad_resident'), 'path', d.get('roofline_path'))"
timeout 300 python bench.py --config 5 > gpurun_out/t14_c5.json 2> gpurun_out/t14_c5.err; tail -1 gpurun_out/t14_c5.json | cut -c1-900
