mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/f4_gputest.log 2>&1; tail -3 gpurun_out/f4_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f4_smoke.txt 2>&1; tail -2 gpurun_out/f4_smoke.txt
timeout 400 python bench.py --impl reference > gpurun_out/f4_benchref.json 2> gpurun_out/f4_benchref.err; tail -1 gpurun_out/f4_benchref.json | cut -c1-300
timeout 400 python bench.py > gpurun_out/f4_bench.json 2> gpurun_out/f4_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/f4_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('kernel_ms_per_step'), 'e2e', d['e2e']['value'], 'dgrad_resident', d.get('dgrad_resident'), 'path', d.get('roofline_path',{}).get('frac'), d['roofline']['frac'], d.get('parity',{}).get('ok'), d.get('gpu_launches'), d.get('clocks'))"
