mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/f5_bench_g8.json 2> gpurun_out/f5_bench_g8.err; python -c "
import json
d=json.loads(open('gpurun_out/f5_bench_g8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('host_link_ceiling'))
for k,v in d.get('gather',{}).items():
    if isinstance(v,dict): print(k, v['value'], v['ms_per_step'], v.get('nvlink_in_gbs_per_receiver'))"
