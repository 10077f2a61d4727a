mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nccl_gather.py -m gpu -x -q > gpurun_out/t17_g2_tests.log 2>&1; tail -3 gpurun_out/t17_g2_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/t17_bench_g2.json 2> gpurun_out/t17_bench_g2.err; python -c "
import json
d=json.loads(open('gpurun_out/t17_bench_g2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'gather', d.get('gather'))"
