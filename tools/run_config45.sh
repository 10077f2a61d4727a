N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --config 5 --frames 10000 --steps 3 --warmup 1 2> gpurun_out/c5_n$N.err | tail -1 > gpurun_out/c5_n$N.json
$TR bench.py --gpus $N --config 4 --utterances 1000 --steps 1 --warmup 1 2> gpurun_out/c4_n$N.err | tail -1 > gpurun_out/c4_n$N.json
tail -c 300 gpurun_out/c5_n$N.err gpurun_out/c4_n$N.err
