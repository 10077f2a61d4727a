# usage (on the GPU box): bash tools/sweep_solve.sh  -- sweeps the solve-schedule knobs, prints per-kernel ms
mkdir -p gpurun_out
for cfg in "1 64 1.0" "1 64 1.25" "16 64 1.25" "24 64 1.25" "24 64 1.5" "32 64 1.25" "48 64 1.25" "24 64 2.0"; do
  set -- $cfg
  echo "subtree=$1 piece=$2 pad=$3" >> gpurun_out/sweep.log
  SDFA_SUBTREE_CAP=$1 SDFA_PIECE_CAP=$2 SDFA_GROUP_PAD=$3 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms_per_step'], d['dgrad_resident']['value'])" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log
