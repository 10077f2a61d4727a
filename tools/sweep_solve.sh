# usage (on the GPU box): bash tools/sweep_solve.sh  -- sweeps the solve-schedule knobs, prints per-kernel ms
mkdir -p gpurun_out
for cfg in "24 64 32 1.25" "16 64 32 1.25" "12 64 32 1.25" "8 64 32 1.25" "16 64 16 1.25" "16 64 24 1.25" "16 64 48 1.25" "16 96 32 1.25" "16 48 32 1.25" "16 64 32 1.1" "16 64 32 1.5"; do
  set -- $cfg
  echo "subtree=$1 piece=$2 supernode=$3 pad=$4" >> gpurun_out/sweep.log
  SDFA_SUBTREE_CAP=$1 SDFA_PIECE_CAP=$2 SDFA_SUPERNODE_CAP=$3 SDFA_GROUP_PAD=$4 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>>gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms_per_step']['solve_ms'], d['dgrad_resident']['value'])" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log
