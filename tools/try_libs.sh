# usage: bash tools/try_libs.sh  -- runs the solve cycle profile + bench with alternative builds of the library
cd "$(dirname "$0")/.."
L=sdfa-2019_b200/lib
cp $L/libsdfa_b200.so /tmp/orig.so
for v in "$@"; do
  echo "=== $v"
  cp $L/libsdfa_b200_$v.so $L/libsdfa_b200.so
  timeout 200 python tools/solve_profile.py 15360 2>&1 | grep -E "total|tasks|barrier"
  timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['kernel_ms_per_step'])"
done
cp /tmp/orig.so $L/libsdfa_b200.so
