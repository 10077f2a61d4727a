#!/bin/bash
# compute-sanitizer over the hot path: memcheck, racecheck (shared-memory hazards of the mbarrier /
# async-proxy protocols), synccheck.  Logs go to gpurun_out/sanitizer_*.log; summaries are copied to profiles/.
# Usage (GPU box): bash tools/run_sanitizers.sh [frames]
set -u
N=${1:-13000}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 20 \
      python tools/sanitize_target.py $N > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|exit " gpurun_out/sanitizer_$tool.log
done
