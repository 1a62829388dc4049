#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench run, then the ncu launch list and full captures of the top kernels.
# Usage: tools/profile.sh <tag> [kernel-regex ...]
set -u
TAG=${1:-r01}; shift || true
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.err; exit 1; }
# launch list of the last (timed) step: 3 warm-up forwards + 1 timed + 2 e2e warm-ups + 1 e2e step
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
for K in "$@"; do
  NAME=$(echo "$K" | tr -c 'A-Za-z0-9\n' '_')
  ncu --set full --clock-control none --import-source on -k regex:$K -s 60 -c 6 -o gpurun_out/prof_${TAG}_${NAME} -f $CMD > gpurun_out/ncu_${TAG}_${NAME}.log 2>&1
done
ls -la gpurun_out | tail -20
