#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench run, then the ncu launch list and full captures of the top kernels.
# Usage: tools/profile.sh <tag> [kernel-regex(demangled) ...]      (set LIST=0 to skip the launch list)
# Each ncu pass only runs after the same command exited 0 without ncu.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r01}; shift || true
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-}"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.err; exit 1; }
if [ "${LIST:-1}" = "1" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}.md 2>/dev/null
fi
i=0
for K in "$@"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s ${SKIP:-30} -c ${COUNT:-4} \
      -o gpurun_out/prof_${TAG}_$i -f $CMD > gpurun_out/ncu_${TAG}_$i.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_$i.log | cut -c1-200
done
ls -la gpurun_out | grep ${TAG}
