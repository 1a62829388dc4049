#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench run, then the ncu launch list and full captures of the top kernels.
# Usage: tools/profile.sh <tag> [kernel-regex(demangled) ...]      (set LIST=0 to skip the launch list)
# Each ncu pass only runs after the same command exited 0 without ncu.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r01}; shift || true
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary ${BENCH_ARGS:-}"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.err; exit 1; }
if [ "${LIST:-1}" = "1" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}.md 2>/dev/null
fi
# DRAM traffic of every launch of the dominant kernel in ONE timed step (bench.py's roofline.traffic = their mean)
if [ -n "${TRAFFIC_KERNEL:-}" ]; then
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
      -k "regex:${TRAFFIC_KERNEL}" -s ${TRAFFIC_SKIP:-300} -c ${TRAFFIC_COUNT:-100} --csv --log-file gpurun_out/traffic_${TAG}.csv $CMD \
      > gpurun_out/ncu_traffic_${TAG}.log 2>&1
  python tools/traffic_summary.py gpurun_out/traffic_${TAG}.csv 2>&1 | tail -3
fi
i=0
for SPEC in "$@"; do      # "<kernel regex>[@<matching launches to skip>]"
  i=$((i+1))
  K="${SPEC%@*}"; S="${SKIP:-30}"
  case "$SPEC" in *@*) S="${SPEC##*@}";; esac
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s $S -c ${COUNT:-4} \
      -o gpurun_out/prof_${TAG}_$i -f $CMD > gpurun_out/ncu_${TAG}_$i.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_$i.log | cut -c1-200
done
ls -la gpurun_out | grep ${TAG}
