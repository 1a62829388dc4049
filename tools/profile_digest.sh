#!/bin/bash
# One profiling session on the GPU box that stays inside gpurun's 64 MiB return limit: the plain run, the ncu launch list, the
# per-launch DRAM traffic of the dominant kernel and a `--set full` capture per listed kernel — each capture is reduced to a text
# digest (tools/ncu_summary.py) on the box and the .ncu-rep is deleted unless KEEP_REP=1 (first capture only).
# Usage: tools/profile_digest.sh <tag> "<kernel regex>[@skip]" ...     (BENCH_ARGS: extra bench.py flags, e.g. "--workload cfg5")
set -u
TAG=${1:-r02}; shift || true
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary ${BENCH_ARGS:-}"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.err; exit 1; }
if [ "${LIST:-1}" = "1" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/launches_${TAG}.csv > gpurun_out/launches_${TAG}.md 2>/dev/null
  rm -f gpurun_out/launches_${TAG}.csv
fi
if [ -n "${TRAFFIC_KERNEL:-}" ]; then
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
      -k "regex:${TRAFFIC_KERNEL}" -s ${TRAFFIC_SKIP:-300} -c ${TRAFFIC_COUNT:-100} --csv --log-file gpurun_out/traffic_${TAG}.csv $CMD > /dev/null 2>&1
  python tools/traffic_summary.py gpurun_out/traffic_${TAG}.csv ${TRAFFIC_WORKLOAD:-cfg2} gpurun_out/ncu_traffic_${TAG}.json 2>&1 | tail -1
fi
i=0
for SPEC in "$@"; do
  i=$((i+1))
  K="${SPEC%@*}"; S="30"
  case "$SPEC" in *@*) S="${SPEC##*@}";; esac
  NAME=$(echo "$K" | tr -cd 'a-zA-Z0-9_')
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s $S -c ${COUNT:-2} \
      -o gpurun_out/prof_${TAG}_${NAME} -f $CMD > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/prof_${TAG}_${NAME}.ncu-rep > gpurun_out/ncu_${TAG}_${NAME}.txt 2>&1
  if [ "${KEEP_REP:-0}" != "1" ] || [ $i -gt 1 ]; then rm -f gpurun_out/prof_${TAG}_${NAME}.ncu-rep; fi
  head -2 gpurun_out/ncu_${TAG}_${NAME}.txt | cut -c1-160
done
du -sh gpurun_out
