#!/bin/bash
# A/B of the tcgen05 GEMM tile order on one box: DRAM bytes per launch (ncu; deterministic, unlike the +-5 % launch timings of a
# power-capped part) for panel budgets / L2 hints, then the cfg3 bench for the candidates.  TCAVP_GEMM_PANEL_MB=0 is row-major order.
# Usage: tools/panel_ab.sh "<MB> <HINT>" ...      (traffic pass per pair, then cfg3 A-B-B-A of the first two pairs)
mkdir -p gpurun_out
[ $# -gt 0 ] || set -- "0 0" "40 1"
timeout 120 python tools/gemm_traffic.py > /dev/null 2>&1 || { echo "gemm_traffic failed"; exit 1; }
for cfg in "$@"; do
  read mb hint <<< "$cfg"
  TCAVP_GEMM_PANEL_MB=$mb TCAVP_GEMM_L2HINT=$hint timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
     --clock-control none --kernel-name-base demangled -k "regex:gemm_tc" --csv --log-file gpurun_out/gt_${mb}_${hint}.csv python tools/gemm_traffic.py > /dev/null 2>&1
  python tools/traffic_summary.py --per-launch gpurun_out/gt_${mb}_${hint}.csv "PANEL_MB=$mb HINT=$hint"
done
for cfg in "$1" "$2" "$2" "$1"; do
  read mb hint <<< "$cfg"
  TCAVP_GEMM_PANEL_MB=$mb TCAVP_GEMM_L2HINT=$hint timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/panel_cfg3.json 2> gpurun_out/panel_cfg3.err
  python -c "
import json; d=json.load(open('gpurun_out/panel_cfg3.json')); print('cfg3 PANEL_MB=$mb HINT=$hint', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
