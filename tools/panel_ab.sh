#!/bin/bash
# A/B of the tcgen05 GEMM tile order on one box: DRAM bytes per launch (ncu; deterministic, unlike the +-5 % launch timings of a
# power-capped part) for panel budgets / L2 hints, then the cfg3 bench for the candidates.  TCAVP_GEMM_PANEL_MB=0 is row-major order.
mkdir -p gpurun_out
timeout 120 python tools/gemm_traffic.py > /dev/null 2>&1 || { echo "gemm_traffic failed"; exit 1; }
for cfg in "0 0" "40 0" "40 1" "40 2" "28 1" "56 1" "80 1"; do
  set -- $cfg
  TCAVP_GEMM_PANEL_MB=$1 TCAVP_GEMM_L2HINT=$2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
     --clock-control none --kernel-name-base demangled -k "regex:gemm_tc" --csv --log-file gpurun_out/gt_$1_$2.csv python tools/gemm_traffic.py > /dev/null 2>&1
  python - "$1" "$2" <<'PY'
import csv, sys, collections
mb, hint = sys.argv[1:3]
lines = [l for l in open(f"gpurun_out/gt_{mb}_{hint}.csv") if l.startswith('"')]
d = collections.OrderedDict()
for r in csv.DictReader(lines):
    d.setdefault(r["ID"], {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}
out = []
for i, (k, m) in enumerate(d.items()):
    if i % 2 == 0:
        continue
    g = lambda n: m[n][0] * sc[m[n][1]]
    out.append(f"rd {g('dram__bytes_read.sum') / 1e9:.2f} wr {g('dram__bytes_write.sum') / 1e9:.2f} {g('gpu__time_duration.sum'):.0f}us")
print(f"PANEL_MB={mb} HINT={hint}: " + " | ".join(out))
PY
done
for cfg in "0 0" "40 1" "40 2" "40 1" "0 0"; do
  set -- $cfg
  TCAVP_GEMM_PANEL_MB=$1 TCAVP_GEMM_L2HINT=$2 timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/panel_cfg3.json 2> gpurun_out/panel_cfg3.err
  python -c "
import json; d=json.load(open('gpurun_out/panel_cfg3.json')); print('cfg3 PANEL_MB=$1 HINT=$2', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
