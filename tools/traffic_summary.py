"""Mean DRAM bytes per launch from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log.
    python tools/traffic_summary.py gpurun_out/traffic_<tag>.csv [workload] [profiles/ncu_traffic.json]
With an output path, the result is merged into that JSON (bench.py reads roofline.traffic from it).
    python tools/traffic_summary.py --per-launch <csv> [label]      one line, every second launch (tools/gemm_traffic.py: steady state)"""
import csv, json, os, sys

per_launch = "--per-launch" in sys.argv
if per_launch:
    sys.argv.remove("--per-launch")
path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
iid, iname, imet, iunit, ival = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
per = {}
for r in rows[1:]:
    d = per.setdefault(r[iid], {"name": r[iname]})
    d[r[imet]] = float(r[ival].replace(",", "")) * SCALE.get(r[iunit], 1.0)
if per_launch:
    cells = [f"rd {d.get('dram__bytes_read.sum', 0) / 1e9:.2f} wr {d.get('dram__bytes_write.sum', 0) / 1e9:.2f} GB {d.get('gpu__time_duration.sum', 0):.0f} us"
             for i, d in enumerate(per.values()) if i % 2 == 1]
    print((sys.argv[2] + ": " if len(sys.argv) > 2 else "") + " | ".join(cells))
    sys.exit(0)
n = len(per)
rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in per.values()) / max(n, 1)
wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in per.values()) / max(n, 1)
us = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values()) / max(n, 1)
name = next(iter(per.values()))["name"].split("(")[0] if per else "?"
out = {"kernel": name, "launches": n, "dram_read_bytes_per_launch": round(rd), "dram_write_bytes_per_launch": round(wr),
       "dram_bytes_per_launch": round(rd + wr), "avg_us_under_ncu": round(us, 1), "source": os.path.basename(path)}
print(json.dumps(out))
if len(sys.argv) > 3:
    wl, dst = sys.argv[2], sys.argv[3]
    table = json.load(open(dst)) if os.path.exists(dst) else {}
    table[wl] = out
    json.dump(table, open(dst, "w"), indent=1)
