"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
    python tools/launch_summary.py gpurun_out/launches_TAG.csv > profiles/launches_TAG.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live event timings, not absolutes."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(int\)", "", name)
    name = name.split("(")[0] if not name.startswith("at::") else name.split("<")[0]
    return name[:90]


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], ns))
    total = sum(r[3] for r in rows) or 1.0
    groups = OrderedDict()
    for name, grid, block, ns in rows:
        g = groups.setdefault(name, dict(n=0, ns=0.0, mx=0.0, grids=set(), block=block))
        g["n"] += 1
        g["ns"] += ns
        g["mx"] = max(g["mx"], ns)
        g["grids"].add(grid)
    print(f"# ncu launch list summary: {path}\n")
    print(f"{len(rows)} launches, {total / 1e6:.3f} ms summed kernel time (serialised, cold cache)\n")
    print("| kernel | launches | total ms | share | avg us | max us | block | grids |")
    print("|---|---:|---:|---:|---:|---:|---|---|")
    for name, g in sorted(groups.items(), key=lambda kv: -kv[1]["ns"]):
        grids = sorted(g["grids"])
        gs = ", ".join(grids[:3]) + (" …" if len(grids) > 3 else "")
        print(f"| `{name}` | {g['n']} | {g['ns'] / 1e6:.3f} | {g['ns'] / total:.4f} | {g['ns'] / g['n'] / 1e3:.1f} | {g['mx'] / 1e3:.1f} | {g['block']} | {gs} |")
    ours = sum(g["ns"] for n, g in groups.items() if not n.startswith("at::"))
    print(f"\nlibtcavp kernels: {ours / total:.4f} of summed kernel time; the rest is torch memset/copy plumbing (allocation zero-fill, H2D staging).")


if __name__ == "__main__":
    main(sys.argv[1])
