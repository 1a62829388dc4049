"""Counts, per kernel of libtcavp.so, the SASS instructions that prove which hardware path it uses (B200_PROFILING.md table):
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, mma.sync -> HMMA, 256-bit global access -> .256.
    python tools/sass_summary.py > profiles/sass_<round>.md          (CPU only: cuobjdump on the built library)"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "traffic-context-augmented-vehicle-trajectory-prediction-framework-using-multimodal-llm_b200", "libtcavp.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n   # noqa: E731
PAT = OrderedDict([("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
                   ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("UTMALDG (TMA load)", r"\bUTMALDG"), ("UTMASTG (TMA store)", r"\bUTMASTG"), ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (mma.sync)", r"\bHMMA"),
                   ("LDSM (ldmatrix)", r"\bLDSM"), ("LDGSTS (cp.async)", r"\bLDGSTS"), ("LDG/STG .256", r"\b(LDG|STG)[.A-Z0-9]*\.256"),
                   ("FFMA", r"\bFFMA")])
rows, cur, cnt = [], None, Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt))
        cur, cnt = m.group(1), Counter()
        continue
    for k, p in PAT.items():
        if re.search(p, line):
            cnt[k] += 1
if cur:
    rows.append((cur, cnt))
print(f"# SASS evidence per kernel of {os.path.basename(so)} (cuobjdump -sass, sm_100a)\n")
print("| kernel | " + " | ".join(PAT) + " |")
print("|---|" + "---:|" * len(PAT))
for name, c in sorted(rows, key=lambda r: demangle(r[0])):
    d, depth, cut = demangle(name), 0, None
    for i, ch in enumerate(d):          # drop the parameter list: the first '(' outside the template brackets
        depth += ch == "<"
        depth -= ch == ">"
        if ch == "(" and depth == 0:
            cut = i
            break
    d = re.sub(r"^void ", "", d[:cut]).replace("tcavp::", "").replace("(int)", "")
    print(f"| `{d[:90]}` | " + " | ".join(str(c.get(k, 0)) for k in PAT) + " |")
