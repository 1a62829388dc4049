"""Which kernel variant moves the bf16 ADE / FDE of a golden fixture: runs the bf16 forward of every fixture under each switch
combination (one process per combination: some switches are read once by the library) and prints the relative ADE / FDE error.
    python tools/bf16_fde_probe.py [fixture ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

COMBOS = [("all new", {}), ("no fused ffn", {"TCAVP_NO_FFN_FUSED": "1"}), ("no tc head", {"TCAVP_NO_HEAD_TC": "1"}),
          ("no tcgen05 cross-attn", {"TCAVP_ATTNX_TCGEN05": "0"}),
          ("none of the three", {"TCAVP_NO_FFN_FUSED": "1", "TCAVP_NO_HEAD_TC": "1", "TCAVP_ATTNX_TCGEN05": "0"})]


def child(names):
    import torch
    from conftest import load_golden
    from test_model_gpu import _forward
    for name in names:
        fix = load_golden(name)
        m, o = _forward(fix, "bf16")
        g = fix["out"]
        ade, fde, ga, gf = float(o["ade"].mean()), float(o["fde"].mean()), float(g["ade"].mean()), float(g["fde"].mean())
        derr = float((o["decoded"] - g["decoded"]).abs().max())
        print(f"  {name:18s} ADE rel {abs(ade - ga) / ga:.2e}  FDE rel {abs(fde - gf) / gf:.2e}  decoded max|err| {derr:.3e}  (FDE {fde:.3f} vs {gf:.3f})", flush=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        child([a for a in sys.argv[1:] if a != "--child"])
        sys.exit(0)
    names = sys.argv[1:] or ["tiny_b6", "cfg1_b8", "cfg3l2_b16", "gqa_l2_b32", "llama32_1b_l2_b8", "cfg5_b32", "gpt2_tiny_b6", "gpt2_l2_b8"]
    for label, env in COMBOS:
        print(f"[{label}]", flush=True)
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", *names], env=e, timeout=600, check=False)
