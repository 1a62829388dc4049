"""Stage-by-stage diagnostics of the CUDA forward against the golden fixtures (run on the GPU box)."""
import sys, time, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_filled_model, load_golden
import tcavp_b200.lib as L
L.build()
for name in sys.argv[1:] or ["tiny_b6", "cfg1_b8"]:
    fix = load_golden(name)
    for dt in ("fp32", "bf16"):
        try:
            m = build_filled_model(fix, dt, "cuda")
            i = fix["inputs"]
            t0 = time.time()
            o = m.engine().forward(i["x"], i["vision"], i["polygon"], i["poly_len"], i["input_ids"], i["attention_mask"], y=i["y"], norm_stat=i["norm_stat"], keep_intermediates=True)
            torch.cuda.synchronize()
            g = fix["out"]
            def d(a, b):
                a = a.float().cpu(); return f"maxabs={float((a-b).abs().max()):.3e} rel={float(((a-b).abs()/(b.abs()+1e-2)).max()):.3e}"
            n = g["final_hidden_head"].shape[0]
            print(name, dt, f"{time.time()-t0:.2f}s", "| poly", d(o["poly_emb"], g["poly_emb"]), "| enc", d(o["enc"], g["enc"]),
                  "| fh", d(o["final_hidden"][:n], g["final_hidden_head"]), "| dec", d(o["decoded"], g["decoded"]),
                  "| loss", float(o["loss"]), float(g["loss"]), "| ade", float(o["ade"].mean()), float(g["ade"].mean()), flush=True)
        except Exception as e:
            print(name, dt, "FAILED", repr(e)[:500], flush=True)
