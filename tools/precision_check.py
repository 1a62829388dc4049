"""ADE / FDE of the bf16 engine variants against the fp32 parity mode on the benchmark batch (cfg2, 1024 scenes, seeded weights)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcavp_b200 as T
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = dict(T.MODEL_PRESETS["cfg1"]); lc = T.resolve_llama(cfg["base_model_name"])
s = T.make_scenes(B, cfg["seq_len"], cfg["out_len"], l_text=128, vocab=lc["vocab_size"], seed=1234, ragged_text=False)
def run(dtype):
    m = T.MultiModalTrajectoryModel(**cfg, compute_dtype=dtype); T.deterministic_fill_(m.state_dict(), 1); m = m.cuda().eval()
    o = m.engine().forward(s["x"], s["vision"], s["polygon"], s["poly_len"], s["input_ids"], s["attention_mask"], y=s["y"], norm_stat=s["norm_stat"])
    torch.cuda.synchronize()
    return o["decoded"].float().cpu(), float(o["sum_ade"]) / B, float(o["sum_fde"]) / B
ref, a0, f0 = run("fp32")
print(f"fp32: ADE {a0:.3f} FDE {f0:.3f}")
for env in ({}, {"TCAVP_NO_ABSORB": "1"}, {"TCAVP_NO_ABSORB": "1", "TCAVP_NO_SPLIT_SMALL": "1"}, {"TCAVP_NO_ABSORB": "1", "TCAVP_NO_FUSE_RSTD": "1"}):
    for k in ("TCAVP_NO_ABSORB", "TCAVP_NO_SPLIT_SMALL", "TCAVP_NO_FUSE_RSTD"): os.environ.pop(k, None)
    os.environ.update(env)
    d, a, f = run("bf16")
    err = (d - ref).abs()
    print(f"bf16 {env}: ADE {a:.3f} ({(a-a0)/a0*100:+.2f}%) FDE {f:.3f} ({(f-f0)/f0*100:+.2f}%)  max|d| {float(err.max()):.4f} mean|d| {float(err.mean()):.5f}  ref scale {float(ref.abs().max()):.2f}")
