import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from conftest import build_filled_model, load_golden
fix = load_golden(sys.argv[1] if len(sys.argv) > 1 else "gqa_l2_b32")
g = fix["out"]; i = fix["inputs"]
for env in ({}, {"TCAVP_NO_ABSORB": "1"}, {"TCAVP_NO_ABSORB": "1", "TCAVP_NO_SPLIT_SMALL": "1"}, {"TCAVP_NO_ABSORB": "1", "TCAVP_NO_FUSE_RSTD": "1"}):
    for k in ("TCAVP_NO_ABSORB", "TCAVP_NO_SPLIT_SMALL", "TCAVP_NO_FUSE_RSTD"): os.environ.pop(k, None)
    os.environ.update(env)
    m = build_filled_model(fix, "bf16", "cuda")
    o = m.engine().forward(i["x"], i["vision"], i["polygon"], i["poly_len"], i["input_ids"], i["attention_mask"], y=i["y"], norm_stat=i["norm_stat"], keep_intermediates=True)
    torch.cuda.synchronize()
    d = o["decoded"].float().cpu()
    fh = o["final_hidden"].float().cpu()
    n = g["final_hidden_head"].shape[0]
    print(env, "ADE %.3f (%.3f) FDE %.3f (%.3f)" % (float(o["ade"].mean()), float(g["ade"].mean()), float(o["fde"].mean()), float(g["fde"].mean())),
          "max|d| %.5f mean|d| %.6f" % (float((d - g["decoded"]).abs().max()), float((d - g["decoded"]).abs().mean())),
          "fh rel err %.4f" % float((fh[:n] - g["final_hidden_head"]).abs().mean() / g["final_hidden_head"].abs().mean()),
          "mean signed d last step %.6f" % float((d[:, :, -1] - g["decoded"][:, :, -1]).mean()))
