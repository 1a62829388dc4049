"""Calibration of the bf16 backbone bands of tests/test_model_gpu.py (run on the B200 box):
relative L2 of final_hidden / image tokens against the reference goldens for every fixture, with and without 2 % multiplicative
noise injected into the weights of one decoder layer.    python tools/bf16_error_probe.py > gpurun_out/bf16_probe.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_filled_model, load_golden  # noqa: E402
from test_model_gpu import FIXTURES, _expected_image_rows, _rel_l2  # noqa: E402


def run(fix, inject):
    m = build_filled_model(fix, "bf16", "cuda")
    if inject:
        layer = m.mllm.llama_wrapper.causal_lm().model.layers[1]
        g = torch.Generator(device="cuda").manual_seed(5)
        with torch.no_grad():
            for p in layer.parameters():
                if p.dim() == 2:
                    p.mul_(1.0 + inject * torch.randn(p.shape, generator=g, device="cuda"))
    i = fix["inputs"]
    o = m.engine().forward(i["x"], i["vision"], i["polygon"], i["poly_len"], i["input_ids"], i["attention_mask"], y=i["y"],
                           norm_stat=i["norm_stat"], keep_intermediates=True)
    torch.cuda.synchronize()
    o = {k: (v.float().cpu() if torch.is_tensor(v) else v) for k, v in o.items()}
    g = fix["out"]
    n = g["final_hidden_head"].shape[0]
    img = _expected_image_rows(m, g)
    am = g["final_hidden_absmean"]
    return dict(fh=_rel_l2(o["final_hidden"][:n], g["final_hidden_head"]), img=_rel_l2(o["image_tokens_plus_mod"][:img.shape[0]], img),
                rowmean=float(((o["final_hidden"].mean(-1) - g["final_hidden_rowmean"]).abs() / am).max()),
                absmean=float(((o["final_hidden"].abs().mean(-1) - am).abs() / am).max()),
                maxabs=float((o["final_hidden"][:n] - g["final_hidden_head"]).abs().max() / g["final_hidden_head"].abs().max()),
                dec=float((o["decoded"] - g["decoded"]).abs().max()),
                ade=abs(float(o["ade"].mean()) - float(g["ade"].mean())) / float(g["ade"].mean()),
                fde=abs(float(o["fde"].mean()) - float(g["fde"].mean())) / float(g["fde"].mean()))


if __name__ == "__main__":
    for name in FIXTURES:
        fix = load_golden(name)
        for inject in (0.0, 0.01, 0.02):
            r = run(fix, inject)
            print(name, f"inject={inject}", " ".join(f"{k}={v:.3e}" for k, v in r.items()), flush=True)
