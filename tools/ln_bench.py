"""LayerNorm over the GPT-2 token stream (147 456 x 768 bf16) and poly_embed at the cfg5 shape: device-event timings.
    TCAVP_LN_PIPE=0 python tools/ln_bench.py      # register-resident kernel only (A/B against the bulk-copy ring)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcavp_b200.lib as L  # noqa: E402

L.build()
from tcavp_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("TCAVP_LN_PIPE =", os.environ.get("TCAVP_LN_PIPE", "1"))
for rows, cols in ((147456, 768), (73728, 768), (147456, 512), (147456, 1024), (16384, 768), (36864, 768)):
    x = torch.randn(rows, cols, device=dev).bfloat16()
    w, b = torch.ones(cols, device=dev), torch.zeros(cols, device=dev)
    out = torch.empty_like(x)
    us = timed(lambda: ops.layernorm(x, w, b, out))
    print(f"layernorm {rows} x {cols} bf16: {us:7.1f} us  {rows * cols * 4 / us / 1e3:7.0f} GB/s", flush=True)
    c = torch.empty_like(x)
    us = timed(lambda: ops.cast(x, c, rows=rows, cols=cols))
    print(f"   (copy of the same bytes: {us:7.1f} us  {rows * cols * 4 / us / 1e3:7.0f} GB/s)", flush=True)
print("TCAVP_NORM_BWD_PIPE =", os.environ.get("TCAVP_NORM_BWD_PIPE", "1"))
for rows, cols, ld in ((73728, 768, 784), (73728, 768, 768)):
    x = torch.randn(rows, ld, device=dev).bfloat16()
    dy, add, dx = (torch.randn(rows, cols, device=dev).bfloat16() for _ in range(3))
    w = torch.ones(cols, device=dev)
    us = timed(lambda: ops.rmsnorm_bwd(dy, x, dx, rows=rows, cols=cols, eps=1e-6, add=add, ldx=ld))
    print(f"rmsnorm_bwd {rows} x {cols} (ldx {ld}) bf16, 4 streams: {us:7.1f} us  {rows * cols * 8 / us / 1e3:7.0f} GB/s", flush=True)
    us = timed(lambda: ops.layernorm_bwd_dx(dy, x, w, dx, rows=rows, cols=cols, eps=1e-5, add=add, ldx=ld))
    print(f"layernorm_bwd_dx {rows} x {cols} (ldx {ld}) bf16, 4 streams: {us:7.1f} us  {rows * cols * 8 / us / 1e3:7.0f} GB/s", flush=True)
B, P, D = 4096, 48, 64
poly = torch.rand(B, P, 2, device=dev) * 1000
lens = torch.full((B,), 40, dtype=torch.int32, device=dev)
w, b, pos = torch.randn(D, 2, device=dev), torch.randn(D, device=dev), torch.randn(64, D, device=dev)
out = torch.empty(B * P, D, device=dev)
km = torch.empty(B, P, dtype=torch.int32, device=dev)
us = timed(lambda: ops.poly_embed(poly, lens, w, b, pos, out, km, B=B, P=P, D=D))
print(f"poly_embed {B} x {P} x {D}: {us:6.1f} us  {B * P * D * 4 / us / 1e3:6.0f} GB/s (output bytes)")
