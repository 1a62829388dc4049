"""GPU check of the panel-major tile order of the tcgen05 GEMMs (csrc/gemm.cu: unit_to_tile) with the panel width FORCED
(TCAVP_GEMM_PANEL_FORCE, read once per process — hence a script, run by tests/test_kernels_gpu.py in a subprocess):
every kernel variant, ragged M / N, panel widths that do and do not divide the tile count, whole output compared."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcavp_b200.lib as L  # noqa: E402

L.build()
from tcavp_b200 import ops  # noqa: E402

assert int(os.environ.get("TCAVP_GEMM_PANEL_FORCE", "0")) > 0, "set TCAVP_GEMM_PANEL_FORCE"
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(11)
# (M, N, K): single-CTA tiles (M < 2048), pair tile, wide tile (K >= 2048, M >= 8192, >= 4 waves of 512 x 256 tiles)
SHAPES = [(1000, 1100, 136), (1900, 96, 64), (4096 + 77, 2304 + 40, 200), (2048, 3000, 72), (20480 + 5, 4096 + 24, 2048)]
for (M, N, K) in SHAPES:
    a = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev)
    ops.gemm(a, w, out, bias=bias)
    torch.cuda.synchronize()
    want = a.float() @ w.float().t() + bias
    bad = ~torch.isclose(out.float(), want, rtol=2e-2, atol=2e-2)
    if bool(bad.any()):
        idx = bad.nonzero()[0].tolist()
        print(f"M{M} N{N} K{K}: {int(bad.sum())} mismatches, first at {idx}: got {float(out[idx[0], idx[1]])} want {float(want[idx[0], idx[1]])}")
        sys.exit(1)
    print(f"M{M} N{N} K{K}: ok", flush=True)
print("panel_order_check ok")
