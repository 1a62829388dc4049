"""One launch per 7B-class shape of tcavp_gemm for an ncu DRAM-traffic pass (tools/panel_ab.sh):
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_tc python tools/gemm_traffic.py
Prints the algorithmic bytes (A + W + out, each once) per shape so the capture can be read against them."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcavp_b200.lib as L  # noqa: E402

L.build()
from tcavp_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [(36864, 12288, 4096), (36864, 4096, 4096), (36864, 22016, 4096), (36864, 4096, 11008)]
for (M, N, K) in SHAPES:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(2):      # the second launch is the steady-state one (W partly resident from the first, as between layers it is not)
        ops.gemm(a, w, out)
    torch.cuda.synchronize()
    print(f"M{M} N{N} K{K}: algorithmic {(M * K + N * K + M * N) * 2 / 1e9:.2f} GB", flush=True)
