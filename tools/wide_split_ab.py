"""A/B of the 512 x 256 wide GEMM's MMA issue order on the K = 768 shapes of cfg2 (and one long-K shape as a control).

The routing switches are read once per process, so each configuration is a child process:
    python tools/wide_split_ab.py                  # runs every configuration below, ~1 s sustained per shape
    python tools/wide_split_ab.py --child          # one configuration (environment set by the parent)
Shapes are timed with the epilogue the model uses for them (SwiGLU for gate/up, residual + row statistics for o_proj)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONFIGS = [  # (label, env)
    ("pair (default routing)", {"TCAVP_GEMM_WIDE_K": "2048"}),
    ("pair kernel for every shape", {"TCAVP_GEMM_WIDE_K": "0"}),
    ("tmastore off, pair kernel for every shape", {"TCAVP_GEMM_WIDE_K": "0", "TCAVP_GEMM_TMASTORE": "0"}),
    ("wide k-major", {"TCAVP_GEMM_WIDE_K": "512", "TCAVP_GEMM_WIDE_SPLIT": "1"}),
    ("wide split 2", {"TCAVP_GEMM_WIDE_K": "512", "TCAVP_GEMM_WIDE_SPLIT": "2"}),
    ("wide split 3", {"TCAVP_GEMM_WIDE_K": "512", "TCAVP_GEMM_WIDE_SPLIT": "3"}),
    ("wide split 4", {"TCAVP_GEMM_WIDE_K": "512", "TCAVP_GEMM_WIDE_SPLIT": "4"}),
]
SHAPES = [  # (name, M, N, K, epilogue)
    ("gate/up", 147456, 6144, 768, "swiglu"), ("qkv", 147456, 2304, 784, "plain"), ("o_proj", 147456, 768, 768, "res"),
    ("down", 147456, 768, 3072, "res"), ("7b o", 36864, 4096, 4096, "res"), ("7b gate/up", 36864, 22016, 4096, "swiglu"),
    ("7b down", 36864, 4096, 11008, "res"),
]


def child():
    import torch
    import tcavp_b200.lib as L
    L.build()
    from tcavp_b200 import ops
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for name, M, N, K, epi in SHAPES:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        kw = {}
        if epi == "swiglu":
            out = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
            kw["act"] = ops.ACT_SWIGLU
        else:
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            if epi == "res":
                kw["residual"] = torch.randn(M, N, device=dev).bfloat16()
                kw["sumsq_out"] = torch.zeros(M, dtype=torch.int64, device=dev)
        for _ in range(3):
            ops.gemm(a, w, out, **kw)
        # correctness on sampled rows (a mis-ordered barrier shows up as garbage, a hang is caught by the caller's timeout)
        rows = torch.randint(0, M, (256,), device=dev)
        acc = a[rows].float() @ w.float().t()
        if epi == "swiglu":
            want = torch.nn.functional.silu(acc[:, 0::2]) * acc[:, 1::2]
        elif epi == "res":
            want = acc + kw["residual"][rows].float()
        else:
            want = acc
        err = float((out[rows].float() - want).abs().max())
        ok = err < 0.02 * float(want.abs().max()) + 0.05
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(a, w, out, **kw); e1.record(); torch.cuda.synchronize()
        n = max(10, int(1000.0 / e0.elapsed_time(e1)))
        e0.record()
        for _ in range(n):
            ops.gemm(a, w, out, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"  {name:8s} M{M} N{N} K{K} {epi:6s}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TF/s  err {err:.3f} {'ok' if ok else 'FAIL'}", flush=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
        sys.exit(0)
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    for rep in range(2):           # A-B-...-A-B: two passes so drift of the box shows
        for label, env in CONFIGS:
            if only and not any(o in label for o in only):
                continue
            print(f"[{label}] pass {rep}", flush=True)
            e = dict(os.environ); e.update(env)
            try:
                subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=e, timeout=240, check=False)
            except subprocess.TimeoutExpired:
                print("  TIMEOUT (hang)", flush=True)
