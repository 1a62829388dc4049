"""Per-kernel gradient parity of the fine-tune step: every backward C-ABI entry point against torch autograd of the same
op evaluated in fp32 on the CPU (the plain-tensor restatement the oracle uses).  fp32 storage: rtol 1e-4; bf16 storage:
compared on bf16-rounded inputs with tolerances that cover output rounding only."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restated as R  # noqa: E402

DEV = "cuda"


@pytest.fixture(scope="module")
def ops(lib_built):
    from tcavp_b200 import ops as o
    assert torch.cuda.is_available()
    return o


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _td(name):
    return torch.float32 if name == "fp32" else torch.bfloat16


def _tol(name):
    return dict(rtol=1e-4, atol=1e-5) if name == "fp32" else dict(rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("batch,rows,cols", [(1, 70, 33), (1, 1000, 64), (5, 2, 15), (3, 64, 40)])
def test_transpose(ops, dtype, batch, rows, cols):
    x = _rand(batch, rows, cols, seed=1).to(_td(dtype))
    ldo = (rows + 7) // 8 * 8
    out = torch.zeros(batch, cols, ldo, dtype=torch.float32, device=DEV)
    ops.transpose(x.to(DEV), out, rows=rows, cols=cols, ldo=ldo, batch=batch, in_bstride=rows * cols, out_bstride=cols * ldo)
    torch.testing.assert_close(out.cpu()[:, :, :rows], x.float().transpose(1, 2))
    assert torch.count_nonzero(out[:, :, rows:]) == 0


@pytest.mark.parametrize("batch,rows,cols", [(1, 70, 34), (1, 1000, 64), (2, 144, 1536), (3, 65, 130), (1, 7, 2)])
def test_transpose_bf16_tiles(ops, batch, rows, cols):
    """bf16 -> bf16 (the 64 x 64-tile kernel with two-element accesses): exact, ragged tiles, odd row counts, padded output rows untouched."""
    x = _rand(batch, rows, cols, seed=2).bfloat16()
    ldo = (rows + 7) // 8 * 8
    out = torch.full((batch, cols, ldo), 5.0, dtype=torch.bfloat16, device=DEV)
    ops.transpose(x.to(DEV), out, rows=rows, cols=cols, ldo=ldo, batch=batch, in_bstride=rows * cols, out_bstride=cols * ldo)
    assert torch.equal(out.cpu()[:, :, :rows], x.transpose(1, 2))
    assert bool((out[:, :, rows:] == 5.0).all())


@pytest.mark.parametrize("period", [1, 7, 16])
def test_period_sum(ops, period):
    rows, cols = period * 37, 200
    x = _rand(rows, cols, seed=2)
    out = torch.ones(period, cols, device=DEV)
    ops.period_sum(x.to(DEV), out, rows=rows, cols=cols, period=period)
    torch.testing.assert_close(out.cpu(), 1.0 + x.view(37, period, cols).sum(0), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_relu_bwd_axpby_swiglu(ops, dtype):
    td = _td(dtype)
    rows, cols = 129, 72
    y, dy = _rand(rows, cols, seed=3).to(td), _rand(rows, cols, seed=4).to(td)
    dx = torch.empty(rows, cols, dtype=td, device=DEV)
    ops.relu_bwd(dy.to(DEV), y.to(DEV), dx, rows=rows, cols=cols)
    torch.testing.assert_close(dx.cpu().float(), torch.where(y.float() > 0, dy.float(), torch.zeros(())))
    out = torch.empty(rows, cols, dtype=torch.float32, device=DEV)
    ops.axpby(y.to(DEV), out, rows=rows, cols=cols, alpha=0.5, b=dy.to(DEV), beta=-2.0)
    torch.testing.assert_close(out.cpu(), 0.5 * y.float() - 2.0 * dy.float())
    I = 40
    gu = _rand(rows, 2 * I, seed=5).to(td).float().requires_grad_(True)
    want = torch.nn.functional.silu(gu[:, 0::2]) * gu[:, 1::2]
    d = _rand(rows, I, seed=6).to(td)
    want.backward(d.float())
    got = torch.empty(rows, I, dtype=td, device=DEV)
    ops.swiglu(gu.detach().to(td).to(DEV), got, rows=rows, I=I)
    torch.testing.assert_close(got.cpu().float(), want.detach(), **_tol(dtype))
    dgu = torch.empty(rows, 2 * I, dtype=td, device=DEV)
    ops.swiglu_bwd(d.to(DEV), gu.detach().to(td).to(DEV), dgu, rows=rows, I=I)
    torch.testing.assert_close(dgu.cpu().float(), gu.grad, **_tol(dtype))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,cols,ld,with_add", [(300, 768, 768, True), (37, 128, 136, True), (65, 1024, 1024, False), (9, 2048, 2048, True),
                                                   (5, 40, 40, True), (130, 256, 264, False)])
def test_layernorm_bwd_dx_frozen(ops, dtype, rows, cols, ld, with_add):
    """dx-only LayerNorm backward (frozen LayerNorm of a GPT-2-arch backbone) with the residual-branch gradient added in the same pass:
    register path (bf16, <= 1024 columns), generic path, padded row strides — against autograd through F.layer_norm."""
    td = _td(dtype)
    x = _rand(rows, cols, seed=14).to(td).float().requires_grad_(True)
    w = 1.0 + 0.1 * _rand(cols, seed=15)
    b = 0.1 * _rand(cols, seed=16)
    y = torch.nn.functional.layer_norm(x, (cols,), w, b, 1e-5)
    dy = _rand(rows, cols, seed=17).to(td)
    y.backward(dy.float())
    add = _rand(rows, cols, seed=18).to(td) if with_add else None
    want = x.grad + (add.float() if with_add else 0.0)

    def pad(t):
        o = torch.zeros(rows, ld, dtype=td, device=DEV)
        o[:, :cols] = t.to(td).to(DEV)
        return o
    dx = torch.full((rows, ld), 3.0, dtype=td, device=DEV)
    ops.layernorm_bwd_dx(pad(dy), pad(x.detach()), w.to(DEV), dx, rows=rows, cols=cols, eps=1e-5, add=pad(add) if with_add else None,
                         lddy=ld, ldx=ld, ldadd=ld, lddx=ld)
    tol = dict(rtol=1e-4, atol=2e-5) if dtype == "fp32" else dict(rtol=2e-2, atol=3e-2)
    torch.testing.assert_close(dx[:, :cols].cpu().float(), want, **tol)
    assert bool((dx[:, cols:] == 3.0).all())


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,cols,ld", [(129, 72, 72), (300, 3072, 3072), (33, 40, 48), (17, 37, 37)])
def test_gelu_new_forward_and_backward(ops, dtype, rows, cols, ld):
    """HF ACT2FN["gelu_new"] as its own pass + backward on the stored pre-activation (GPT-2 fine-tune step), vector and scalar paths,
    padded rows — against torch's tanh GELU and autograd through it."""
    td = _td(dtype)
    x = (2.5 * _rand(rows, cols, seed=12)).to(td).float().requires_grad_(True)
    want = torch.nn.functional.gelu(x, approximate="tanh")
    dy = _rand(rows, cols, seed=13).to(td)
    want.backward(dy.float())
    xd = torch.zeros(rows, ld, dtype=td, device=DEV)
    xd[:, :cols] = x.detach().to(td).to(DEV)
    out = torch.full((rows, ld), 7.0, dtype=td, device=DEV)
    ops.gelu_tanh(xd, out, rows=rows, cols=cols, ldx=ld, ldo=ld)
    torch.testing.assert_close(out[:, :cols].cpu().float(), want.detach(), **_tol(dtype))
    assert bool((out[:, cols:] == 7.0).all())
    dyd = torch.zeros(rows, ld, dtype=td, device=DEV)
    dyd[:, :cols] = dy.to(DEV)
    ops.gelu_tanh_bwd(dyd, xd, dyd, rows=rows, cols=cols, lddy=ld, ldx=ld, lddx=ld)          # in place on the gradient, as the step uses it
    torch.testing.assert_close(dyd[:, :cols].cpu().float(), x.grad, **_tol(dtype))


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,cols,with_res", [(37, 64, True), (300, 768, False), (5, 40, True)])
def test_layernorm_bwd(ops, dtype, rows, cols, with_res):
    td = _td(dtype)
    x = _rand(rows, cols, seed=7).to(td).float().requires_grad_(True)
    r = _rand(rows, cols, seed=8).to(td).float() if with_res else None
    w = (1.0 + 0.1 * _rand(cols, seed=9)).requires_grad_(True)
    b = (0.1 * _rand(cols, seed=10)).requires_grad_(True)
    y = R.layer_norm(x + r if with_res else x, w, b)
    dy = _rand(rows, cols, seed=11).to(td)
    y.backward(dy.float())
    dx = torch.empty(rows, cols, dtype=td, device=DEV)
    dw, db = torch.zeros(cols, device=DEV), torch.zeros(cols, device=DEV)
    ops.layernorm_bwd(dy.to(DEV), x.detach().to(td).to(DEV), w.detach().to(DEV), residual=None if r is None else r.to(td).to(DEV),
                      dx=dx, dw=dw, db=db)
    torch.testing.assert_close(dx.cpu().float(), x.grad, **_tol(dtype))
    torch.testing.assert_close(dw.cpu(), w.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(db.cpu(), b.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("unit_w", [True, False])
def test_rmsnorm_bwd(ops, dtype, unit_w):
    td = _td(dtype)
    rows, cols, ld = 77, 128, 136
    xb = _rand(rows, ld, seed=12).to(td)
    x = xb[:, :cols].float().requires_grad_(True)
    w = torch.ones(cols) if unit_w else 1.0 + 0.1 * _rand(cols, seed=13)
    y = R.rms_norm(x, w, 1e-6)
    dy = _rand(rows, cols, seed=14).to(td)
    add = _rand(rows, cols, seed=15).to(td)
    y.backward(dy.float())
    dx = torch.empty(rows, cols, dtype=td, device=DEV)
    ops.rmsnorm_bwd(dy.to(DEV), xb.to(DEV), dx, rows=rows, cols=cols, eps=1e-6, w=None if unit_w else w.to(DEV), add=add.to(DEV), ldx=ld)
    torch.testing.assert_close(dx.cpu().float(), x.grad + add.float(), **_tol(dtype))


@pytest.mark.parametrize("kind", ["rms", "rms_unit_w", "ln"])
@pytest.mark.parametrize("rows,cols,ldx,with_add", [(32768 + 19, 768, 784, True), (33000, 512, 512, False), (32768, 1024, 1024, True), (40001, 256, 264, True)])
def test_norm_backward_over_many_bf16_rows_through_the_bulk_copy_ring(ops, kind, rows, cols, ldx, with_add):
    """>= 32768 bf16 rows: RMSNorm backward and the dX-only LayerNorm backward run through norm_bwd_pipe_kernel (rows of x / dy / add
    staged by cp.async.bulk, three row triples in flight per warp) — against autograd, with a padded x stride (the K-extended operand
    of the Llama path), row counts that leave ring slots empty, with and without the residual-branch gradient."""
    g = torch.Generator().manual_seed(31)
    xb = torch.zeros(rows, ldx)
    xb[:, :cols] = torch.randn(rows, cols, generator=g) * 1.5 + 0.25 * torch.randn(rows, 1, generator=g)
    xb = xb.bfloat16()
    x = xb[:, :cols].float().requires_grad_(True)
    w = torch.ones(cols) if kind == "rms_unit_w" else 1.0 + 0.1 * _rand(cols, seed=32)
    y = torch.nn.functional.layer_norm(x, (cols,), w, 0.1 * _rand(cols, seed=33), 1e-5) if kind == "ln" else R.rms_norm(x, w, 1e-6)
    dy = torch.randn(rows, cols, generator=g).bfloat16()
    y.backward(dy.float())
    add = torch.randn(rows, cols, generator=g).bfloat16() if with_add else None
    want = x.grad + (add.float() if with_add else 0.0)
    dx = torch.empty(rows, cols, dtype=torch.bfloat16, device=DEV)
    if kind == "ln":
        ops.layernorm_bwd_dx(dy.to(DEV), xb.to(DEV), w.to(DEV), dx, rows=rows, cols=cols, eps=1e-5, add=add.to(DEV) if with_add else None, ldx=ldx)
    else:
        ops.rmsnorm_bwd(dy.to(DEV), xb.to(DEV), dx, rows=rows, cols=cols, eps=1e-6, w=None if kind == "rms_unit_w" else w.to(DEV),
                        add=add.to(DEV) if with_add else None, ldx=ldx)
    torch.testing.assert_close(dx.cpu().float(), want, rtol=2e-2, atol=3e-2)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_rope_adjacent_matches_fused_gemm_layout_and_inverts(ops, dtype):
    td = _td(dtype)
    L, dh, heads, B = 12, 32, 3, 4
    rows, cols = B * L, heads * dh
    table = ops.rope_table(L, dh, 10000.0, DEV, layout=1)
    x = _rand(rows, cols + 8, seed=16).to(td)
    buf = x.clone().to(DEV)
    ops.rope_adjacent_(buf, rows=rows, L=L, ld=cols + 8, cols=cols, dh=dh, table=table)
    # reference: HF rotate_half on the un-permuted layout; adjacent pair (2i, 2i+1) of a head = HF pair (i, i + dh/2)
    cos, sin = R.rope_cos_sin(L, dh, 10000.0)
    xh = x[:, :cols].float().view(B, L, heads, dh)
    unperm = torch.cat([xh[..., 0::2], xh[..., 1::2]], dim=-1)                      # adjacent -> HF order
    rot = unperm * cos[None, :, None, :] + R.rotate_half(unperm) * sin[None, :, None, :]
    want = torch.stack([rot[..., : dh // 2], rot[..., dh // 2:]], dim=-1).reshape(B, L, heads, dh).reshape(rows, cols)
    torch.testing.assert_close(buf.cpu().float()[:, :cols], want, **_tol(dtype))
    torch.testing.assert_close(buf.cpu()[:, cols:], x[:, cols:])
    if dtype == "fp32":
        ops.rope_adjacent_(buf, rows=rows, L=L, ld=cols + 8, cols=cols, dh=dh, table=table, inverse=True)
        torch.testing.assert_close(buf.cpu(), x, rtol=1e-5, atol=1e-5)


def test_copy_rows_scatter_gather(ops):
    B, Q, L, H = 3, 4, 10, 16
    img = _rand(B * Q, H, seed=17)
    fused = torch.zeros(B * L, H, device=DEV)
    ops.copy_rows(img.to(DEV), fused, rows=B * Q, cols=H, out_remap=(Q, L, 0))
    assert torch.equal(fused.cpu().view(B, L, H)[:, :Q].reshape(B * Q, H), img)
    back = torch.empty(B * (L - Q), H, dtype=torch.bfloat16, device=DEV)
    ops.copy_rows(fused, back, rows=B * (L - Q), cols=H, in_remap=(L - Q, L, Q))
    assert torch.count_nonzero(back) == 0
    back2 = torch.empty(B * Q, H, device=DEV)
    ops.copy_rows(fused, back2, rows=B * Q, cols=H, in_remap=(Q, L, 0))
    assert torch.equal(back2.cpu(), img)


def test_masked_mean_bwd(ops):
    B, P, D = 5, 16, 8
    lens = [0, 16, 1, 7, 3]
    x = _rand(B, P, D, seed=18).requires_grad_(True)
    valid = (torch.arange(P)[None, :] < torch.tensor(lens)[:, None]).float()[:, :, None]
    out = (x * valid).sum(1) / torch.tensor(lens).clamp(min=1)[:, None]
    d = _rand(B, D, seed=19)
    out.backward(d)
    dx = torch.empty(B, P, D, device=DEV)
    ops.masked_mean_bwd(d.to(DEV), torch.tensor(lens, dtype=torch.int32, device=DEV), dx, B=B, P=P, D=D)
    torch.testing.assert_close(dx.cpu(), x.grad)


@pytest.mark.parametrize("T_in,T_out", [(6, 12), (15, 25), (15, 15)])
def test_nlinear_bwd(ops, T_in, T_out):
    B, C = 9, 64
    x = _rand(B, T_in, C, seed=20).requires_grad_(True)
    w = _rand(T_out, T_in, C, seed=21, scale=0.3).requires_grad_(True)          # [t][s][c]
    last = x[:, -1:, :]
    out = torch.einsum("tsc,bsc->btc", w, x - last) + last
    g = _rand(B, T_out, C, seed=22)
    out.backward(g)
    din = torch.empty(B, T_in, C, device=DEV)
    dw = torch.zeros(T_out, T_in, C, device=DEV)
    ops.nlinear_bwd(g.to(DEV), B=B, C=C, T_in=T_in, T_out=T_out, x_in=x.detach().to(DEV), w=w.detach().to(DEV), din=din, dw=dw)
    torch.testing.assert_close(din.cpu(), x.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dw.cpu(), w.grad, rtol=1e-4, atol=1e-5)
    # forward kernel agrees with the same formula (no lane adjust, zero bias)
    dec = torch.empty(B, T_out, C, device=DEV)
    ops.nlinear_decode(x.detach().to(DEV), w.detach().to(DEV), torch.zeros(T_out, C, device=DEV), None, dec, B=B, C=C, T_in=T_in, T_out=T_out)
    torch.testing.assert_close(dec.cpu(), out.detach(), rtol=1e-4, atol=1e-5)


def test_head_assemble_and_loss_bwd(ops):
    B, T_in, T_out = 7, 6, 12
    o = _rand(B, T_out, 2, seed=23).requires_grad_(True)
    x, y = torch.rand(B, 2, T_in), torch.rand(B, 2, T_out)
    ns = [(10.0 * b, 10.0 * b + 100.0 + b, 700.0, 705.0 + 9 * b) for b in range(B)]
    dec = o.permute(0, 2, 1) + x[:, :, -1:]
    loss = R.mse_loss(dec, y, ns)
    loss.backward(torch.tensor(0.5))
    decoded = torch.empty(B, 2, T_out, device=DEV)
    ops.head_assemble(o.detach().reshape(B * T_out, 2).to(DEV), x.to(DEV), decoded, B=B, T_in=T_in, T_out=T_out)
    torch.testing.assert_close(decoded.cpu(), dec.detach())
    d_o = torch.empty(B * T_out, 2, device=DEV)
    ops.traj_loss_bwd(decoded, y.to(DEV), torch.tensor(ns, device=DEV), d_o, B=B, T_out=T_out, gscale=torch.tensor([0.5], device=DEV))
    # the oracle de-normalises both operands before subtracting (fp32 cancellation near min_y ~ 700); the kernel uses (dec - y) * range^2
    torch.testing.assert_close(d_o.cpu().view(B, T_out, 2), o.grad, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("J", [4, 16, 32])
def test_skinny_dw(ops, dtype, J):
    td = _td(dtype)
    M, N, ld = 1000, 300, 312
    Y = _rand(M, ld, seed=24).to(td)
    Z = _rand(M, J + 8, seed=25).to(td)
    rs = torch.rand(M) + 0.5
    out = torch.zeros(N, J, device=DEV)
    ops.skinny_dw(Y.to(DEV), Z.to(DEV), out, M=M, N=N, J=J, ldy=ld, ldz=J + 8, row_scale=rs.to(DEV))
    want = (Y[:, :N].float() * rs[:, None]).t() @ Z[:, :J].float()
    torch.testing.assert_close(out.cpu(), want, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("J", [16, 32])
def test_skinny_dw_tensor_core_path(ops, J):
    """bf16 operands with 16-byte aligned rows take the mma.sync kernel (ldmatrix.trans on the row-major operands); the row factor is
    applied to Z in bf16 on its way into shared memory, so the reference rounds the same product."""
    M, N, ld = 5000, 304, 312
    Y = _rand(M, ld, seed=24).bfloat16()
    Z = _rand(M, J + 8, seed=25).bfloat16()
    rs = torch.rand(M) + 0.5
    out = torch.zeros(N, J, device=DEV)
    ops.skinny_dw(Y.to(DEV), Z.to(DEV), out, M=M, N=N, J=J, ldy=ld, ldz=J + 8, row_scale=rs.to(DEV))
    want = Y[:, :N].float().t() @ (Z[:, :J].float() * rs[:, None]).bfloat16().float()
    torch.testing.assert_close(out.cpu(), want, rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("dtype,M,N,K", [("fp32", 1000, 200, 70), ("fp32", 3000, 64, 2), ("bf16", 1000, 200, 72), ("bf16", 4100, 384, 256),
                                         ("bf16", 700, 130, 36), ("mixed", 900, 64, 64)])
def test_dw(ops, dtype, M, N, K):
    """out[n][k] += sum_m dy[m][n] x[m][k] on the row-major operands (no transposes): FFMA kernel (fp32 / mixed / ragged) and
    tensor-core kernel (bf16, N % 8 == K % 8 == 0)."""
    tdy = torch.float32 if dtype in ("fp32", "mixed") else torch.bfloat16
    tdx = torch.float32 if dtype == "fp32" else torch.bfloat16
    dy = _rand(M, N + 8, seed=31).to(tdy)
    x = _rand(M, K + 16, seed=32).to(tdx)
    out = torch.zeros(N, K + 3, device=DEV)
    out[:, K:] = 7.0
    ops.dw(dy.to(DEV)[:, :N], x.to(DEV)[:, :K], out, M=M, N=N, K=K)
    want = dy[:, :N].float().t() @ x[:, :K].float()
    torch.testing.assert_close(out[:, :K].cpu(), want, rtol=1e-3, atol=2e-3)
    assert (out[:, K:] == 7.0).all()
    ops.dw(dy.to(DEV)[:, :N], x.to(DEV)[:, :K], out, M=M, N=N, K=K)          # accumulates
    torch.testing.assert_close(out[:, :K].cpu(), 2 * want, rtol=1e-3, atol=4e-3)


ATTN_CASES = [  # B, H, Hkv, Tq, Tk, dh, causal, masked
    (3, 4, 2, 24, 24, 32, True, True),     # LLM-like: causal + padding + GQA
    (2, 12, 12, 144, 144, 64, True, True),
    (2, 2, 2, 25, 40, 384, False, False),  # LTSF cross-attention: wide heads, no mask
    (4, 4, 4, 64, 64, 16, False, True),    # lane polygon encoder
    (3, 8, 8, 16, 15, 96, False, False),   # Q-Former cross-attention
    (2, 2, 2, 15, 15, 32, False, False),   # temporal self-attention
]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("B,H,Hkv,Tq,Tk,dh,causal,masked", ATTN_CASES)
def test_attention_bwd(ops, dtype, B, H, Hkv, Tq, Tk, dh, causal, masked):
    td = _td(dtype)
    q = _rand(B, Tq, H, dh, seed=26).to(td).float().requires_grad_(True)
    k = _rand(B, Tk, Hkv, dh, seed=27).to(td).float().requires_grad_(True)
    v = _rand(B, Tk, Hkv, dh, seed=28).to(td).float().requires_grad_(True)
    valid = torch.ones(B, Tk, dtype=torch.bool)
    if masked:
        for b in range(B):
            valid[b, Tk - 1 - 2 * b:] = False
    scale = dh ** -0.5
    kk = k.repeat_interleave(H // Hkv, dim=2)
    vv = v.repeat_interleave(H // Hkv, dim=2)
    s = torch.einsum("bihd,bjhd->bhij", q, kk) * scale
    allow = valid[:, None, None, :].expand(B, H, Tq, Tk)
    if causal:
        allow = allow & torch.tril(torch.ones(Tq, Tk, dtype=torch.bool))[None, None]
    s = s.masked_fill(~allow, float("-inf"))
    o = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, dim=-1), vv)
    do = _rand(B, Tq, H, dh, seed=29).to(td)
    o.backward(do.float())
    qd, kd, vd, dod = (t.detach().to(td).to(DEV).contiguous() for t in (q, k, v, do))
    dq = torch.empty_like(qd)
    dk = torch.zeros(B, Tk, Hkv, dh, device=DEV)
    dv = torch.zeros(B, Tk, Hkv, dh, device=DEV)
    km = valid.to(torch.int32).to(DEV) if masked else None
    ops.attention_bwd(qd, kd, vd, dod, dq, dk, dv, B=B, H=H, Hkv=Hkv, Tq=Tq, Tk=Tk, dh=dh, q_strides=(Tq * H * dh, H * dh),
                      k_strides=(Tk * Hkv * dh, Hkv * dh), v_strides=(Tk * Hkv * dh, Hkv * dh), do_strides=(Tq * H * dh, H * dh),
                      dq_strides=(Tq * H * dh, H * dh), dk_strides=(Tk * Hkv * dh, Hkv * dh), dv_strides=(Tk * Hkv * dh, Hkv * dh),
                      scale=scale, causal=causal, key_mask=km)
    tol = dict(rtol=1e-3, atol=1e-4) if dtype == "fp32" else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dq.cpu().float(), q.grad, **tol)
    torch.testing.assert_close(dk.cpu(), k.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(dv.cpu(), v.grad, rtol=1e-3, atol=1e-4)


def test_adamw_matches_torch(ops):
    n = 10_001
    p0, g = _rand(n, seed=30), _rand(n, seed=31)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=5e-4, weight_decay=1e-4)
    p, m, v = p0.clone().to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        gs = g * step
        ref.grad = gs.clone()
        opt.step()
        ops.adamw_(p, (gs * 2.0).to(DEV), m, v, lr=5e-4, weight_decay=1e-4, step=step, grad_scale=0.5)
    torch.testing.assert_close(p.cpu(), ref.detach(), rtol=1e-5, atol=1e-6)


TC_ATTN_CASES = [  # B, H, Tq, Tk, dh, causal, masked   (bf16, H == Hkv, forward output supplied -> tensor-core kernel)
    (2, 12, 144, 144, 64, True, True),     # LLM cfg-1/2
    (2, 4, 144, 144, 128, True, True),     # LLM 7B-class head
    (3, 4, 24, 24, 32, True, True),
    (4, 4, 64, 64, 16, False, True),       # lane polygon encoder
    (3, 8, 16, 15, 96, False, False),      # Q-Former cross-attention
    (3, 8, 15, 15, 96, False, False),
    (2, 2, 15, 15, 32, False, False),      # temporal self-attention
    (1, 2, 200, 256, 64, False, True),
    (2, 3, 256, 256, 64, True, False),
]


@pytest.mark.parametrize("B,H,Tq,Tk,dh,causal,masked", TC_ATTN_CASES)
def test_attention_bwd_tensor_core(ops, B, H, Tq, Tk, dh, causal, masked):
    td = torch.bfloat16
    q = _rand(B, Tq, H, dh, seed=40).to(td).float().requires_grad_(True)
    k = _rand(B, Tk, H, dh, seed=41).to(td).float().requires_grad_(True)
    v = _rand(B, Tk, H, dh, seed=42).to(td).float().requires_grad_(True)
    valid = torch.ones(B, Tk, dtype=torch.bool)
    if masked:
        for b in range(B):
            valid[b, Tk - 1 - 3 * b:] = False
    scale = dh ** -0.5
    s = torch.einsum("bihd,bjhd->bhij", q, k) * scale
    allow = valid[:, None, None, :].expand(B, H, Tq, Tk)
    if causal:
        allow = allow & torch.tril(torch.ones(Tq, Tk, dtype=torch.bool))[None, None]
    s = s.masked_fill(~allow, float("-inf"))
    o = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, dim=-1), v)
    do = _rand(B, Tq, H, dh, seed=43).to(td)
    o.backward(do.float())
    qd, kd, vd, dod, od = (t.detach().to(td).to(DEV).contiguous() for t in (q, k, v, do, o))
    dq = torch.full_like(qd, float("nan"))
    dk = torch.zeros(B, Tk, H, dh, device=DEV)
    dv = torch.zeros(B, Tk, H, dh, device=DEV)
    km = valid.to(torch.int32).to(DEV) if masked else None
    sq, sk = (Tq * H * dh, H * dh), (Tk * H * dh, H * dh)
    n0 = ops.launch_count()
    ops.attention_bwd(qd, kd, vd, dod, dq, dk, dv, B=B, H=H, Hkv=H, Tq=Tq, Tk=Tk, dh=dh, q_strides=sq, k_strides=sk, v_strides=sk,
                      do_strides=sq, dq_strides=sq, dk_strides=sk, dv_strides=sk, scale=scale, causal=causal, key_mask=km, o=od, o_strides=sq)
    assert ops.launch_count() - n0 == 1
    # bf16 P / dS operands inside the kernel: tolerance scaled to each tensor's magnitude
    for name, got, want in (("dq", dq.float().cpu(), q.grad), ("dk", dk.cpu(), k.grad), ("dv", dv.cpu(), v.grad)):
        err = float((got - want).abs().max()) / (float(want.abs().max()) + 1e-12)
        assert err < 2e-2, (name, err)
        assert torch.isfinite(got).all(), name


@pytest.mark.parametrize("B,H,Tq,Tk,dh,masked", [(3, 2, 25, 144, 384, False), (2, 2, 50, 144, 384, True), (2, 2, 12, 40, 2048, False),
                                                 (2, 2, 64, 250, 128, True)])
@pytest.mark.parametrize("dkv", ["bf16", "fp32"])
def test_attention_bwd_few_queries_wide_heads(ops, B, H, Tq, Tk, dh, masked, dkv):
    """tcavp_attention_bwd_owned on the LTSF cross-attention shape (few queries, head_dim = H/2): tensor-core kernel, one CTA per
    (batch, head), dq / dk / dv stored directly (dk / dv in the caller's dtype, here straight into a packed [K | V] gradient buffer)."""
    td = torch.bfloat16
    q = _rand(B, Tq, H, dh, seed=50).to(td).float().requires_grad_(True)
    k = _rand(B, Tk, H, dh, seed=51).to(td).float().requires_grad_(True)
    v = _rand(B, Tk, H, dh, seed=52).to(td).float().requires_grad_(True)
    valid = torch.ones(B, Tk, dtype=torch.bool)
    if masked:
        for b in range(B):
            valid[b, Tk - 1 - 5 * b:] = False
    scale = dh ** -0.5
    s = (torch.einsum("bihd,bjhd->bhij", q, k) * scale).masked_fill(~valid[:, None, None, :], float("-inf"))
    o = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, dim=-1), v)
    do = _rand(B, Tq, H, dh, seed=53).to(td)
    o.backward(do.float())
    qd, kd, vd, dod = (t.detach().to(td).to(DEV).contiguous() for t in (q, k, v, do))
    dq = torch.full_like(qd, float("nan"))
    E = H * dh
    dkv_buf = torch.full((B * Tk, 2 * E), float("nan"), device=DEV, dtype=torch.bfloat16 if dkv == "bf16" else torch.float32)
    km = valid.to(torch.int32).to(DEV) if masked else None
    sq, sk = (Tq * E, E), (Tk * E, E)
    assert ops.attention_bwd_owned_ok(qd, H=H, Hkv=H, Tq=Tq, Tk=Tk, dh=dh, o=None)
    n0 = ops.launch_count()
    ops.attention_bwd_owned(qd, kd, vd, dod, dq, dkv_buf, dkv_buf[:, E:], B=B, H=H, Tq=Tq, Tk=Tk, dh=dh, q_strides=sq, k_strides=sk, v_strides=sk,
                            do_strides=sq, dq_strides=sq, dk_strides=(Tk * 2 * E, 2 * E), dv_strides=(Tk * 2 * E, 2 * E), scale=scale, key_mask=km)
    assert ops.launch_count() - n0 == 1
    got_k = dkv_buf[:, :E].float().cpu().view(B, Tk, H, dh)
    got_v = dkv_buf[:, E:].float().cpu().view(B, Tk, H, dh)
    for name, got, want in (("dq", dq.float().cpu(), q.grad), ("dk", got_k, k.grad), ("dv", got_v, v.grad)):
        assert torch.isfinite(got).all(), name
        err = float((got - want).abs().max()) / (float(want.abs().max()) + 1e-12)
        assert err < 2e-2, (name, err)


@pytest.mark.parametrize("M,I,H", [(4300, 384, 256), (300, 64, 128), (9000, 1024, 2112)])
def test_swiglu_backward_fused_into_the_gemm(ops, M, I, H):
    """act = SWIGLU_BWD: d(mid) = dx . W_down (the GEMM) never leaves the tile; the epilogue turns it into interleaved (d gate, d up) from the
    stashed (gate, up) pairs — against autograd through silu(gate) * up followed by the down projection."""
    dx = _rand(M, H, seed=60, scale=0.5).bfloat16()
    wdT = _rand(I, H, seed=61, scale=H ** -0.5).bfloat16()          # W_down^T : [I, H]  (d mid = dx . W_down = dx @ wdT.T)
    gu = _rand(M, 2 * I, seed=62).bfloat16()
    g = gu[:, 0::2].float().requires_grad_(True)
    u = gu[:, 1::2].float().requires_grad_(True)
    mid = torch.nn.functional.silu(g) * u
    dmid = dx.float() @ wdT.float().t()
    mid.backward(dmid)
    out = torch.empty(M, 2 * I, dtype=torch.bfloat16, device=DEV)
    ops.gemm(dx.to(DEV), wdT.to(DEV), out, act=ops.ACT_SWIGLU_BWD, aux_out=gu.to(DEV))
    got = out.float().cpu()
    torch.testing.assert_close(got[:, 0::2], g.grad, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(got[:, 1::2], u.grad, rtol=3e-2, atol=3e-2)
