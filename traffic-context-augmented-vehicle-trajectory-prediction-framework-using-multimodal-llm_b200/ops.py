"""Thin torch-tensor front end of the C ABI in include/tcavp.h.

torch is used for device memory and streams only: every function here hands raw device pointers to
libtcavp.so and raises TcavpError on a non-zero return code.  Nothing falls back to torch math."""
import ctypes
from ctypes import POINTER, Structure, byref, c_float, c_int, c_longlong, c_void_p

import torch

from . import lib as _lib

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_SWIGLU = 0, 1, 2
_DT = {torch.float32: F32, torch.bfloat16: BF16}


class GemmArgs(Structure):
    _fields_ = [("M", c_int), ("N", c_int), ("K", c_int),
                ("A", c_void_p), ("lda", c_int),
                ("W", c_void_p), ("ldw", c_int),
                ("in_dtype", c_int),
                ("out", c_void_p), ("ldo", c_int), ("out_dtype", c_int),
                ("bias", c_void_p),
                ("residual", c_void_p), ("ldr", c_int), ("res_dtype", c_int),
                ("act", c_int),
                ("remap_gi", c_int), ("remap_go", c_int), ("remap_off", c_int),
                ("rope_cos_sin", c_void_p), ("rope_L", c_int), ("rope_dh", c_int), ("rope_cols", c_int),
                ("row_scale", c_void_p)]


class AttnArgs(Structure):
    _fields_ = [("B", c_int), ("H", c_int), ("Hkv", c_int), ("Tq", c_int), ("Tk", c_int), ("dh", c_int),
                ("q", c_void_p), ("q_sb", c_longlong), ("q_st", c_longlong),
                ("k", c_void_p), ("k_sb", c_longlong), ("k_st", c_longlong),
                ("v", c_void_p), ("v_sb", c_longlong), ("v_st", c_longlong),
                ("out", c_void_p), ("o_sb", c_longlong), ("o_st", c_longlong),
                ("dtype", c_int), ("scale", c_float), ("causal", c_int), ("key_mask", c_void_p)]


_PROF = None


class LaunchProfiler:
    """Times every libtcavp launch with CUDA events on the launching stream (bench.py's roofline numbers).
    Events are recorded around each launch inside the timed region; nothing synchronises until summary()."""

    def __init__(self):
        self.rec = []

    def __enter__(self):
        global _PROF
        _PROF = self
        return self

    def __exit__(self, *a):
        global _PROF
        _PROF = None

    def add(self, kernel, flops, nbytes, e0, e1):
        self.rec.append((kernel, flops, nbytes, e0, e1))

    def summary(self):
        torch.cuda.synchronize()
        groups, base = {}, {}
        for kernel, flops, nbytes, e0, e1 in self.rec:
            ms = e0.elapsed_time(e1)
            for key, table in ((kernel, groups), (kernel.split("[")[0], base)):
                g = table.setdefault(key, dict(kernel=key, launches=0, time_ms=0.0, flops=0.0, bytes=0.0))
                g["launches"] += 1
                g["time_ms"] += ms
                g["flops"] += flops
                g["bytes"] += nbytes
        for g in list(groups.values()) + list(base.values()):
            g["avg_ms"] = g["time_ms"] / max(g["launches"], 1)
            g["tflops"] = g["flops"] / (g["time_ms"] * 1e-3) / 1e12 if g["time_ms"] > 0 else 0.0
            g["gbs"] = g["bytes"] / (g["time_ms"] * 1e-3) / 1e9 if g["time_ms"] > 0 else 0.0
        order = sorted(groups.values(), key=lambda g: -g["time_ms"])
        total = sum(g["time_ms"] for g in order) or 1.0
        brief = [dict(kernel=g["kernel"], launches=g["launches"], time_ms=round(g["time_ms"], 3), share=round(g["time_ms"] / total, 4),
                      tflops=round(g["tflops"], 1), gbs=round(g["gbs"], 1)) for g in order]
        dominant = max(base.values(), key=lambda g: g["time_ms"]) if base else None
        return {"dominant": dominant, "groups": brief}


class _Timed:
    __slots__ = ("kernel", "flops", "nbytes", "e0")

    def __init__(self, kernel, flops=0.0, nbytes=0.0):
        self.kernel, self.flops, self.nbytes = kernel, flops, nbytes

    def __enter__(self):
        if _PROF is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if _PROF is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _PROF.add(self.kernel, self.flops, self.nbytes, self.e0, e1)


def _nb(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}") from None


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.TcavpError("tcavp ops need CUDA tensors (there is no CPU fallback)")


def gemm(a, w, out, *, M=None, N=None, K=None, lda=None, ldw=None, ldo=None, bias=None, residual=None, ldr=None,
         act=ACT_NONE, remap=(0, 0, 0), rope=None, row_scale=None):
    """out = act(a @ w.T + bias) + residual.  a: [M, >=K] row-major (lda = a.stride(0)), w: [N, >=K]."""
    _need_cuda(a, w, out, bias, residual)
    g = GemmArgs()
    g.M = a.shape[0] if M is None else M
    g.N = w.shape[0] if N is None else N
    g.K = a.shape[1] if K is None else K
    g.A, g.lda = a.data_ptr(), (a.stride(0) if lda is None else lda)
    g.W, g.ldw = w.data_ptr(), (w.stride(0) if ldw is None else ldw)
    if a.dtype != w.dtype:
        raise TypeError(f"gemm: a is {a.dtype}, w is {w.dtype}")
    g.in_dtype = dt(a)
    g.out, g.ldo, g.out_dtype = out.data_ptr(), (out.stride(0) if ldo is None else ldo), dt(out)
    g.bias = None if bias is None else bias.data_ptr()
    if bias is not None and bias.dtype != torch.float32:
        raise TypeError("gemm: bias must be fp32")
    if residual is not None:
        g.residual, g.ldr, g.res_dtype = residual.data_ptr(), (residual.stride(0) if ldr is None else ldr), dt(residual)
    g.act = act
    g.remap_gi, g.remap_go, g.remap_off = remap
    if row_scale is not None:
        if row_scale.dtype != torch.float32:
            raise TypeError("gemm: row_scale must be fp32")
        g.row_scale = row_scale.data_ptr()
    if rope is not None:     # (table, L, dh, cols)
        g.rope_cos_sin, g.rope_L, g.rope_dh, g.rope_cols = rope[0].data_ptr(), rope[1], rope[2], rope[3]
    if g.in_dtype == BF16:
        kern = "gemm_tc_kernel<%d>[N%d,K%d]" % (32 if g.N <= 32 else 64 if g.N <= 64 else 128 if g.N <= 128 else 256, g.N, g.K)
    else:
        kern = "gemm_simt_kernel"
    esz = a.element_size()
    with _Timed(kern, 2.0 * g.M * g.N * g.K, float(g.M * g.K * esz + g.N * g.K * esz + g.M * g.N * out.element_size())):
        _lib.check(_lib.load().tcavp_gemm(byref(g), _stream()), "tcavp_gemm")
    return out


def attention(q, k, v, out, *, B, H, Hkv, Tq, Tk, dh, q_strides, k_strides, v_strides, o_strides, scale, causal=False,
              key_mask=None):
    """q/k/v/out are tensors whose data_ptr() is element (0,0,0,0); *_strides = (batch stride, time stride) in elements."""
    _need_cuda(q, k, v, out, key_mask)
    a = AttnArgs()
    a.B, a.H, a.Hkv, a.Tq, a.Tk, a.dh = B, H, Hkv, Tq, Tk, dh
    a.q, (a.q_sb, a.q_st) = q.data_ptr(), q_strides
    a.k, (a.k_sb, a.k_st) = k.data_ptr(), k_strides
    a.v, (a.v_sb, a.v_st) = v.data_ptr(), v_strides
    a.out, (a.o_sb, a.o_st) = out.data_ptr(), o_strides
    a.dtype, a.scale, a.causal = dt(q), scale, int(causal)
    if key_mask is not None:
        if key_mask.dtype != torch.int32:
            raise TypeError("attention: key_mask must be int32")
        a.key_mask = key_mask.data_ptr()
    if a.dtype == BF16 and dh in (16, 32, 64, 96, 128) and max(Tq, Tk) <= 256:
        kern = f"attn_flash_kernel[dh{dh},q{Tq},k{Tk}]"
    elif a.dtype == BF16 and dh > 128 and dh % 64 == 0 and Tq <= 32 and Tk <= 256 and not causal:
        kern = "attn_x_kernel"
    elif dh in (16, 32) and Tk <= 128:
        kern = f"attn_row_kernel[dh{dh},q{Tq},k{Tk}]"
    else:
        kern = f"attn_warp_kernel[dh{dh},q{Tq},k{Tk}]"
    fl = 4.0 * B * H * Tq * Tk * dh * (0.5 if causal else 1.0)
    with _Timed(kern, fl, 2.0 * q.element_size() * B * dh * (H * Tq + Hkv * Tk)):
        _lib.check(_lib.load().tcavp_attention(byref(a), _stream()), "tcavp_attention")
    return out


def layernorm(x, w, b, out, *, residual=None, eps=1e-5, remap=(0, 0, 0), rowvec=None, rows=None, cols=None):
    _need_cuda(x, w, b, out, residual, rowvec)
    rows = x.numel() // x.shape[-1] if rows is None else rows
    cols = x.shape[-1] if cols is None else cols
    with _Timed("layernorm_kernel"):
        _lib.check(_lib.load().tcavp_layernorm(_p(x), _p(residual), _p(w), _p(b), _p(out), rows, cols, c_float(eps), dt(x), dt(out),
                                           remap[0], remap[1], remap[2], _p(rowvec), _stream()), "tcavp_layernorm")
    return out


def rmsnorm(x, w, out, *, eps, rows=None, cols=None, ldi=None, ldo=None):
    _need_cuda(x, w, out)
    rows = x.numel() // x.shape[-1] if rows is None else rows
    cols = x.shape[-1] if cols is None else cols
    ldo = cols if ldo is None else ldo
    ldi = cols if ldi is None else ldi
    with _Timed("rmsnorm_kernel", 0.0, float(rows * cols * (x.element_size() + out.element_size()))):
        _lib.check(_lib.load().tcavp_rmsnorm(_p(x), ldi, _p(w), _p(out), rows, cols, ldo, c_float(eps), dt(x), dt(out), _stream()),
                   "tcavp_rmsnorm")
    return out


def row_rstd(x, out, *, rows, cols, ldx, eps):
    _need_cuda(x, out)
    with _Timed("row_rstd_kernel", 0.0, float(rows * cols * x.element_size())):
        _lib.check(_lib.load().tcavp_row_rstd(_p(x), ldx, rows, cols, c_float(eps), dt(x), _p(out), _stream()), "tcavp_row_rstd")
    return out


def rope_table(L, dh, theta, device, layout=0):
    # HF:86-88 inv_freq, evaluated on the host exactly as transformers does
    inv = (1.0 / (theta ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))).to(device)
    t = torch.empty((L, dh // 2, 2) if layout == 0 else (dh // 4, L, 4), dtype=torch.float32, device=device)
    with _Timed("rope_table_kernel"):
        _lib.check(_lib.load().tcavp_rope_table(_p(t), _p(inv), L, dh, layout, _stream()), "tcavp_rope_table")
    return t


def rope_(qkv, *, rows, L, ld, n_q_heads, n_k_heads, dh, table):
    _need_cuda(qkv, table)
    with _Timed("rope_kernel"):
        _lib.check(_lib.load().tcavp_rope(_p(qkv), rows, L, ld, n_q_heads, n_k_heads, dh, _p(table), dt(qkv), _stream()), "tcavp_rope")
    return qkv


def embed_text(ids, attn_mask, embed, text_mod, fused, mask_out, *, B, L_text, n_img, H):
    _need_cuda(ids, attn_mask, embed, text_mod, fused, mask_out)
    if ids.dtype != torch.int64 or (attn_mask is not None and attn_mask.dtype != torch.int64):
        raise TypeError("embed_text: ids / attention_mask must be int64")
    with _Timed("embed_text_kernel"):
        _lib.check(_lib.load().tcavp_embed_text(_p(ids), _p(attn_mask), _p(embed), dt(embed), _p(text_mod), _p(fused), dt(fused),
                                            _p(mask_out), B, L_text, n_img, H, embed.shape[0], _stream()), "tcavp_embed_text")


def add_rowvec(x, rowvec, out, *, rows, cols, remap=(0, 0, 0)):
    _need_cuda(x, rowvec, out)
    with _Timed("add_rowvec_kernel"):
        _lib.check(_lib.load().tcavp_add_rowvec(_p(x), _p(rowvec), _p(out), rows, cols, dt(x), dt(out), remap[0], remap[1], remap[2],
                                            _stream()), "tcavp_add_rowvec")
    return out


def cast(x, out, *, rows, cols, ldi=None, ldo=None, in_row_mod=0):
    _need_cuda(x, out)
    with _Timed("cast_kernel"):
        _lib.check(_lib.load().tcavp_cast(_p(x), cols if ldi is None else ldi, dt(x), _p(out), cols if ldo is None else ldo, dt(out),
                                      rows, cols, in_row_mod, _stream()), "tcavp_cast")
    return out


def poly_embed(polygon, lens, w, bias, pos, out, key_mask, *, B, P, D):
    _need_cuda(polygon, lens, w, bias, pos, out, key_mask)
    with _Timed("poly_embed_kernel"):
        _lib.check(_lib.load().tcavp_poly_embed(_p(polygon), _p(lens), _p(w), _p(bias), _p(pos), _p(out), dt(out), _p(key_mask), B, P, D,
                                            _stream()), "tcavp_poly_embed")


def masked_mean(x, lens, out, *, B, P, D):
    _need_cuda(x, lens, out)
    with _Timed("masked_mean_kernel"):
        _lib.check(_lib.load().tcavp_masked_mean(_p(x), dt(x), _p(lens), _p(out), dt(out), B, P, D, _stream()), "tcavp_masked_mean")
    return out


def ltsf_encode(x, wt, bt, we, be, pos, enc, *, B, F, C, T_in):
    _need_cuda(x, wt, bt, we, be, pos, enc)
    with _Timed("ltsf_encode_kernel"):
        _lib.check(_lib.load().tcavp_ltsf_encode(_p(x), _p(wt), _p(bt), _p(we), _p(be), _p(pos), _p(enc), dt(enc), B, F, C, T_in,
                                             _stream()), "tcavp_ltsf_encode")
    return enc


def nlinear_decode(enc, wd, bd, lane_adj, dec, *, B, C, T_in, T_out):
    _need_cuda(enc, wd, bd, lane_adj, dec)
    with _Timed("nlinear_decode_kernel"):
        _lib.check(_lib.load().tcavp_nlinear_decode(_p(enc), dt(enc), _p(wd), _p(bd), _p(lane_adj), 0 if lane_adj is None else dt(lane_adj),
                                                _p(dec), dt(dec), B, C, T_in, T_out, _stream()), "tcavp_nlinear_decode")
    return dec


def fusion_head(fused, ln_w, ln_b, w1, b1, w2, b2, wo, bo, x, decoded, *, y=None, norm_stat=None, metrics=None, per_scene=None,
                B, C, T_in, T_out):
    _need_cuda(fused, x, decoded, y, norm_stat, metrics, per_scene)
    with _Timed("fusion_head_kernel"):
        _lib.check(_lib.load().tcavp_fusion_head(_p(fused), dt(fused), _p(ln_w), _p(ln_b), _p(w1), _p(b1), _p(w2), _p(b2), _p(wo), _p(bo),
                                             _p(x), _p(decoded), _p(y), _p(norm_stat), _p(metrics), _p(per_scene), B, C, T_in, T_out,
                                             _stream()), "tcavp_fusion_head")
    return decoded


def traj_metrics(decoded, y, norm_stat, metrics, per_scene, *, B, T_out):
    _need_cuda(decoded, y, norm_stat, metrics, per_scene)
    with _Timed("traj_metrics_kernel"):
        _lib.check(_lib.load().tcavp_traj_metrics(_p(decoded), _p(y), _p(norm_stat), _p(metrics), _p(per_scene), B, T_out, _stream()),
               "tcavp_traj_metrics")


def launch_count():
    return int(_lib.load().tcavp_launch_count())


def device_info():
    sm, maj, mnr = c_int(), c_int(), c_int()
    _lib.check(_lib.load().tcavp_device_info(byref(sm), byref(maj), byref(mnr)), "tcavp_device_info")
    return sm.value, maj.value, mnr.value
