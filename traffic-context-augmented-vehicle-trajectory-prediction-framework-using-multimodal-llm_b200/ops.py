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
                ("remap_gi", c_int), ("remap_go", c_int), ("remap_off", c_int)]


class AttnArgs(Structure):
    _fields_ = [("B", c_int), ("H", c_int), ("Hkv", c_int), ("Tq", c_int), ("Tk", c_int), ("dh", c_int),
                ("q", c_void_p), ("q_sb", c_longlong), ("q_st", c_longlong),
                ("k", c_void_p), ("k_sb", c_longlong), ("k_st", c_longlong),
                ("v", c_void_p), ("v_sb", c_longlong), ("v_st", c_longlong),
                ("out", c_void_p), ("o_sb", c_longlong), ("o_st", c_longlong),
                ("dtype", c_int), ("scale", c_float), ("causal", c_int), ("key_mask", c_void_p)]


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
         act=ACT_NONE, remap=(0, 0, 0)):
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
    _lib.check(_lib.load().tcavp_attention(byref(a), _stream()), "tcavp_attention")
    return out


def layernorm(x, w, b, out, *, residual=None, eps=1e-5, remap=(0, 0, 0), rowvec=None, rows=None, cols=None):
    _need_cuda(x, w, b, out, residual, rowvec)
    rows = x.numel() // x.shape[-1] if rows is None else rows
    cols = x.shape[-1] if cols is None else cols
    _lib.check(_lib.load().tcavp_layernorm(_p(x), _p(residual), _p(w), _p(b), _p(out), rows, cols, c_float(eps), dt(x), dt(out),
                                           remap[0], remap[1], remap[2], _p(rowvec), _stream()), "tcavp_layernorm")
    return out


def rmsnorm(x, w, out, *, eps, rows=None, cols=None, ldo=None):
    _need_cuda(x, w, out)
    rows = x.numel() // x.shape[-1] if rows is None else rows
    cols = x.shape[-1] if cols is None else cols
    ldo = cols if ldo is None else ldo
    _lib.check(_lib.load().tcavp_rmsnorm(_p(x), _p(w), _p(out), rows, cols, ldo, c_float(eps), dt(x), dt(out), _stream()),
               "tcavp_rmsnorm")
    return out


def rope_table(L, dh, theta, device):
    t = torch.empty(L, dh // 2, 2, dtype=torch.float32, device=device)
    _lib.check(_lib.load().tcavp_rope_table(_p(t), L, dh, c_float(theta), _stream()), "tcavp_rope_table")
    return t


def rope_(qkv, *, rows, L, ld, n_q_heads, n_k_heads, dh, table):
    _need_cuda(qkv, table)
    _lib.check(_lib.load().tcavp_rope(_p(qkv), rows, L, ld, n_q_heads, n_k_heads, dh, _p(table), dt(qkv), _stream()), "tcavp_rope")
    return qkv


def embed_text(ids, attn_mask, embed, text_mod, fused, mask_out, *, B, L_text, n_img, H):
    _need_cuda(ids, attn_mask, embed, text_mod, fused, mask_out)
    if ids.dtype != torch.int64 or (attn_mask is not None and attn_mask.dtype != torch.int64):
        raise TypeError("embed_text: ids / attention_mask must be int64")
    _lib.check(_lib.load().tcavp_embed_text(_p(ids), _p(attn_mask), _p(embed), dt(embed), _p(text_mod), _p(fused), dt(fused),
                                            _p(mask_out), B, L_text, n_img, H, embed.shape[0], _stream()), "tcavp_embed_text")


def add_rowvec(x, rowvec, out, *, rows, cols, remap=(0, 0, 0)):
    _need_cuda(x, rowvec, out)
    _lib.check(_lib.load().tcavp_add_rowvec(_p(x), _p(rowvec), _p(out), rows, cols, dt(x), dt(out), remap[0], remap[1], remap[2],
                                            _stream()), "tcavp_add_rowvec")
    return out


def cast(x, out, *, rows, cols, ldi=None, ldo=None, in_row_mod=0):
    _need_cuda(x, out)
    _lib.check(_lib.load().tcavp_cast(_p(x), cols if ldi is None else ldi, dt(x), _p(out), cols if ldo is None else ldo, dt(out),
                                      rows, cols, in_row_mod, _stream()), "tcavp_cast")
    return out


def poly_embed(polygon, lens, w, bias, pos, out, key_mask, *, B, P, D):
    _need_cuda(polygon, lens, w, bias, pos, out, key_mask)
    _lib.check(_lib.load().tcavp_poly_embed(_p(polygon), _p(lens), _p(w), _p(bias), _p(pos), _p(out), dt(out), _p(key_mask), B, P, D,
                                            _stream()), "tcavp_poly_embed")


def masked_mean(x, lens, out, *, B, P, D):
    _need_cuda(x, lens, out)
    _lib.check(_lib.load().tcavp_masked_mean(_p(x), dt(x), _p(lens), _p(out), dt(out), B, P, D, _stream()), "tcavp_masked_mean")
    return out


def ltsf_encode(x, wt, bt, we, be, pos, enc, *, B, F, C, T_in):
    _need_cuda(x, wt, bt, we, be, pos, enc)
    _lib.check(_lib.load().tcavp_ltsf_encode(_p(x), _p(wt), _p(bt), _p(we), _p(be), _p(pos), _p(enc), dt(enc), B, F, C, T_in,
                                             _stream()), "tcavp_ltsf_encode")
    return enc


def nlinear_decode(enc, wd, bd, lane_adj, dec, *, B, C, T_in, T_out):
    _need_cuda(enc, wd, bd, lane_adj, dec)
    _lib.check(_lib.load().tcavp_nlinear_decode(_p(enc), dt(enc), _p(wd), _p(bd), _p(lane_adj), 0 if lane_adj is None else dt(lane_adj),
                                                _p(dec), dt(dec), B, C, T_in, T_out, _stream()), "tcavp_nlinear_decode")
    return dec


def fusion_head(fused, ln_w, ln_b, w1, b1, w2, b2, wo, bo, x, decoded, *, y=None, norm_stat=None, metrics=None, per_scene=None,
                B, C, T_in, T_out):
    _need_cuda(fused, x, decoded, y, norm_stat, metrics, per_scene)
    _lib.check(_lib.load().tcavp_fusion_head(_p(fused), dt(fused), _p(ln_w), _p(ln_b), _p(w1), _p(b1), _p(w2), _p(b2), _p(wo), _p(bo),
                                             _p(x), _p(decoded), _p(y), _p(norm_stat), _p(metrics), _p(per_scene), B, C, T_in, T_out,
                                             _stream()), "tcavp_fusion_head")
    return decoded


def traj_metrics(decoded, y, norm_stat, metrics, per_scene, *, B, T_out):
    _need_cuda(decoded, y, norm_stat, metrics, per_scene)
    _lib.check(_lib.load().tcavp_traj_metrics(_p(decoded), _p(y), _p(norm_stat), _p(metrics), _p(per_scene), B, T_out, _stream()),
               "tcavp_traj_metrics")


def launch_count():
    return int(_lib.load().tcavp_launch_count())


def device_info():
    sm, maj, mnr = c_int(), c_int(), c_int()
    _lib.check(_lib.load().tcavp_device_info(byref(sm), byref(maj), byref(mnr)), "tcavp_device_info")
    return sm.value, maj.value, mnr.value
