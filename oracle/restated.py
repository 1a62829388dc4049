"""TEST INFRASTRUCTURE — not product code (see oracle/README.md).

CPU restatement (plain tensor arithmetic, fp32, no nn.Module, no HF, no peft) of the reference forward
hot path, driven by a reference-layout state_dict.  Every function cites the reference lines it follows
("HF:" = transformers 5.5.0 models/llama/modeling_llama.py, "torch:" = torch 2.11 nn/functional.py /
nn/modules/transformer.py — third-party code the reference calls but does not vendor or pin).

Pinning: tests/test_oracle_cpu.py checks this file against tests/golden/*.pt, which
oracle/make_golden.py minted by running the UNMODIFIED reference classes (oracle/ref_loader.py) in the
authoring container.  The reference itself ships no tests or golden vectors, so parity is
pinned to the reference's own code executed here, not to reference-owned fixtures.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def layer_norm(x, w, b, eps=1e-5):
    """torch nn.LayerNorm (biased variance, eps inside the sqrt)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def rms_norm(x, w, eps):
    """HF:53-70 LlamaRMSNorm: fp32 variance, weight multiplies the re-cast activation."""
    xf = x.float()
    var = (xf * xf).mean(-1, keepdim=True)
    return w * (xf * torch.rsqrt(var + eps)).to(x.dtype)


# Train-mode dropout: oracle/dropout.py's DropOracle, set by `with dropout(...)`; None = eval mode (every p = 0, the default).
DROP = None


def _D(t, mod, layer, kind):
    """One dropout site (identity in eval mode); `t` is in the canonical batch-major layout."""
    return t if DROP is None else DROP.apply(t, mod, layer, kind)


class dropout:
    """with restated.dropout(DropOracle(...)): forward(...)   — runs the restatement in train mode with those masks."""

    def __init__(self, oracle):
        self.oracle = oracle

    def __enter__(self):
        global DROP
        self.prev, DROP = DROP, self.oracle
        return self.oracle

    def __exit__(self, *a):
        global DROP
        DROP = self.prev


def mha(q_in, k_in, v_in, w_in, b_in, w_out, b_out, nheads, key_padding_mask=None, site=None):
    """torch: F.multi_head_attention_forward (packed in_proj, batch-first view here).
    q_in (B,Tq,E), k_in/v_in (B,Tk,E); key_padding_mask (B,Tk) bool, True = ignore.  site = (module, layer, kind) of the
    dropout on the attention probabilities (train mode)."""
    B, Tq, E = q_in.shape
    Tk = k_in.shape[1]
    dh = E // nheads
    wq, wk, wv = w_in[:E], w_in[E:2 * E], w_in[2 * E:]
    bq, bk, bv = b_in[:E], b_in[E:2 * E], b_in[2 * E:]
    q = linear(q_in, wq, bq).view(B, Tq, nheads, dh).transpose(1, 2)
    k = linear(k_in, wk, bk).view(B, Tk, nheads, dh).transpose(1, 2)
    v = linear(v_in, wv, bv).view(B, Tk, nheads, dh).transpose(1, 2)
    s = (q * (1.0 / math.sqrt(dh))) @ k.transpose(-1, -2)
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    if site is not None:
        p = _D(p, *site)
    o = (p @ v).transpose(1, 2).reshape(B, Tq, E)
    return linear(o, w_out, b_out)


def _g(sd, prefix, name):
    return sd[prefix + name]


def encoder_layer(sd, p, x, nheads, key_padding_mask=None, mod=None, li=0):
    """torch: nn.TransformerEncoderLayer, norm_first=False, activation=relu:
    x = norm1(x + dropout1(sa(x))); x = norm2(x + dropout2(linear2(dropout(relu(linear1(x))))))   (dropout sites: train mode only)."""
    a = mha(x, x, x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
            sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nheads, key_padding_mask, site=(mod, li, "sa_attn"))
    x = layer_norm(x + _D(a, mod, li, "drop1"), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    h = _D(torch.relu(linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])), mod, li, "ffn")
    f = linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return layer_norm(x + _D(f, mod, li, "drop2"), sd[p + "norm2.weight"], sd[p + "norm2.bias"])


def decoder_layer(sd, p, x, mem, nheads, li=0):
    """torch: nn.TransformerDecoderLayer, norm_first=False, relu, no masks (reference train.py:413)."""
    mod = "qdec"
    a = mha(x, x, x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
            sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nheads, site=(mod, li, "sa_attn"))
    x = layer_norm(x + _D(a, mod, li, "drop1"), sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    c = mha(x, mem, mem, sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"],
            sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"], nheads, site=(mod, li, "ca_attn"))
    x = layer_norm(x + _D(c, mod, li, "drop2"), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    h = _D(torch.relu(linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])), mod, li, "ffn")
    f = linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return layer_norm(x + _D(f, mod, li, "drop3"), sd[p + "norm3.weight"], sd[p + "norm3.bias"])


def _count_layers(sd, prefix):
    n = 0
    while any(k.startswith(f"{prefix}{n}.") for k in sd):
        n += 1
    return n


# --------------------------------------------------------------------------------------------------
# reference scripts/train.py:352-383  LanePolygonEncoder
# --------------------------------------------------------------------------------------------------


def lane_polygon_encoder(sd, polygon, poly_len, nheads=4, p="lane_polygon_encoder."):
    B, P, _ = polygon.shape
    x = linear(polygon, sd[p + "input_proj.weight"], sd[p + "input_proj.bias"]) + sd[p + "pos_embedding"][:, :P]
    lens = torch.as_tensor(list(poly_len), dtype=torch.long)
    pad = torch.arange(P)[None, :] >= lens[:, None]                       # train.py:366-370
    for i in range(_count_layers(sd, p + "encoder.layers.")):
        x = encoder_layer(sd, f"{p}encoder.layers.{i}.", x, nheads, pad, mod="poly", li=i)
    valid = (~pad).to(x.dtype)[:, :, None]                                # train.py:373-382 masked mean, 0 if len==0
    x = torch.nan_to_num(x, nan=0.0)  # rows with len==0 are all-masked (NaN softmax) and discarded by the reference
    s = (x * valid).sum(1)
    return torch.where(lens[:, None] > 0, s / lens.clamp(min=1)[:, None].to(x.dtype), torch.zeros_like(s))


# --------------------------------------------------------------------------------------------------
# reference scripts/train.py:388-414  BlipQFormer
# --------------------------------------------------------------------------------------------------


def qformer(sd, vision, nheads=8, p="mllm.qformer."):
    B = vision.shape[0]
    x = linear(vision, sd[p + "vision_proj.weight"], sd[p + "vision_proj.bias"])
    for i in range(_count_layers(sd, p + "encoder.layers.")):
        x = encoder_layer(sd, f"{p}encoder.layers.{i}.", x, nheads, mod="qenc", li=i)
    q = sd[p + "query_tokens"].unsqueeze(0).expand(B, -1, -1)
    for i in range(_count_layers(sd, p + "decoder.layers.")):
        q = decoder_layer(sd, f"{p}decoder.layers.{i}.", q, x, nheads, li=i)
    return q


# --------------------------------------------------------------------------------------------------
# HF LlamaModel (HF:375-427) with peft lora.Linear on q_proj / v_proj (oracle/peft_shim.py)
# --------------------------------------------------------------------------------------------------


def find_llm_prefix(sd):
    """Prefix up to and including 'model.' of LlamaModel (or 'transformer.' of GPT2Model)."""
    for k in sd:
        if k.endswith("model.embed_tokens.weight"):
            return k[: -len("embed_tokens.weight")]
        if k.endswith("transformer.wte.weight"):
            return k[: -len("wte.weight")]
    raise KeyError("no embed_tokens / wte in state_dict")


def _lora_linear(sd, p, x, scaling, site=None):
    """peft lora.Linear.forward: base(x) + B(A(dropout(x))) * alpha/r (dropout in train mode only).  Accepts peft>=0.7 (.base_layer)
    and the older layout without it (reference ablation_study_without_lora.py:1071-1079)."""
    w = sd.get(p + "base_layer.weight", sd.get(p + "weight"))
    y = linear(x, w)
    if p + "lora_A.default.weight" in sd:
        xd = x if site is None else _D(x, *site)
        y = y + linear(linear(xd, sd[p + "lora_A.default.weight"]), sd[p + "lora_B.default.weight"]) * scaling
    return y


def llama3_scale_inv_freq(inv, rs):
    """transformers modeling_rope_utils._compute_llama3_parameters (rope_type "llama3", Llama-3.1 / 3.2), restated per frequency:
    long wavelengths (> ctx / low_freq_factor) are slowed by `factor`, short ones (< ctx / high_freq_factor) are untouched, the band
    in between blends the two.  Pinned by the llama32_1b_l2_b8 golden (minted through the real HF LlamaRotaryEmbedding)."""
    ctx, f, lo, hi = rs["original_max_position_embeddings"], rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"]
    out = inv.clone()
    for j in range(inv.numel()):
        w = inv[j:j + 1]                                     # keep fp32 tensor arithmetic (same rounding as HF's vectorised form)
        wavelen = 2 * math.pi / w
        if bool(wavelen > ctx / lo):
            out[j] = (w / f)[0]
        elif not bool(wavelen < ctx / hi):
            s = (ctx / wavelen - lo) / (hi - lo)
            out[j] = ((1 - s) * w / f + s * w)[0]
    return out


def rope_cos_sin(L, dh, theta, rope_scaling=None):
    """HF:73-136 LlamaRotaryEmbedding, default / llama3 rope type, positions 0..L-1 (inputs_embeds path: HF:392-397)."""
    inv = 1.0 / (theta ** (torch.arange(0, dh, 2, dtype=torch.int64).float() / dh))
    if rope_scaling and rope_scaling.get("rope_type", "default") == "llama3":
        inv = llama3_scale_inv_freq(inv, rope_scaling)
    fr = torch.arange(L, dtype=torch.float32)[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def llama_stack(sd, llama_cfg, embeds, attn_mask, lora_scaling, p=None):
    """embeds (B,L,H); attn_mask (B,L) 1 = valid.  Returns hidden_states[-1] (post final norm, HF:421)."""
    p = p or find_llm_prefix(sd)
    B, L, H = embeds.shape
    nh = llama_cfg["num_attention_heads"]
    nkv = llama_cfg.get("num_key_value_heads", nh)
    dh = llama_cfg.get("head_dim", H // nh)
    eps = llama_cfg.get("rms_norm_eps", 1e-6)
    cos, sin = rope_cos_sin(L, dh, float(llama_cfg.get("rope_theta", 10000.0)), llama_cfg.get("rope_scaling"))
    causal = torch.tril(torch.ones(L, L, dtype=torch.bool))
    allowed = causal[None, :, :] & attn_mask.bool()[:, None, :]           # HF:399 create_causal_mask
    x = embeds
    for i in range(llama_cfg["num_hidden_layers"]):
        lp = f"{p}layers.{i}."
        h = rms_norm(x, sd[lp + "input_layernorm.weight"], eps)
        q = _lora_linear(sd, lp + "self_attn.q_proj.", h, lora_scaling, ("llm", i, "lora_q")).view(B, L, nh, dh).transpose(1, 2)
        k = _lora_linear(sd, lp + "self_attn.k_proj.", h, lora_scaling, ("llm", i, "lora_k")).view(B, L, nkv, dh).transpose(1, 2)
        v = _lora_linear(sd, lp + "self_attn.v_proj.", h, lora_scaling, ("llm", i, "lora_v")).view(B, L, nkv, dh).transpose(1, 2)
        q = q * cos + rotate_half(q) * sin                                  # HF:146-170
        k = k * cos + rotate_half(k) * sin
        if nkv != nh:
            k = k.repeat_interleave(nh // nkv, dim=1)                       # HF:173-182 repeat_kv
            v = v.repeat_interleave(nh // nkv, dim=1)
        s = (q @ k.transpose(-1, -2)) * (dh ** -0.5)
        s = s.masked_fill(~allowed[:, None], float("-inf"))
        a = torch.softmax(s.float(), dim=-1).to(q.dtype) @ v
        a = a.transpose(1, 2).reshape(B, L, nh * dh)
        x = x + _lora_linear(sd, lp + "self_attn.o_proj.", a, lora_scaling)
        h = rms_norm(x, sd[lp + "post_attention_layernorm.weight"], eps)
        g = _lora_linear(sd, lp + "mlp.gate_proj.", h, lora_scaling)
        u = _lora_linear(sd, lp + "mlp.up_proj.", h, lora_scaling)
        x = x + _lora_linear(sd, lp + "mlp.down_proj.", F.silu(g) * u, lora_scaling)   # HF:182-196
    return rms_norm(x, sd[p + "norm.weight"], eps)


def _conv1d(sd, p, x, scaling, site=None):
    """transformers Conv1D (y = x W + b, W [in, out]) with an optional peft LoRA pair around it (fan_in_fan_out: the adapters are plain
    Linears): base_layer.{weight, bias} + lora_B(lora_A(dropout(x))) alpha / r (dropout in train mode only)."""
    if p + "base_layer.weight" in sd:
        y = x @ sd[p + "base_layer.weight"] + sd[p + "base_layer.bias"]
        xd = x if site is None else _D(x, *site)
        return y + (xd @ sd[p + "lora_A.default.weight"].t()) @ sd[p + "lora_B.default.weight"].t() * scaling
    return x @ sd[p + "weight"] + sd[p + "bias"]


def gpt2_stack(sd, cfg, embeds, attn_mask, lora_scaling, p=None):
    """HF GPT2Model over inputs_embeds (modeling_gpt2.py: GPT2Model.forward / GPT2Block / GPT2Attention / GPT2MLP): + wpe[0..L), pre-norm
    blocks with LayerNorm (bias), fused c_attn -> (q, k, v), causal + key-padding mask, scores / sqrt(dh), gelu_new MLP, ln_f.
    Returns hidden_states[-1] (after ln_f)."""
    p = p or find_llm_prefix(sd)
    B, L, H = embeds.shape
    nh = cfg["num_attention_heads"]
    dh = H // nh
    eps = cfg.get("layer_norm_epsilon", 1e-5)
    allowed = torch.tril(torch.ones(L, L, dtype=torch.bool))[None] & attn_mask.bool()[:, None, :]
    x = _D(embeds + sd[p + "wpe.weight"][:L], "llm", 0, "embd")                               # GPT2Model.drop (train mode only)
    for i in range(cfg["num_hidden_layers"]):
        bp = f"{p}h.{i}."
        h = layer_norm(x, sd[bp + "ln_1.weight"], sd[bp + "ln_1.bias"], eps)
        q, k, v = _conv1d(sd, bp + "attn.c_attn.", h, lora_scaling, ("llm", i, "lora_c")).split(H, dim=-1)
        q, k, v = (t.view(B, L, nh, dh).transpose(1, 2) for t in (q, k, v))
        s = (q @ k.transpose(-1, -2)) * (dh ** -0.5)
        s = s.masked_fill(~allowed[:, None], float("-inf"))
        pr = _D(torch.softmax(s.float(), dim=-1).to(q.dtype), "llm", i, "attn")                # attn_dropout on the probabilities
        a = (pr @ v).transpose(1, 2).reshape(B, L, H)
        x = x + _D(_conv1d(sd, bp + "attn.c_proj.", a, lora_scaling), "llm", i, "resid1")      # resid_dropout
        h = layer_norm(x, sd[bp + "ln_2.weight"], sd[bp + "ln_2.bias"], eps)
        h = F.gelu(_conv1d(sd, bp + "mlp.c_fc.", h, lora_scaling), approximate="tanh")      # ACT2FN["gelu_new"]
        x = x + _D(_conv1d(sd, bp + "mlp.c_proj.", h, lora_scaling), "llm", i, "resid2")       # GPT2MLP.dropout
    return layer_norm(x, sd[p + "ln_f.weight"], sd[p + "ln_f.bias"], eps)


# --------------------------------------------------------------------------------------------------
# reference scripts/train.py:504-554  LlamaMultiModal.forward (ids branch)
# --------------------------------------------------------------------------------------------------


def mllm_forward(sd, cfg, llama_cfg, vision, input_ids, attention_mask, p="mllm."):
    img = qformer(sd, vision, cfg.get("q_nhead", 8), p + "qformer.")
    if p + "q_proj.weight" in sd:
        img = linear(img, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    img = img + sd[p + "vision_modality_embedding"]
    lp = find_llm_prefix(sd)
    gpt2 = llama_cfg.get("arch") == "gpt2"
    txt = sd[lp + ("wte.weight" if gpt2 else "embed_tokens.weight")][input_ids] + sd[p + "text_modality_embedding"]
    fused = torch.cat([img, txt], dim=1)
    mask = torch.cat([torch.ones(img.shape[:2], dtype=attention_mask.dtype), attention_mask], dim=1)
    scaling = cfg.get("lora_alpha", 32) / cfg.get("lora_r", 8)
    return (gpt2_stack if gpt2 else llama_stack)(sd, llama_cfg, fused, mask, scaling, lp), img


# --------------------------------------------------------------------------------------------------
# reference scripts/train.py:659-842  TransformerLTSF
# --------------------------------------------------------------------------------------------------


def stack_individual(sd, prefix, C):
    w = torch.stack([sd[f"{prefix}{i}.weight"] for i in range(C)])         # (C, T_out, T_in)
    b = torch.stack([sd[f"{prefix}{i}.bias"] for i in range(C)])           # (C, T_out)
    return w, b


def ltsf_encoder(sd, x, p="ltsf."):
    """train.py:837-839 token_proj (1x1 conv) -> NLinear encoder (701-716) -> + pos_encoding."""
    w = sd[p + "token_proj.weight"][:, :, 0]
    xp = torch.einsum("cf,bft->bct", w, x) + sd[p + "token_proj.bias"][None, :, None]
    C = xp.shape[1]
    last = xp[:, :, -1:]
    W, Bv = stack_individual(sd, p + "nlinear_encoder.encoder_linears.", C)
    enc = torch.einsum("cot,bct->bco", W, xp - last) + Bv[None] + last
    return enc + sd[p + "pos_encoding"][:, :, : enc.shape[2]]


def self_attention_block(sd, enc, nheads, p="ltsf.attn_block."):
    """train.py:674-686 — note the residuals are taken from the NORMALISED tensors."""
    x = enc.permute(0, 2, 1)                                                # (B,T,E) batch-first view of (T,B,E)
    xn = layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    a = mha(xn, xn, xn, sd[p + "mha.in_proj_weight"], sd[p + "mha.in_proj_bias"],
            sd[p + "mha.out_proj.weight"], sd[p + "mha.out_proj.bias"], nheads, site=("ltsf", 0, "sa_attn"))
    r = layer_norm(xn + _D(a, "ltsf", 0, "drop1"), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    h = _D(torch.relu(linear(r, sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"])), "ltsf", 0, "ffn")
    f = linear(h, sd[p + "ffn.3.weight"], sd[p + "ffn.3.bias"])
    return (r + _D(f, "ltsf", 0, "drop2")).permute(0, 2, 1)


def ltsf_decoder(sd, enc, poly_emb, final_hidden, out_len, p="ltsf.decoder."):
    """train.py:767-806."""
    B, C, T = enc.shape
    last = enc[:, :, -1:]
    W, Bv = stack_individual(sd, p + "decoder_linears.", C)
    dec = torch.einsum("cot,bct->bco", W, enc - last) + Bv[None] + last
    dec = dec + linear(poly_emb, sd[p + "lane_fc.weight"], sd[p + "lane_fc.bias"]).view(B, C, out_len)
    if p + "post_mlp.0.weight" in sd:                                       # replaces, no residual (787-791)
        h = _D(torch.relu(linear(dec.reshape(B, -1), sd[p + "post_mlp.0.weight"], sd[p + "post_mlp.0.bias"])), "dec", 0, "post")
        dec = linear(h, sd[p + "post_mlp.3.weight"], sd[p + "post_mlp.3.bias"]).view(B, C, out_len)
    dec_t = dec.permute(0, 2, 1)                                            # (B,T_out,C)
    q = linear(dec_t, sd[p + "dec_proj.weight"], sd[p + "dec_proj.bias"])
    cross = mha(q, final_hidden, final_hidden, sd[p + "cross_attn.in_proj_weight"], sd[p + "cross_attn.in_proj_bias"],
                sd[p + "cross_attn.out_proj.weight"], sd[p + "cross_attn.out_proj.bias"], 2, site=("dec", 0, "cross_attn"))   # no key mask (798)
    fused = dec_t + linear(cross, sd[p + "dec_unproj.weight"], sd[p + "dec_unproj.bias"])
    f = layer_norm(fused, sd[p + "fusion_layer.0.weight"], sd[p + "fusion_layer.0.bias"])
    f = linear(torch.relu(linear(f, sd[p + "fusion_layer.1.weight"], sd[p + "fusion_layer.1.bias"])),
               sd[p + "fusion_layer.3.weight"], sd[p + "fusion_layer.3.bias"])
    out = linear(f, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])     # (B,T_out,2)
    return out.permute(0, 2, 1)


# --------------------------------------------------------------------------------------------------
# reference scripts/train.py:914-964 forward; 1302-1322 ADE/FDE; ablation_study_without_lora.py:1237 RMSE
# --------------------------------------------------------------------------------------------------


def denorm(t, norm_stat):
    ns = torch.as_tensor(norm_stat, dtype=t.dtype).view(-1, 4)
    out = t.clone()
    out[:, 0, :] = t[:, 0, :] * (ns[:, 1] - ns[:, 0])[:, None] + ns[:, 0][:, None]
    out[:, 1, :] = t[:, 1, :] * (ns[:, 3] - ns[:, 2])[:, None] + ns[:, 2][:, None]
    return out


def mse_loss(decoded, y, norm_stat):
    d, g = denorm(decoded, norm_stat), denorm(y, norm_stat)
    return ((d[:, 0] - g[:, 0]) ** 2).mean() + ((d[:, 1] - g[:, 1]) ** 2).mean()


def ade_fde(decoded, y, norm_stat):
    """Returns per-scene (ade, fde) on de-normalised coordinates."""
    e = torch.sqrt(((denorm(decoded, norm_stat) - denorm(y, norm_stat)) ** 2).sum(dim=1))
    return e.mean(dim=1), e[:, -1]


@torch.no_grad()
def forward(sd, cfg, llama_cfg, x, vision, polygon, poly_len, input_ids, attention_mask, y=None, norm_stat=None,
            final_hidden=None, with_lm_head=False):
    """cfg: reference ctor kwargs (seq_len, out_len, lane_polygon_nhead, q_nhead, ltsf_nhead, lora_r, lora_alpha).
    Returns a dict with every intermediate the golden fixtures record.  `with_lm_head`: also evaluate the vocabulary logits the
    shipped reference always computes and throws away (HF:487-491 via train.py:445-453, 547-554) — timing only ("as shipped")."""
    sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    out = {}
    out["poly_emb"] = lane_polygon_encoder(sd, polygon, poly_len, cfg.get("lane_polygon_nhead", 4))
    if final_hidden is None:
        out["final_hidden"], out["image_tokens"] = mllm_forward(sd, cfg, llama_cfg, vision, input_ids, attention_mask)
        if with_lm_head:
            lp = find_llm_prefix(sd)
            head = lp[: -len("model.")] + "lm_head.weight"
            out["logits_absmax"] = linear(out["final_hidden"], sd[head]).abs().max()
    else:
        out["final_hidden"] = final_hidden
    enc = ltsf_encoder(sd, x)
    out["enc"] = self_attention_block(sd, enc, cfg.get("ltsf_nhead", 1))
    dec = ltsf_decoder(sd, out["enc"], out["poly_emb"], out["final_hidden"], cfg["out_len"])
    out["decoded"] = dec + x[:, :, -1:]
    if y is not None and norm_stat is not None:
        out["loss"] = mse_loss(out["decoded"], y, norm_stat)
        out["ade"], out["fde"] = ade_fde(out["decoded"], y, norm_stat)
    return out


# --------------------------------------------------------------------------------------------------
# fine-tune step: reference gradients = torch autograd through the restatement above
# (reference scripts/im_kim_train_GRN.py:1029-1039 with every dropout p = 0)
# --------------------------------------------------------------------------------------------------


def trainable_keys(sd):
    """peft semantics (oracle/peft_shim.py): every LLM tensor is frozen except lora_A / lora_B; everything else trains."""
    return [k for k, v in sd.items() if v.is_floating_point() and ("llama_model" not in k or "lora_" in k)]


def loss_and_grads(sd, cfg, llama_cfg, x, vision, polygon, poly_len, input_ids, attention_mask, y, norm_stat, keys=None):
    """Returns (loss, decoded, {key: d loss / d sd[key]}) for the trainable keys (fp32, CPU)."""
    sd = {k: (v.detach().float().clone() if v.is_floating_point() else v) for k, v in sd.items()}
    keys = list(keys) if keys is not None else trainable_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
    with torch.enable_grad():
        out = forward.__wrapped__(sd, cfg, llama_cfg, x, vision, polygon, poly_len, input_ids, attention_mask, y, norm_stat)
        grads = torch.autograd.grad(out["loss"], [sd[k] for k in keys], allow_unused=True)
    return out["loss"].detach(), out["decoded"].detach(), {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(keys, grads)}


# --------------------------------------------------------------------------------------------------
# stage 1: CausalLM objective (reference scripts/check_generation.py:131-151; the same call inside scripts/train.py:533-547)
# --------------------------------------------------------------------------------------------------


def stage1_labels(input_ids, attention_mask, prompt_len):
    """Labels of a synthetic stage-1 batch: the answer part of every sequence (the reference's dataset masks the prompt and the padding
    with -100)."""
    lab = input_ids.clone()
    lab[:, :prompt_len] = -100
    lab[attention_mask == 0] = -100
    return lab


def causal_lm_loss(sd, cfg, llama_cfg, vision, input_ids, attention_mask, labels):
    """HF LlamaForCausalLM.forward(inputs_embeds=[image tokens | text], labels=[-100 x n_img | labels]).loss (HF:487-491 ForCausalLMLoss:
    logits upcast to fp32, position t scored against token t + 1, ignore_index -100, mean over the labelled positions)."""
    fh, img = mllm_forward(sd, cfg, llama_cfg, vision, input_ids, attention_mask)
    lp = find_llm_prefix(sd)
    logits = linear(fh, sd[lp.rsplit(".", 2)[0] + ".lm_head.weight"]).float()       # "...model." (Llama) / "...transformer." (GPT-2) -> sibling lm_head
    B, Q = img.shape[:2]
    fused_labels = torch.cat([torch.full((B, Q), -100, dtype=labels.dtype), labels], dim=1)
    return torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), fused_labels[:, 1:].reshape(-1), ignore_index=-100)


def stage1_loss_and_grads(sd, cfg, llama_cfg, vision, input_ids, attention_mask, labels, keys=None):
    """(loss, {key: gradient}) of the stage-1 objective for the trainable mllm.* tensors (autograd through the restatement)."""
    sd = {k: (v.detach().float().clone() if v.is_floating_point() else v) for k, v in sd.items()}
    keys = list(keys) if keys is not None else [k for k in trainable_keys(sd) if k.startswith("mllm.")]
    for k in keys:
        sd[k].requires_grad_(True)
    with torch.enable_grad():
        loss = causal_lm_loss(sd, cfg, llama_cfg, vision, input_ids, attention_mask, labels)
        grads = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    return loss.detach(), {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(keys, grads)}


def compress_grad(g, big=50_000):
    """Golden-fixture form of a gradient: full tensor when small, else first rows + row / column sums."""
    g = g.detach().float()
    if g.numel() <= big or g.dim() < 2:
        return {"full": g.clone()}
    g2 = g.reshape(g.shape[0], -1)
    return {"head": g2[:4].clone(), "rowsum": g2.double().sum(1).float(), "colsum": g2.double().sum(0).float(), "shape": tuple(g.shape)}


def best_of_k(candidates, y, norm_stat):
    """reference scripts/test.py:1336-1368: candidates (B, K, 2, T) -> (min ADE, min FDE, min RMSE) per scene."""
    B = candidates.shape[0]
    ns = torch.tensor(norm_stat, dtype=torch.float32) if not torch.is_tensor(norm_stat) else norm_stat.float()
    min_x, max_x, min_y, max_y = (ns[:, i].view(B, 1, 1) for i in range(4))
    rx, ry = max_x - min_x, max_y - min_y
    pred = candidates.clone().float()
    pred[..., 0, :] = pred[..., 0, :] * rx + min_x
    pred[..., 1, :] = pred[..., 1, :] * ry + min_y
    yd = y.clone().float().unsqueeze(1)
    yd[..., 0, :] = yd[..., 0, :] * rx + min_x
    yd[..., 1, :] = yd[..., 1, :] * ry + min_y
    errors = torch.sqrt(((pred - yd) ** 2).sum(dim=2))
    ade = errors.mean(dim=-1)
    fde = errors[..., -1]
    rmse = torch.sqrt(torch.mean((pred - yd) ** 2, dim=[2, 3]))
    return ade.min(dim=1).values, fde.min(dim=1).values, rmse.min(dim=1).values
