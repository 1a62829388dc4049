"""TEST INFRASTRUCTURE — not product code (see oracle/README.md).

Dropout for the oracle side of the fine-tune / best-of-K parity tests.  The reference trains with lora_dropout = ltsf_dropout = 0.1 and
torch's default 0.1 inside every nn.TransformerEncoderLayer / DecoderLayer (reference scripts/train.py:358, 402-405, 432-440, 663-671,
745-754; im_kim_train_GRN.py:1019-1041 runs in train() mode) and evaluates best-of-K in train() mode as well (scripts/test.py:1308-1338).
torch's own RNG stream cannot be reproduced by another implementation, so parity is defined with the mask function of
include/tcavp.h (tcavp_dropout) substituted for torch's Bernoulli draw on BOTH sides:

  * `DropOracle` evaluates that mask function on the host (numpy, 32-bit wrap-around arithmetic);
  * `patch_reference_dropout` makes the UNMODIFIED reference model use it: torch.nn.functional.dropout (every nn.Dropout, peft's
    lora_dropout, nn.MultiheadAttention's need_weights=True path) and F.scaled_dot_product_attention (the need_weights=False path of the
    nn.Transformer layers) are replaced for the duration of one forward; the k-th call is mapped to its site by `reference_call_sequence`,
    which is also checked to be consumed exactly (so a wrong assumption about the reference's call order fails loudly);
  * oracle/restated.py applies the same masks at the same sites (restated.DROP).

Site ids and element indices are the contract shared with the CUDA path (train_engine.py: MOD / KIND / site_id):
  site = module << 20 | layer << 8 | kind;  activations: index = row * cols + col with row = b * T + t (batch-major);
  attention probabilities: index = ((b * H + h) * Tq + i) * Tk + j.
"""
import contextlib

import numpy as np
import torch

MOD = dict(poly=1, qenc=2, qdec=3, llm=4, ltsf=5, dec=6)
KIND = dict(sa_attn=0, drop1=1, ffn=2, drop2=3, ca_attn=4, drop3=5, post=6, cross_attn=7, lora_q=8, lora_k=9, lora_v=10, lora_c=11, embd=12, attn=13, resid1=14, resid2=15)
_M32 = np.uint64(0xFFFFFFFF)


def site_id(mod, layer, kind):
    return (MOD[mod] << 20) | (layer << 8) | KIND[kind]


def _mix32(x):
    """x: uint64 array holding 32-bit values."""
    x = x ^ (x >> np.uint64(16))
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x = x ^ (x >> np.uint64(15))
    x = (x * np.uint64(0x846CA68B)) & _M32
    x = x ^ (x >> np.uint64(16))
    return x


def drop_threshold(p):
    return min(max(int(round(float(p) * 4294967296.0)), 0), 4294967295)


def keep_mask(seed, step, site, p, n):
    """bool[n]: element idx is kept iff hash(idx; seed, step, site) >= round(p * 2^32)."""
    s0, s1 = np.uint64(int(seed) & 0xFFFFFFFF), np.uint64(int(step) & 0xFFFFFFFF)
    inner = _mix32(np.array([(int(site) + 0x9E3779B9 * (int(s1) + 1)) & 0xFFFFFFFF], dtype=np.uint64))
    key = _mix32(s0 ^ inner)[0]
    idx = np.arange(n, dtype=np.uint64)
    u = _mix32((idx & _M32) ^ key)
    hi = idx >> np.uint64(32)
    if n > 0xFFFFFFFF:
        u = np.where(hi != 0, _mix32(u ^ ((hi * np.uint64(0x85EBCA6B)) & _M32)), u)
    return u >= np.uint64(drop_threshold(p))


class DropOracle:
    """probs: {(module, kind): p}.  `apply(t, mod, layer, kind)` = t * keep / (1 - p) with t in the canonical (batch-major) layout."""

    def __init__(self, seed, step, probs):
        self.seed, self.step, self.probs = int(seed), int(step), dict(probs)
        self.used = []

    def factor(self, shape, mod, layer, kind, p=None, dtype=torch.float32):
        p = self.probs.get((mod, kind), 0.0) if p is None else p
        if p <= 0.0:
            return None
        n = int(np.prod(shape))
        keep = keep_mask(self.seed, self.step, site_id(mod, layer, kind), p, n)
        self.used.append((mod, layer, kind))
        return torch.from_numpy(keep.astype(np.float64) / (1.0 - p)).to(dtype).view(*shape)

    def apply(self, t, mod, layer, kind, p=None):
        f = self.factor(tuple(t.shape), mod, layer, kind, p, t.dtype)
        return t if f is None else t * f


def default_probs(model_cfg, transformer_p=0.1, llama_cfg=None):
    """Dropout probability of every site for a reference constructor-kwargs dict (train.py:848-872 defaults: 0.1 everywhere)."""
    lp, tp = float(model_cfg.get("lora_dropout", 0.1)), float(model_cfg.get("ltsf_dropout", 0.1))
    pr = {}
    for mod in ("poly", "qenc", "qdec"):
        for kind in ("sa_attn", "drop1", "ffn", "drop2") + (("ca_attn", "drop3") if mod == "qdec" else ()):
            pr[(mod, kind)] = float(transformer_p)
    for kind in ("sa_attn", "drop1", "ffn", "drop2"):
        pr[("ltsf", kind)] = tp
    pr[("dec", "post")] = pr[("dec", "cross_attn")] = tp
    if model_cfg.get("use_lora", True):
        pr[("llm", "lora_q")] = pr[("llm", "lora_v")] = pr[("llm", "lora_c")] = lp      # lora_c: the c_attn target of a GPT-2-arch backbone
    if llama_cfg is not None and llama_cfg.get("arch") == "gpt2":      # HF GPT2Config's own dropouts
        pr[("llm", "embd")], pr[("llm", "attn")] = float(llama_cfg.get("embd_pdrop", 0.0)), float(llama_cfg.get("attn_pdrop", 0.0))
        pr[("llm", "resid1")] = pr[("llm", "resid2")] = float(llama_cfg.get("resid_pdrop", 0.0))
    return pr


def reference_call_sequence(model_cfg, n_llm_layers, lora_targets=("q_proj", "v_proj"), arch="llama", llama_cfg=None):
    """[(module, layer, kind, layout)] in the order the reference's forward reaches its dropout calls (train mode):
    lane_polygon_encoder -> mllm (Q-Former encoder, decoder, LoRA-Llama) -> ltsf (attention block, decoder) — train.py:926-939.
    layout: "flat" = the tensor's own row-major order is the canonical one; "tbe" = (T, B, E) tensors of the batch_first=False
    modules (SelfAttentionBlock, train.py:674-686)."""
    seq = []

    def enc(mod, n):
        for l in range(n):
            seq.extend([(mod, l, "sa_attn", "flat"), (mod, l, "drop1", "flat"), (mod, l, "ffn", "flat"), (mod, l, "drop2", "flat")])
    enc("poly", model_cfg.get("lane_polygon_layers", 2))
    enc("qenc", model_cfg.get("q_enc_layers", 4))
    for l in range(model_cfg.get("q_dec_layers", 4)):        # torch TransformerDecoderLayer: _sa_block, _mha_block, _ff_block
        seq.extend([("qdec", l, "sa_attn", "flat"), ("qdec", l, "drop1", "flat"), ("qdec", l, "ca_attn", "flat"), ("qdec", l, "drop2", "flat"),
                    ("qdec", l, "ffn", "flat"), ("qdec", l, "drop3", "flat")])
    if arch == "gpt2":
        # HF modeling_gpt2.py in train mode: GPT2Model.drop; per block c_attn (peft's lora_dropout on its input), attention-probability
        # dropout, resid_dropout after attn.c_proj, GPT2MLP.dropout after mlp.c_proj.  Calls with p == 0 never reach the patched function.
        g = llama_cfg or {}
        lora = model_cfg.get("use_lora", True) and float(model_cfg.get("lora_dropout", 0.1)) > 0
        if float(g.get("embd_pdrop", 0.0)) > 0:
            seq.append(("llm", 0, "embd", "flat"))
        for l in range(n_llm_layers):
            if lora:
                seq.append(("llm", l, "lora_c", "flat"))
            if float(g.get("attn_pdrop", 0.0)) > 0:
                seq.append(("llm", l, "attn", "flat"))
            if float(g.get("resid_pdrop", 0.0)) > 0:
                seq.extend([("llm", l, "resid1", "flat"), ("llm", l, "resid2", "flat")])
    elif model_cfg.get("use_lora", True):
        for l in range(n_llm_layers):                         # HF LlamaAttention.forward: q_proj, k_proj, v_proj in this order
            for t in ("q_proj", "k_proj", "v_proj"):
                if t in lora_targets:
                    seq.append(("llm", l, "lora_" + t[0], "flat"))
    seq.extend([("ltsf", 0, "sa_attn", "flat"), ("ltsf", 0, "drop1", "tbe"), ("ltsf", 0, "ffn", "tbe"), ("ltsf", 0, "drop2", "tbe")])
    if model_cfg.get("use_post_mlp", True):
        seq.append(("dec", 0, "post", "flat"))
    seq.append(("dec", 0, "cross_attn", "flat"))
    return seq


@contextlib.contextmanager
def patch_reference_dropout(oracle, sequence):
    """Runs the unmodified reference model with `oracle`'s masks: see the module docstring."""
    import torch.nn.functional as F
    it = iter(sequence)
    state = {"calls": 0}
    orig_dropout, orig_sdpa = F.dropout, F.scaled_dot_product_attention

    def dropout(input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        try:
            mod, layer, kind, layout = next(it)
        except StopIteration:
            raise AssertionError("the reference made more dropout calls than reference_call_sequence lists") from None
        state["calls"] += 1
        if layout == "tbe":
            T, B, E = input.shape
            f = oracle.factor((B, T, E), mod, layer, kind, p, input.dtype).permute(1, 0, 2)
        else:
            f = oracle.factor(tuple(input.shape), mod, layer, kind, p, input.dtype)
        return input * f

    def sdpa(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, **kw):
        if dropout_p == 0.0:
            return orig_sdpa(query, key, value, attn_mask=attn_mask, dropout_p=0.0, is_causal=is_causal, scale=scale, **kw)
        s = (query @ key.transpose(-2, -1)) * (query.shape[-1] ** -0.5 if scale is None else scale)
        if is_causal:
            L, S = s.shape[-2:]
            s = s.masked_fill(~torch.ones(L, S, dtype=torch.bool).tril(), float("-inf"))
        if attn_mask is not None:
            s = s.masked_fill(~attn_mask, float("-inf")) if attn_mask.dtype == torch.bool else s + attn_mask
        p = dropout(torch.softmax(s, dim=-1), dropout_p, True)
        return p @ value

    F.dropout, F.scaled_dot_product_attention = dropout, sdpa
    try:
        yield state
        leftover = list(it)
        assert not leftover, f"reference_call_sequence lists {len(leftover)} dropout calls the reference never made: {leftover[:3]}"
    finally:
        F.dropout, F.scaled_dot_product_attention = orig_dropout, orig_sdpa
