"""Backbone shape presets and model-constructor presets for BASELINE.json's configs.

The reference hard-codes a hub name (reference scripts/train.py:1347) and lets HF resolve the shape; there is
no network here, so a `base_model_name` is resolved to one of these shapes (or to an explicit dict)."""

LLAMA_PRESETS = {
    # "GPT-2-small-class" Llama-arch backbone: cfg 1, 2, 5
    "llama-768": dict(vocab_size=32000, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                      num_attention_heads=12, num_key_value_heads=12, head_dim=64, rms_norm_eps=1e-6,
                      rope_theta=10000.0),
    # Llama-2-7B shape (reference default name, scripts/train.py:461): cfg 3, 4
    "llama-7b": dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                     num_attention_heads=32, num_key_value_heads=32, head_dim=128, rms_norm_eps=1e-5,
                     rope_theta=10000.0),
    # GQA fidelity shape (Llama-3.2-1B geometry with plain rope)
    "llama-1b-gqa": dict(vocab_size=32000, hidden_size=2048, intermediate_size=8192, num_hidden_layers=16,
                         num_attention_heads=32, num_key_value_heads=8, head_dim=64, rms_norm_eps=1e-5,
                         rope_theta=500000.0),
    # meta-llama/Llama-3.2-1B — the one backbone name the reference hard-codes (scripts/train.py:1347).  Values from the public
    # model card / config.json: GQA 32/8, head_dim 64, rope_theta 5e5 with the "llama3" frequency scaling,
    # tied input / output embeddings, vocabulary 128256.
    "llama-3.2-1b": dict(vocab_size=128256, hidden_size=2048, intermediate_size=8192, num_hidden_layers=16,
                         num_attention_heads=32, num_key_value_heads=8, head_dim=64, rms_norm_eps=1e-5,
                         rope_theta=500000.0, max_position_embeddings=131072, tie_word_embeddings=True,
                         rope_scaling=dict(rope_type="llama3", factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                           original_max_position_embeddings=8192)),
    # GPT-2 architecture (HF GPT2LMHeadModel through the same AutoModelForCausalLM call, reference scripts/train.py:427-431; peft's default
    # LoRA target for model_type "gpt2" is the fused c_attn projection): LayerNorm with bias, learned position embeddings, Conv1D
    # projections ([in, out] weights) with biases, gelu_new MLP, tied lm_head.  Values of the public gpt2 (small) config.json.
    "gpt2": dict(arch="gpt2", vocab_size=50257, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                 n_positions=1024, layer_norm_epsilon=1e-5, tie_word_embeddings=True, attn_pdrop=0.1, resid_pdrop=0.1, embd_pdrop=0.1),
    "gpt2-tiny": dict(arch="gpt2", vocab_size=97, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      n_positions=64, layer_norm_epsilon=1e-5, tie_word_embeddings=True, attn_pdrop=0.1, resid_pdrop=0.1, embd_pdrop=0.1),
    # tiny shape for golden fixtures that carry their full state_dict
    "llama-tiny": dict(vocab_size=97, hidden_size=128, intermediate_size=256, num_hidden_layers=2,
                       num_attention_heads=4, num_key_value_heads=2, head_dim=32, rms_norm_eps=1e-6,
                       rope_theta=10000.0),
}
_ALIASES = {
    "meta-llama/Llama-2-7b-hf": "llama-7b", "meta-llama/Llama-7B": "llama-7b", "meta-llama/Llama-2-7b": "llama-7b",
    "huggyllama/llama-7b": "llama-7b",
    "meta-llama/Llama-3.2-1B": "llama-3.2-1b", "meta-llama/Llama-3.2-1B-Instruct": "llama-3.2-1b",
    "gpt2-small-class": "llama-768",
    "openai-community/gpt2": "gpt2", "gpt2-small": "gpt2",
}


def rope_inv_freq(cfg):
    """inv_freq[dh/2] (fp32, host) exactly as transformers computes it: HF:86-88 for the default rope and
    `modeling_rope_utils._compute_llama3_parameters` for rope_scaling = {"rope_type": "llama3", ...} (Llama-3.1 / 3.2:
    wavelengths longer than original_ctx / low_freq_factor are divided by `factor`, those shorter than
    original_ctx / high_freq_factor are kept, the band in between is interpolated)."""
    import math

    import torch
    dh = cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"]
    theta = float(cfg.get("rope_theta", 10000.0))
    inv = 1.0 / (theta ** (torch.arange(0, dh, 2, dtype=torch.int64).to(dtype=torch.float) / dh))
    rs = cfg.get("rope_scaling") or None
    if not rs or rs.get("rope_type", rs.get("type", "default")) == "default":
        return inv
    kind = rs.get("rope_type", rs.get("type"))
    if kind != "llama3":
        raise NotImplementedError(f"rope_scaling type {kind!r} (only 'default' and 'llama3' are implemented)")
    factor, lo, hi = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"]
    old_ctx = rs["original_max_position_embeddings"]
    low_wl, high_wl = old_ctx / lo, old_ctx / hi
    wavelen = 2 * math.pi / inv
    inv_l = torch.where(wavelen > low_wl, inv / factor, inv)
    smooth = (old_ctx / wavelen - lo) / (hi - lo)
    smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
    is_medium = ~(wavelen < high_wl) * ~(wavelen > low_wl)
    return torch.where(is_medium, smoothed, inv_l)


def resolve_llama(name_or_cfg):
    if isinstance(name_or_cfg, dict):
        return dict(name_or_cfg)
    key = _ALIASES.get(name_or_cfg, name_or_cfg)
    if key not in LLAMA_PRESETS:
        raise KeyError(f"unknown backbone {name_or_cfg!r}: pass a preset name {sorted(LLAMA_PRESETS)} or a shape dict "
                       "(there is no hub access to resolve checkpoint names)")
    return dict(LLAMA_PRESETS[key])


# ctor kwargs of MultiModalTrajectoryModel (reference scripts/train.py:848-872, values from 1100-1125)
MODEL_PRESETS = {
    "cfg1": dict(seq_len=15, out_len=25, individual=True, d_model=64, base_model_name="llama-768", use_lora=True,
                 lora_r=8, lora_alpha=32, ltsf_nhead=2),
    # the same model on the GPT-2 architecture itself (HF gpt2-small: LayerNorm, wpe, Conv1D projections, gelu_new, c_attn LoRA)
    "cfg1-gpt2": dict(seq_len=15, out_len=25, individual=True, d_model=64, base_model_name="gpt2", use_lora=True,
                      lora_r=8, lora_alpha=32, ltsf_nhead=2),
    "cfg3": dict(seq_len=15, out_len=25, individual=True, d_model=64, base_model_name="llama-7b", use_lora=True,
                 lora_r=16, lora_alpha=32, ltsf_nhead=2),
    "cfg5": dict(seq_len=15, out_len=50, individual=True, d_model=64, base_model_name="llama-768", use_lora=False,
                 ltsf_nhead=2),
    "tiny": dict(seq_len=6, out_len=12, individual=True, d_model=64, base_model_name="llama-tiny", use_lora=True,
                 lora_r=4, lora_alpha=16, q_hidden_size=64, q_nhead=4, q_enc_layers=1, q_dec_layers=2,
                 q_num_query_tokens=16, vision_dim=32, ltsf_nhead=2),
}
