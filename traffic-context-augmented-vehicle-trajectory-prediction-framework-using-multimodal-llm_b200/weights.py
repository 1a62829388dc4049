"""Deterministic, library-independent weight filler: every tensor of a reference-layout state_dict is drawn
from a generator seeded by (seed, crc32(key)), so the authoring container (reference model) and the GPU box
(this package) materialise bit-identical weights without shipping them.  Zero-initialised reference
parameters (pos_embedding, pos_encoding, lora_B) get small normals so they are exercised."""
import zlib

import torch


def _scale_for(key: str, shape) -> tuple:
    """-> (mean, std)"""
    leaf = key.rsplit(".", 1)[-1]
    is_norm = ("norm" in key.split(".")[-2] if "." in key else False) or key.endswith("fusion_layer.0.weight") \
        or key.endswith("fusion_layer.0.bias")
    if is_norm:
        return (1.0, 0.1) if leaf == "weight" else (0.0, 0.05)
    if leaf == "bias" or key.endswith("in_proj_bias"):
        return 0.0, 0.02
    if "lora_B" in key:
        return 0.0, 0.05
    if key.endswith("embed_tokens.weight"):
        return 0.0, 0.5
    if "pos_embedding" in key or "pos_encoding" in key:
        return 0.0, 0.1
    if "modality_embedding" in key or key.endswith("query_tokens"):
        return 0.0, 1.0
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return 0.0, fan_in ** -0.5
    return 0.0, 0.02


@torch.no_grad()
def deterministic_fill_(state_dict, seed: int = 0):
    """Fills the (floating-point) tensors of `state_dict` IN PLACE and returns it."""
    for key in sorted(state_dict):
        t = state_dict[key]
        if not t.is_floating_point():
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
        mean, std = _scale_for(key, tuple(t.shape))
        v = torch.randn(tuple(t.shape), generator=g, dtype=torch.float32) * std + mean
        t.copy_(v.to(t.dtype))
    return state_dict
