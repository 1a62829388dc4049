"""TEST INFRASTRUCTURE — mints tests/golden/*.pt by running the UNMODIFIED reference classes
(/root/reference/scripts/train.py via oracle/ref_loader.py) in the authoring container.

    python -m oracle.make_golden            # writes every fixture listed in FIXTURES

Weights are not stored: both sides materialise them with tcavp_b200.deterministic_fill_(state_dict, seed)
(a crc32(key)-seeded CPU generator), and each fixture records a few per-key checksums so drift of the filler
is detected.  Inputs ARE stored (they are small).  The reference ships no golden vectors of its own
(SURVEY.md §4), so these files are the pin.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import tcavp_b200 as T  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (model preset, overrides, B, l_text, weight seed, scene seed, explicit poly_len or None, script)
FIXTURES = {
    "tiny_b6": ("tiny", {}, 6, 24, 11, 101, [0, 64, 1, 33, 14, 22], "scripts/train.py"),
    "cfg1_b8": ("cfg1", {}, 8, 128, 7, 3, None, "scripts/train.py"),
    "cfg5_b32": ("cfg5", {}, 32, 128, 5, 9, None, "scripts/train.py"),
    # 7B-class geometry (H 4096, 32 heads of 128, I 11008, LoRA r 16) cut to 2 decoder layers and a 4096-entry vocabulary so the
    # unmodified reference finishes on the authoring container's CPU (~10 min); 16 scenes = 2304 rows reach the CTA-pair GEMM.
    "cfg3l2_b16": ("cfg3", {"base_model_name": dict(vocab_size=4096, hidden_size=4096, intermediate_size=11008, num_hidden_layers=2,
                                                   num_attention_heads=32, num_key_value_heads=32, head_dim=128, rms_norm_eps=1e-5,
                                                   rope_theta=10000.0)}, 16, 128, 13, 17, None, "scripts/train.py"),
    # grouped-query attention (Llama-3.2-1B geometry: 32 query heads on 8 KV heads, rope_theta 5e5), two layers, ragged text
    "gqa_l2_b32": ("cfg1", {"lora_r": 16, "base_model_name": dict(vocab_size=4096, hidden_size=2048, intermediate_size=8192, num_hidden_layers=2,
                                                                 num_attention_heads=32, num_key_value_heads=8, head_dim=64, rms_norm_eps=1e-5,
                                                                 rope_theta=500000.0)}, 32, 96, 23, 23, None, "scripts/train.py"),   # weight seed picked so FDE (645 px) is not degenerate against ADE (411 px)
    # the reference's own hard-coded backbone (scripts/train.py:1347 "meta-llama/Llama-3.2-1B"): GQA 32/8, rope_theta 5e5 with the
    # llama3 frequency scaling (through the real HF LlamaRotaryEmbedding), TIED input / output embeddings, LoRA r 8 (the reference's
    # default) — cut to two layers and an 8192-entry vocabulary so the fixture stays small and fast
    "llama32_1b_l2_b8": ("cfg1", {"base_model_name": dict(T.LLAMA_PRESETS["llama-3.2-1b"], num_hidden_layers=2, vocab_size=8192)},
                         8, 128, 31, 37, None, "scripts/train.py"),
    # GPT-2 architecture through the same AutoModelForCausalLM call (HF GPT2LMHeadModel; peft's default LoRA target c_attn, a Conv1D):
    # the tiny shape with a full state_dict, and gpt2-small's geometry (768 / 12 heads / 3072, LayerNorm, wpe, gelu_new) cut to two
    # layers and an 8192-entry vocabulary
    "gpt2_tiny_b6": ("tiny", {"base_model_name": "gpt2-tiny"}, 6, 24, 41, 43, [0, 64, 1, 33, 14, 22], "scripts/train.py"),
    "gpt2_l2_b8": ("cfg1", {"base_model_name": dict(T.LLAMA_PRESETS["gpt2"], num_hidden_layers=2, vocab_size=8192)},
                   8, 128, 47, 53, None, "scripts/train.py"),
}


def weight_checksums(sd, n=12):
    keys = sorted(k for k in sd if sd[k].is_floating_point())
    pick = keys[:: max(1, len(keys) // n)][:n]
    return {k: float(sd[k].double().sum()) for k in pick}


def make(name):
    preset, over, B, l_text, wseed, sseed, poly_len, script = FIXTURES[name]
    mc = dict(T.MODEL_PRESETS[preset])
    mc.update(over)
    lc = T.resolve_llama(mc["base_model_name"])
    mod = ref_loader.load_reference(script, lc)
    model = ref_loader.build_reference_model(mod, mc, lc)
    sd = model.state_dict()
    T.deterministic_fill_(sd, wseed)
    model.load_state_dict(sd, strict=True)
    s = T.make_scenes(B, mc["seq_len"], mc["out_len"], vision_dim=mc.get("vision_dim", 512), l_text=l_text,
                      vocab=lc["vocab_size"], seed=sseed)
    if poly_len is not None:
        g = torch.Generator().manual_seed(sseed + 1)
        s["poly_len"] = list(poly_len)
        pts = torch.rand(B, 64, 2, generator=g) * torch.tensor([3839.0, 750.0]) + torch.tensor([0.0, 700.0])
        keep = torch.arange(64)[None, :] < torch.tensor(poly_len)[:, None]
        s["polygon"] = torch.where(keep[..., None], pts, torch.zeros_like(pts))
    with torch.no_grad():
        loss, decoded = model(s["x"], s["vision"], s["context_str"], s["polygon"], s["poly_len"], y=s["y"],
                              norm_stat=s["norm_stat"], input_ids=s["input_ids"], attention_mask=s["attention_mask"],
                              labels=None)
        decoded_only = model(s["x"], s["vision"], s["context_str"], s["polygon"], s["poly_len"],
                             input_ids=s["input_ids"], attention_mask=s["attention_mask"])
        assert torch.equal(decoded, decoded_only)
        poly_emb = model.lane_polygon_encoder(s["polygon"], s["poly_len"])
        image_tokens = model.mllm.qformer(s["vision"])
        final_hidden, n_img = model.mllm(s["vision"], s["context_str"], input_ids=s["input_ids"],
                                         attention_mask=s["attention_mask"])
        lt = model.ltsf
        enc = lt.attn_block(lt.nlinear_encoder(lt.token_proj(s["x"])) + lt.pos_encoding)
        # ADE / FDE exactly as reference scripts/train.py:1302-1322
        ns = torch.tensor(s["norm_stat"])
        pred_den, y_den = decoded.clone(), s["y"].clone()
        rx, ry = (ns[:, 1] - ns[:, 0])[:, None], (ns[:, 3] - ns[:, 2])[:, None]
        pred_den[:, 0, :] = pred_den[:, 0, :] * rx + ns[:, 0][:, None]
        pred_den[:, 1, :] = pred_den[:, 1, :] * ry + ns[:, 2][:, None]
        y_den[:, 0, :] = y_den[:, 0, :] * rx + ns[:, 0][:, None]
        y_den[:, 1, :] = y_den[:, 1, :] * ry + ns[:, 2][:, None]
        err = torch.sqrt(((pred_den - y_den) ** 2).sum(dim=1))
        ade, fde = err.mean(dim=1), err[:, -1]
    keep_fh = min(B, 2)
    fix = {
        "name": name, "model_cfg": mc, "llama_cfg": lc, "weight_seed": wseed, "scene_seed": sseed,
        "l_text": l_text, "script": script, "n_state_keys": len(sd), "state_shapes": {k: tuple(v.shape) for k, v in sd.items()},
        "weight_checksums": weight_checksums(sd),
        "inputs": {k: s[k] for k in ("x", "y", "vision", "polygon", "poly_len", "norm_stat", "input_ids", "attention_mask")},
        "out": {
            "decoded": decoded, "loss": loss, "poly_emb": poly_emb, "image_tokens": image_tokens[:4].clone(), "enc": enc,
            "final_hidden_head": final_hidden[:keep_fh].clone(),
            "final_hidden_rowmean": final_hidden.mean(dim=-1), "final_hidden_absmean": final_hidden.abs().mean(dim=-1),
            "ade": ade, "fde": fde, "n_img": int(n_img),
        },
        "versions": {"torch": str(torch.__version__), "transformers": __import__("transformers").__version__},
    }
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".pt")
    torch.save(fix, path)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  loss={float(loss):.4f} ade={float(ade.mean()):.4f} fde={float(fde.mean()):.4f}")


# name -> (model preset, B, l_text, polygon lengths[, ctor overrides])
GRAD_FIXTURES = {
    "tiny_b5_grads": ("tiny", 5, 24, [64, 1, 33, 14, 22]),
    # the real 768-class geometry (12 layers, 12 heads of 64, L = 144, LoRA r 8): gradients are stored compressed (restated.compress_grad)
    "cfg1_b3_grads": ("cfg1", 3, 128, [33, 14, 22]),
    # 7B-class geometry (H 4096, 32 heads of 128, I 11008, LoRA r 16), two decoder layers: the backward through W^T at K = 11008 / 22016
    # (the wide GEMM) and the 67 M-parameter cross-attention of the fusion block
    "cfg3l2_b4_grads": ("cfg3", 4, 128, [33, 14, 22, 32], {"base_model_name": FIXTURES["cfg3l2_b16"][1]["base_model_name"]}),
    # GPT-2 architecture (HF GPT2LMHeadModel, peft LoRA on the fused c_attn Conv1D): the tiny shape and gpt2-small's geometry cut to two layers
    "gpt2_tiny_b5_grads": ("tiny", 5, 24, [64, 1, 33, 14, 22], {"base_model_name": "gpt2-tiny"}),
    "gpt2_l2_b3_grads": ("cfg1", 3, 128, [33, 14, 22], {"base_model_name": FIXTURES["gpt2_l2_b8"][1]["base_model_name"]}),
}


# train-mode fixtures: name -> (p = 0 fixture it shares model / inputs with, dropout seed, step)
DROP_FIXTURES = {
    "tiny_b5_grads_drop": ("tiny_b5_grads", 20261, 3),
    "cfg1_b3_grads_drop": ("cfg1_b3_grads", 77001, 11),
    "gpt2_tiny_b5_grads_drop": ("gpt2_tiny_b5_grads", 31337, 5),
    "gpt2_l2_b3_grads_drop": ("gpt2_l2_b3_grads", 52007, 2),       # LoRA r 8, hidden 768: the fused lora-dropout kernels in bf16 mode
}


def make_grads(name="tiny_b5_grads"):
    """Gradients of the UNMODIFIED reference model (autograd on): the pin of the fine-tune step (reference
    scripts/im_kim_train_GRN.py:1029-1039).  Plain fixtures run in eval mode (every dropout off); `*_drop` fixtures run in train()
    mode — the reference's lora_dropout = ltsf_dropout = 0.1 and torch's 0.1 in every nn.Transformer layer — with torch's Bernoulli
    draw replaced by the counter-based mask of include/tcavp.h (oracle/dropout.py: patch_reference_dropout)."""
    from oracle import dropout as OD
    from oracle import restated
    drop = None
    if name in DROP_FIXTURES:
        base, dseed, dstep = DROP_FIXTURES[name]
        drop = (dseed, dstep)
    else:
        base = name
    preset, B, l_text, poly_len, *over = GRAD_FIXTURES[base]
    mc = dict(T.MODEL_PRESETS[preset])
    mc.update(over[0] if over else {})
    lc = T.resolve_llama(mc["base_model_name"])
    mod = ref_loader.load_reference("scripts/train.py", lc)
    model = ref_loader.build_reference_model(mod, mc, lc)
    sd = model.state_dict()
    T.deterministic_fill_(sd, 13)
    model.load_state_dict(sd, strict=True)
    s = T.make_scenes(B, mc["seq_len"], mc["out_len"], vision_dim=mc.get("vision_dim", 512), l_text=l_text, vocab=lc["vocab_size"], seed=77)
    poly_len = list(poly_len)             # no empty polygon: the reference's own gradients are NaN there (all-masked softmax)
    g = torch.Generator().manual_seed(78)
    pts = torch.rand(B, 64, 2, generator=g) * torch.tensor([3839.0, 750.0]) + torch.tensor([0.0, 700.0])
    keep = torch.arange(64)[None, :] < torch.tensor(poly_len)[:, None]
    s["poly_len"], s["polygon"] = poly_len, torch.where(keep[..., None], pts, torch.zeros_like(pts))
    # The reference's own fp32 backward is noise-dominated in the first lane-polygon attention (raw-pixel inputs give logits
    # ~1e6: its fp32 gradients of pos_embedding / input_proj / layer-0 in_proj differ by 20-30 % from its fp64 gradients, while
    # every other tensor agrees to 1e-6).  The pin is therefore the UNMODIFIED reference run in float64.
    import contextlib
    n_llm = lc["num_hidden_layers"]

    def patched():
        if drop is None:
            return contextlib.nullcontext()
        return OD.patch_reference_dropout(OD.DropOracle(drop[0], drop[1], {}), OD.reference_call_sequence(mc, n_llm, arch=lc.get("arch", "llama"), llama_cfg=lc))
    if drop is not None:
        model.train()
    with torch.no_grad(), patched():
        loss32, _ = model(s["x"], s["vision"], s["context_str"], s["polygon"], s["poly_len"], y=s["y"], norm_stat=s["norm_stat"],
                          input_ids=s["input_ids"], attention_mask=s["attention_mask"], labels=None)
    model = model.double()
    model.zero_grad()
    with patched() as st:
        loss, decoded = model(s["x"].double(), s["vision"].double(), s["context_str"], s["polygon"].double(), s["poly_len"], y=s["y"].double(),
                              norm_stat=s["norm_stat"], input_ids=s["input_ids"], attention_mask=s["attention_mask"], labels=None)
    if drop is not None:
        print(f"  train mode: {st['calls']} dropout calls mapped to sites")
    loss.backward()
    assert abs(float(loss) - float(loss32)) < 1e-4 * float(loss), (float(loss), float(loss32))
    grads = {k: p.grad for k, p in model.named_parameters() if p.requires_grad and p.grad is not None}
    frozen = sorted(k for k, p in model.named_parameters() if not p.requires_grad)
    assert all("llama_model" in k and "lora_" not in k for k in frozen), frozen[:3]
    fix = {"name": name, "model_cfg": mc, "llama_cfg": lc, "weight_seed": 13,
           "inputs": {k: s[k] for k in ("x", "y", "vision", "polygon", "poly_len", "norm_stat", "input_ids", "attention_mask")},
           "loss": loss.detach().float(), "loss_fp32_run": loss32.detach(), "decoded": decoded.detach().float(), "n_trainable": len(grads),
           "precision": "reference executed in float64 (see make_grads)",
           "dropout": None if drop is None else {"seed": drop[0], "step": drop[1], "probs": {f"{k[0]}.{k[1]}": v for k, v in OD.default_probs(mc, llama_cfg=lc).items()},
                                                  "mode": "train(): counter-based masks of include/tcavp.h substituted for torch's RNG"},
           "grads": {k: restated.compress_grad(v) for k, v in grads.items()},
           "versions": {"torch": str(torch.__version__), "transformers": __import__("transformers").__version__}}
    path = os.path.join(GOLDEN_DIR, name + ".pt")
    torch.save(fix, path)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  loss={float(loss.detach()):.4f}  {len(grads)} gradient tensors")


# stage-1 fixtures: name -> (model preset, B, l_text, prompt length)
STAGE1_FIXTURES = {
    "tiny_b5_stage1": ("tiny", 5, 24, 9),
    "cfg1_b2_stage1": ("cfg1", 2, 128, 70),       # the 768-class geometry, vocabulary 32000
    "gpt2_tiny_b5_stage1": ("tiny", 5, 24, 9, {"base_model_name": "gpt2-tiny"}),      # HF GPT2LMHeadModel (tied lm_head) under the same call
}


def make_stage1(name="tiny_b5_stage1"):
    """Stage-1 (CausalLM) objective of the UNMODIFIED reference classes (scripts/train.py:516-547 builds exactly this call; the stage-1
    script scripts/check_generation.py:131-151 returns its `outputs`): Q-Former image tokens + text embeddings (+ modality vectors) through
    the peft-wrapped HF LlamaForCausalLM with labels -> outputs.loss, autograd gradients of every trainable mllm tensor.  float64 run,
    eval mode (dropout off), like the other gradient fixtures."""
    from oracle import restated
    preset, B, l_text, prompt_len, *over = STAGE1_FIXTURES[name]
    mc = dict(T.MODEL_PRESETS[preset])
    mc.update(over[0] if over else {})
    lc = T.resolve_llama(mc["base_model_name"])
    mod = ref_loader.load_reference("scripts/train.py", lc)
    model = ref_loader.build_reference_model(mod, mc, lc)
    sd = model.state_dict()
    T.deterministic_fill_(sd, 13)
    model.load_state_dict(sd, strict=True)
    s = T.make_scenes(B, mc["seq_len"], mc["out_len"], vision_dim=mc.get("vision_dim", 512), l_text=l_text, vocab=lc["vocab_size"], seed=79)
    labels = restated.stage1_labels(s["input_ids"], s["attention_mask"], prompt_len)

    def ref_loss(m, vision):
        mm = m.mllm                                                   # the reference's LlamaMultiModal, its own sub-modules
        img = mm.q_proj(mm.qformer(vision)) + mm.vision_modality_embedding
        txt = mm.llama_wrapper.llama_model.get_input_embeddings()(s["input_ids"]) + mm.text_modality_embedding
        fused = torch.cat([img, txt], dim=1)
        fmask = torch.cat([torch.ones((B, img.size(1)), dtype=s["attention_mask"].dtype), s["attention_mask"]], dim=1)
        flab = torch.full((B, img.size(1) + labels.size(1)), -100, dtype=labels.dtype)
        flab[:, img.size(1):] = labels
        return mm.llama_wrapper(inputs_embeds=fused, attention_mask=fmask, labels=flab, output_hidden_states=True).loss
    model.eval()
    with torch.no_grad():
        loss32 = ref_loss(model, s["vision"])
    model = model.double()
    model.zero_grad()
    loss = ref_loss(model, s["vision"].double())
    loss.backward()
    assert abs(float(loss) - float(loss32)) < 1e-4 * float(loss), (float(loss), float(loss32))
    grads = {k: p.grad for k, p in model.named_parameters() if p.requires_grad and p.grad is not None}
    assert grads and all(k.startswith("mllm.") for k in grads), [k for k in grads if not k.startswith("mllm.")][:3]
    fix = {"name": name, "model_cfg": mc, "llama_cfg": lc, "weight_seed": 13,
           "inputs": {"vision": s["vision"], "input_ids": s["input_ids"], "attention_mask": s["attention_mask"], "labels": labels},
           "loss": loss.detach().float(), "loss_fp32_run": loss32.detach(), "n_tokens": int((labels != -100).sum()), "n_trainable": len(grads),
           "precision": "reference executed in float64 (see make_stage1)",
           "grads": {k: restated.compress_grad(v) for k, v in grads.items()},
           "versions": {"torch": str(torch.__version__), "transformers": __import__("transformers").__version__}}
    path = os.path.join(GOLDEN_DIR, name + ".pt")
    torch.save(fix, path)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  loss={float(loss.detach()):.4f}  {fix['n_tokens']} labelled tokens, {len(grads)} gradient tensors")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    for n in (sys.argv[1:] or list(FIXTURES) + list(GRAD_FIXTURES) + list(DROP_FIXTURES) + list(STAGE1_FIXTURES)):
        make_stage1(n) if n in STAGE1_FIXTURES else (make_grads(n) if "_grads" in n else make(n))
