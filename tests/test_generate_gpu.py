"""Text-generation side path (reference scripts/train.py:577-654): next-token logits over the image-token + prompt prefix against
the oracle's backbone + lm_head, a greedy continuation against the same loop run through the oracle, and the generate_batch call of
train.py:1231-1241 with a stand-in tokenizer.  Sampling itself depends on the RNG and is only checked for its invariants."""
import pytest
import torch

gpu = pytest.mark.gpu

import tcavp_b200 as T  # noqa: E402
from conftest import build_filled_model, load_golden  # noqa: E402
from oracle import restated  # noqa: E402
from tcavp_b200 import generate as G  # noqa: E402


def _oracle_logits(sd, cfg, lc, vision, ids_prefix, new_ids):
    """lm_head(hidden_states[-1][:, -1]) for prefix = image tokens (+vision modality) + prompt (+text modality) + plain new-token rows."""
    sd = {k: v.float() for k, v in sd.items()}
    p = "mllm."
    img = restated.qformer(sd, vision, cfg.get("q_nhead", 8), p + "qformer.")
    if p + "q_proj.weight" in sd:
        img = restated.linear(img, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    img = img + sd[p + "vision_modality_embedding"]
    lp = restated.find_llm_prefix(sd)
    gpt2 = lc.get("arch") == "gpt2"
    E = sd[lp + ("wte.weight" if gpt2 else "embed_tokens.weight")]
    parts = [img, E[ids_prefix] + sd[p + "text_modality_embedding"]]
    if new_ids is not None and new_ids.shape[1] > 0:
        parts.append(E[new_ids])
    fused = torch.cat(parts, dim=1)
    mask = torch.ones(fused.shape[:2], dtype=torch.long)
    fh = (restated.gpt2_stack if gpt2 else restated.llama_stack)(sd, lc, fused, mask, cfg.get("lora_alpha", 32) / cfg.get("lora_r", 8), lp)
    head = sd[lp.rsplit(".", 2)[0] + ".lm_head.weight"]
    return fh[:, -1] @ head.t()


@gpu
@pytest.mark.parametrize("fixture", ["tiny_b6", "gpt2_tiny_b6"])
@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-4), ("bf16", 4e-2)])
def test_next_token_logits_and_greedy_continuation_match_the_oracle(lib_built, dtype, tol, fixture):
    fix = load_golden(fixture)
    m = build_filled_model(fix, dtype, "cuda")
    cfg, lc = fix["model_cfg"], fix["llama_cfg"]
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    i = fix["inputs"]
    vision, ids = i["vision"][:2], i["input_ids"][:2, :9]
    eng = m.engine()
    pre = eng.prefix_embeds(vision, ids)
    assert pre.shape == (2, 16 + 9, lc["hidden_size"])
    got = G.next_token_logits(eng, pre).cpu()
    want = _oracle_logits(sd, cfg, lc, vision, ids, None)
    scale = float(want.abs().max())
    torch.testing.assert_close(got, want, rtol=tol, atol=tol * scale)
    if dtype == "fp32":
        # greedy continuation (no sampling, no penalties): the same tokens as the oracle's own loop, with the KV cache (prefill + one-row
        # decode steps) and with the sequence recomputed at every step
        new = torch.zeros(2, 0, dtype=torch.long)
        for _ in range(5):
            nxt = _oracle_logits(sd, cfg, lc, vision, ids, new).argmax(-1)
            new = torch.cat([new, nxt[:, None]], dim=1)
        for kv in (True, False):
            seq = G.generate_ids(m, vision, ids, max_new_tokens=5, do_sample=False, repetition_penalty=1.0, no_repeat_ngram_size=0, kv_cache=kv)
            assert seq.shape == (2, 9 + 5) and torch.equal(seq[:, :9].cpu(), ids)
            assert torch.equal(seq[:, 9:].cpu(), new), kv
    # the decode step itself (Llama: RoPE at the new position; GPT-2: wpe[t]): hidden state of position P + j from the cache == the same
    # row of a full forward over the longer sequence
    extra = eng.token_embeds(i["input_ids"][:2, 9:12].to(pre.device))
    full = torch.cat([pre, extra], dim=1)
    Lf, P, H = full.shape[1], pre.shape[1], pre.shape[2]
    ones = torch.ones(2, Lf, dtype=torch.int32, device=pre.device)
    want_h = eng.llm_forward(full.clone(), ones, 2, Lf).view(2, Lf, H).float().cpu()
    caches = []
    got0 = eng.llm_forward(pre.clone(), ones[:, :P].contiguous(), 2, P, kv_out=(caches, Lf)).view(2, P, H).float().cpu()
    assert len(caches) == lc["num_hidden_layers"] and caches[0].shape[1] == Lf
    ht = dict(rtol=tol, atol=tol * float(want_h.abs().max()))
    torch.testing.assert_close(got0, want_h[:, :P], **ht)
    for j in range(3):
        h = eng.llm_decode_step(extra[:, j].contiguous(), caches, P + j).float().cpu()
        torch.testing.assert_close(h, want_h[:, P + j], **ht)


@gpu
def test_generate_batch_call_of_the_reference_train_loop(lib_built):
    """train.py:1231-1241: model.mllm.generate_batch(vision_embs=..., prompt_ids=..., tokenizer=model.mllm.tokenizer, max_new_tokens=...,
    temperature=0.9, top_k=40, top_p=0.9, device=device) -> list of strings."""
    fix = load_golden("tiny_b6")
    m = build_filled_model(fix, "bf16", "cuda")
    i = fix["inputs"]

    class Tok:
        eos_token_id, pad_token_id = 96, 0

        def decode(self, ids, skip_special_tokens=True):
            return " ".join(f"t{int(t)}" for t in ids if not (skip_special_tokens and int(t) in (0, 96)))
    m.mllm.tokenizer = Tok()
    torch.manual_seed(0)
    texts = m.mllm.generate_batch(vision_embs=i["vision"][0:1].cuda(), prompt_ids=i["input_ids"][0:1, :12].cuda(), tokenizer=m.mllm.tokenizer,
                                  max_new_tokens=16, temperature=0.9, top_k=40, top_p=0.9, device="cuda")
    assert isinstance(texts, list) and len(texts) == 1 and isinstance(texts[0], str)
    toks = texts[0].split()
    assert 12 <= len(toks) <= 12 + 16 and all(t.startswith("t") for t in toks)


def test_logits_processors_follow_hf_semantics():
    """repetition penalty, no-repeat-ngram, top-k, top-p on a hand-made distribution (CPU tensors: pure torch host logic)."""
    logits = torch.tensor([[2.0, 1.0, 0.5, -1.0, 0.0, 3.0]])
    seq = torch.tensor([[5, 1, 5]])
    out = G.process_logits(logits.clone(), seq, 1.0, 0, None, 1.2, 0)
    assert out[0, 5] == pytest.approx(3.0 / 1.2) and out[0, 1] == pytest.approx(1.0 / 1.2) and out[0, 0] == 2.0
    # bigram (5, 1) was seen: after a trailing 5 the token 1 is banned with no_repeat_ngram_size=2
    out = G.process_logits(logits.clone(), seq, 1.0, 0, None, 1.0, 2)
    assert out[0, 1] == float("-inf") and torch.isfinite(out[0, [0, 2, 3, 4, 5]]).all()
    out = G.process_logits(logits.clone(), seq[:, :0], 1.0, 2, None, 1.0, 0)
    assert torch.isfinite(out[0]).sum() == 2 and torch.isfinite(out[0, [0, 5]]).all()
    out = G.process_logits(logits.clone(), seq[:, :0], 1.0, 0, 0.5, 1.0, 0)
    assert torch.isfinite(out[0, 5]) and torch.isfinite(out[0]).sum() <= 2           # 3.0 carries 0.57 of the mass: nucleus = {5} or {5, 0}
    out = G.process_logits(logits.clone(), seq[:, :0], 0.5, 0, None, 1.0, 0)
    torch.testing.assert_close(out, logits / 0.5)
