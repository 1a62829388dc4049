"""Text generation side path (reference scripts/train.py:577-654 `LlamaMultiModal.generate_batch`, called once per epoch at
train.py:1231-1241; the stage-1 script scripts/check_generation.py:131-222 drives the same call).

The reference monkey-patches the backbone's embedding layer and calls HF `generate(do_sample=True, temperature, top_k, top_p,
no_repeat_ngram_size=3, repetition_penalty=1.2)`.  Here the same sampler runs over the decoder kernels of libtcavp.so:
  * prefix = Q-Former image tokens (+ q_proj, + vision modality embedding) followed by the prompt embeddings (+ text modality
    embedding) — `Engine.prefix_embeds`, the assembly of train.py:585-598;
  * every step evaluates the decoder stack over the sequence so far (`Engine.llm_forward`: tcgen05 GEMMs, tcgen05 / flash attention)
    and the vocabulary logits of the LAST position with one M = B GEMM against lm_head;
  * logits processors in HF's order: repetition penalty -> no-repeat-ngram -> temperature -> top-k -> top-p -> multinomial draw.
Decoding uses a KV cache (Llama- and GPT-2-arch backbones): the prefix is evaluated once (`Engine.llm_forward(kv_out=...)` keeps every
layer's rotated keys / values), then every new token is one `Engine.llm_decode_step` — a one-row pass through the same kernels whose
query attends over the cached positions.  `kv_cache=False` recomputes the sequence so far at every step, which is also what the cached
path is tested against.  The sampled tokens depend on torch's RNG, the logits do not.

Deviation from the reference, on purpose: the reference's patched embedding returns `fused_embeds[:, :len(ids)]` on the first call, i.e.
it silently drops the last 16 prompt positions and hands HF an attention mask that is 16 longer than the input (train.py:604-611,
627-632).  This implementation generates after the FULL prefix (image tokens + whole prompt), which is what the surrounding code intends."""
import torch

from . import ops


@torch.no_grad()
def next_token_logits(engine, embeds, mask=None):
    """embeds (B, L, H) in the activation dtype -> fp32 logits (B, V) of the token that follows position L - 1."""
    B, L, H = embeds.shape
    if mask is None:
        mask = torch.ones(B, L, dtype=torch.int32, device=embeds.device)
    fh = engine.llm_forward(embeds.clone(), mask, B, L)                    # (B * L, H), post final norm (= hidden_states[-1])
    last = fh.view(B, L, H)[:, L - 1].contiguous()
    head = engine.lm_head()
    logits = torch.empty(B, head.shape[0], dtype=torch.float32, device=embeds.device)
    return ops.gemm(last, head, logits)


def process_logits(logits, seq, temperature, top_k, top_p, repetition_penalty, no_repeat_ngram_size):
    """HF's LogitsProcessorList of `generate(do_sample=True, ...)` for one step; `seq` (B, T) = token ids so far (prompt + generated)."""
    B, V = logits.shape
    if repetition_penalty and repetition_penalty != 1.0 and seq.shape[1] > 0:      # RepetitionPenaltyLogitsProcessor
        score = torch.gather(logits, 1, seq)
        score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
        logits = logits.scatter(1, seq, score)
    n = int(no_repeat_ngram_size or 0)
    if n > 0 and seq.shape[1] + 1 >= n:                                             # NoRepeatNGramLogitsProcessor
        for b in range(B):
            ids = seq[b].tolist()
            prefix = tuple(ids[len(ids) - (n - 1):]) if n > 1 else ()
            banned = [ids[i + n - 1] for i in range(len(ids) - n + 1) if tuple(ids[i:i + n - 1]) == prefix]
            if banned:
                logits[b, torch.tensor(banned, device=logits.device)] = float("-inf")
    if temperature and temperature != 1.0:                                          # TemperatureLogitsWarper
        logits = logits / temperature
    if top_k and 0 < top_k < V:                                                     # TopKLogitsWarper
        kth = torch.topk(logits, top_k, dim=-1).values[:, -1:]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if top_p is not None and 0.0 < top_p < 1.0:                                     # TopPLogitsWarper (keeps at least one token)
        srt, idx = torch.sort(logits, descending=False, dim=-1)
        cum = srt.softmax(dim=-1).cumsum(dim=-1)
        remove = cum <= (1.0 - top_p)
        remove[:, -1] = False
        logits = logits.masked_fill(torch.zeros_like(remove).scatter(1, idx, remove), float("-inf"))
    return logits


@torch.no_grad()
def generate_ids(model, vision_embs, prompt_ids, max_new_tokens=128, temperature=0.9, top_k=40, top_p=0.9, do_sample=True,
                 repetition_penalty=1.2, no_repeat_ngram_size=3, eos_token_id=None, pad_token_id=None, generator=None, kv_cache=True):
    """-> (B, prompt + new) token ids: the prompt followed by the generated tokens (rows that hit `eos_token_id` are padded)."""
    eng = model.engine()
    dev = eng.dev
    ids = prompt_ids.to(device=dev, dtype=torch.int64)
    B = ids.shape[0]
    prefix = eng.prefix_embeds(vision_embs, ids)                           # (B, Q + Lp, H)
    P, H = prefix.shape[1], prefix.shape[2]
    seq = ids.clone()
    new_embeds = []
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    pad = eos_token_id if pad_token_id is None else pad_token_id
    use_cache = bool(kv_cache) and int(max_new_tokens) > 0
    caches, head, logits = [], eng.lm_head(), None
    if use_cache:      # prefill: one pass over the prefix, keys / values of every layer kept
        ones = torch.ones(B, P, dtype=torch.int32, device=dev)
        fh = eng.llm_forward(prefix.clone(), ones, B, P, kv_out=(caches, P + int(max_new_tokens)))
        logits = ops.gemm(fh.view(B, P, H)[:, P - 1].contiguous(), head, torch.empty(B, head.shape[0], dtype=torch.float32, device=dev))
    for step in range(int(max_new_tokens)):
        if not use_cache:
            embeds = prefix if not new_embeds else torch.cat([prefix] + new_embeds, dim=1)
            logits = next_token_logits(eng, embeds)
        logits = process_logits(logits, seq, temperature if do_sample else 1.0, top_k if do_sample else 0, top_p if do_sample else None,
                                repetition_penalty, no_repeat_ngram_size)
        if do_sample:
            nxt = torch.multinomial(logits.softmax(dim=-1), 1, generator=generator)[:, 0]
        else:
            nxt = logits.argmax(dim=-1)
        if eos_token_id is not None:
            nxt = torch.where(done, torch.full_like(nxt, pad if pad is not None else 0), nxt)
            done = done | (nxt == eos_token_id)
        seq = torch.cat([seq, nxt[:, None]], dim=1)
        emb = eng.token_embeds(nxt[:, None])                               # plain embedding rows (no modality vector): train.py:609-611
        if eos_token_id is not None and bool(done.all()):
            break
        if use_cache:
            if step + 1 < int(max_new_tokens):
                h = eng.llm_decode_step(emb.view(B, H), caches, P + step)
                logits = ops.gemm(h, head, torch.empty(B, head.shape[0], dtype=torch.float32, device=dev))
        else:
            new_embeds.append(emb)
    return seq


def generate_batch(model, vision_embs, prompt_ids, tokenizer, max_new_tokens=128, temperature=0.9, top_k=40, top_p=0.9, device="cuda"):
    """Drop-in for `LlamaMultiModal.generate_batch` (train.py:577-654): list of decoded strings, cut after the reference's end marker."""
    was_training = model.training
    model.eval()
    try:
        out = generate_ids(model, vision_embs, prompt_ids, max_new_tokens=max_new_tokens, temperature=temperature, top_k=top_k, top_p=top_p,
                           eos_token_id=getattr(tokenizer, "eos_token_id", None), pad_token_id=getattr(tokenizer, "pad_token_id", None))
    finally:
        model.train(was_training)
    texts = []
    marker = "No right-following vehicle."                                 # train.py:648-652
    for row in out:
        text = tokenizer.decode(row, skip_special_tokens=True)
        if marker in text:
            text = text[: text.index(marker) + len(marker)]
        texts.append(text)
    return texts
