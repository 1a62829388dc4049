"""Tokenised-prompt cache: the reference re-tokenises an identical prompt for every sliding window
(scripts/train.py:200-238: the prompt differs between windows only through `track_id`, the answer text varies per window).

`PromptCache` memoises the prompt half by its exact text, so a track's windows share one tokenisation, and assembles
(input_ids, attention_mask, labels) exactly as train.py:211-238 does: prompt and answer tokenised separately with
add_special_tokens=False and truncation at max_length, concatenated, labels = -100 on the prompt positions, everything cut to max_length.
Keyed by the whole prompt string, not by pieces of it: BPE / sentencepiece merges can cross a split point, so only whole-string reuse is
guaranteed to give the ids the reference would get.  (A KV-cache of the prompt prefix is NOT valid for this model: the 16 scene-specific
image tokens precede the prompt in the fused sequence, train.py:528.)"""
import torch


class PromptCache:
    def __init__(self, tokenizer, max_length=512, max_entries=100000):
        self.tokenizer, self.max_length, self.max_entries = tokenizer, int(max_length), int(max_entries)
        self._prompts = {}
        self.hits = self.misses = 0

    def _tok(self, text):
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_tensors="pt", add_special_tokens=False)
        return enc["input_ids"], enc["attention_mask"]

    def prompt(self, prompt_text):
        """(input_ids (1, Lp), attention_mask (1, Lp)) of the prompt, tokenised once per distinct text."""
        hit = self._prompts.get(prompt_text)
        if hit is not None:
            self.hits += 1
            return hit
        self.misses += 1
        if len(self._prompts) >= self.max_entries:
            self._prompts.pop(next(iter(self._prompts)))
        ids, mask = self._tok(prompt_text)
        self._prompts[prompt_text] = (ids, mask)
        return ids, mask

    def encode(self, prompt_text, answer_text):
        """dict(input_ids, attention_mask, labels), each (L,) int64 — the per-sample tensors of train.py:211-238, 246-248."""
        p_ids, p_mask = self.prompt(prompt_text)
        a_ids, a_mask = self._tok(answer_text)
        input_ids = torch.cat([p_ids, a_ids], dim=1)
        attention_mask = torch.cat([p_mask, a_mask], dim=1)
        labels = torch.full_like(input_ids, -100)
        n = p_ids.size(1)
        labels[:, n:] = input_ids[:, n:]
        if input_ids.size(1) > self.max_length:
            input_ids, attention_mask, labels = input_ids[:, :self.max_length], attention_mask[:, :self.max_length], labels[:, :self.max_length]
        return {"input_ids": input_ids.squeeze(0), "attention_mask": attention_mask.squeeze(0), "labels": labels.squeeze(0)}
