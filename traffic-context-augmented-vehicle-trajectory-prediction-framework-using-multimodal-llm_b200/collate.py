"""Batch assembly for the hot path: the reference's `custom_collate_fn` (scripts/train.py:301-347) with the
same output keys and values, but built for the device path:

  * every tensor of the batch lands in ONE pinned host arena and travels with one H2D copy (`ScenePack.to_device`);
  * `lane_polygon_len` / `norm_stat` — Python lists in the reference, which `forward()` re-tensorises and copies on every call
    (train.py:946-949) — are additionally packed as `int32[B]` / `float32[B, 4]` tensors (`lane_polygon_len_t`, `norm_stat_t`);
  * token sequences are right-padded exactly like `pad_sequence(..., padding_value=0 / 0 / -100)`.

The reference-facing dict keeps the list-typed entries, so `model(batch["traj_emb"], batch["vision_emb"], batch["context_str"],
batch["lane_polygon"], batch["lane_polygon_len"], y=batch["target_traj"], norm_stat=batch["norm_stat"], input_ids=...)` works
unchanged (train.py:1169-1176)."""
import torch

_STRS = ("context_str", "answer_str", "track_id")


def _arena(spec, pin):
    """spec: [(key, shape, dtype)] -> (flat uint8 buffer, {key: tensor view}); every view starts 16-byte aligned."""
    off, plan = 0, []
    for key, shape, dtype in spec:
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        plan.append((key, shape, dtype, off, nbytes))
        off += (nbytes + 15) // 16 * 16
    buf = torch.empty(max(off, 16), dtype=torch.uint8)
    # a DataLoader worker process must not touch CUDA (pin_memory() initialises it in the forked child): workers hand over pageable
    # memory and the loader's own pin_memory thread (DataLoader(pin_memory=True)) pins the arena in the parent
    if pin and torch.utils.data.get_worker_info() is None and torch.cuda.is_available():
        buf = buf.pin_memory()
    views = {key: buf[o:o + nb].view(dtype).view(*shape) for key, shape, dtype, o, nb in plan}
    return buf, views


class ScenePack(dict):
    """The collated batch (reference keys) plus its arena; `to_device` moves the numeric part with a single async copy.
    The arena and its layout travel as ordinary entries (`_arena`, `_plan`), so they survive `DataLoader(pin_memory=True)`, which
    rebuilds mappings entry by entry (and pins `_arena` like any other tensor)."""

    @property
    def arena(self):
        return self.get("_arena")

    @property
    def plan(self):
        return self.get("_plan")

    def to_device(self, device, non_blocking=True):
        return pack_to_device(self, device, non_blocking)


def pack_to_device(batch, device, non_blocking=True):
    """`ScenePack.to_device` for any mapping that carries `_arena` / `_plan` (e.g. the plain dict a pinning DataLoader hands back).
    After the loader's re-wrap the per-key tensors no longer alias the arena, but both hold the same bytes: the arena is what moves."""
    arena, plan = batch["_arena"], batch["_plan"]
    dbuf = arena.to(device, non_blocking=non_blocking)
    out = ScenePack({k: v for k, v in batch.items() if not torch.is_tensor(v)})
    for key, shape, dtype, o, nb in plan:
        out[key] = dbuf[o:o + nb].view(dtype).view(*shape)
    out["_arena"] = dbuf
    return out


def custom_collate_fn(batch, pin=True):
    """Drop-in for reference scripts/train.py:301-347 (same keys, same values, same dtypes)."""
    B = len(batch)
    t_in, t_out = batch[0]["traj_emb"].shape[0], batch[0]["target_traj"].shape[0]
    tv, dv = batch[0]["vision_emb"].shape
    P = batch[0]["lane_polygon"].shape[0]
    lt = max(int(b["input_ids"].shape[0]) for b in batch)
    spec = [("traj_emb", (B, 2, t_in), torch.float32), ("target_traj", (B, 2, t_out), torch.float32),
            ("vision_emb", (B, tv, dv), batch[0]["vision_emb"].dtype), ("lane_polygon", (B, P, 2), torch.float32),
            ("lane_polygon_len_t", (B,), torch.int32), ("norm_stat_t", (B, 4), torch.float32),
            ("input_ids", (B, lt), torch.int64), ("attention_mask", (B, lt), torch.int64), ("labels", (B, lt), torch.int64)]
    buf, v = _arena(spec, pin)
    v["input_ids"].zero_()
    v["attention_mask"].zero_()
    v["labels"].fill_(-100)
    for i, b in enumerate(batch):
        v["traj_emb"][i] = b["traj_emb"].transpose(0, 1)          # (T_in, 2) -> (2, T_in), train.py:309-310
        v["target_traj"][i] = b["target_traj"].transpose(0, 1)
        v["vision_emb"][i] = b["vision_emb"]
        v["lane_polygon"][i] = b["lane_polygon"]
        v["lane_polygon_len_t"][i] = int(b["lane_polygon_len"])
        v["norm_stat_t"][i] = torch.tensor([float(s) for s in b["norm_stat"]], dtype=torch.float32)
        n = int(b["input_ids"].shape[0])
        v["input_ids"][i, :n] = b["input_ids"]
        v["attention_mask"][i, :n] = b["attention_mask"]
        v["labels"][i, :n] = b["labels"]
    out = ScenePack(v)
    out["lane_polygon_len"] = [b["lane_polygon_len"] for b in batch]
    out["norm_stat"] = [b["norm_stat"] for b in batch]
    for k in _STRS:
        out[k] = [b[k] for b in batch]
    out["_arena"] = buf
    off, plan = 0, []
    for key, shape, dtype in spec:
        nb = v[key].numel() * v[key].element_size()
        plan.append((key, shape, dtype, off, nb))
        off += (nb + 15) // 16 * 16
    out["_plan"] = [tuple(p) for p in plan]
    return out
