"""Scene-parallel evaluation and data-parallel gradient exchange.

Scenes are independent batch rows, so inference shards the scene index range contiguously over ranks with no
exchange during forward and ends with ONE all-reduce of (sum ADE, sum FDE, n) — the fix for the reference's
rank-0-only evaluation (reference scripts/train.py:1255-1273).  Fine-tuning is data parallel: one all-reduce per
step over a single flat buffer holding only the trainable gradients (the reference gets the same effect from
DistributedDataParallel's buckets, train.py:1127-1132).  One process per GPU; the backend is whatever the caller
initialised (`nccl` on the B200 box, `gloo` in the CPU tests)."""
import torch
import torch.distributed as dist


def scene_shard(n_scenes, rank=None, world=None):
    """Contiguous [lo, hi) slice of the scene range owned by `rank` (sizes differ by at most one)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_scenes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_scenes(scenes, rank=None, world=None):
    """Slices every per-scene entry (tensor rows / list items) of a make_scenes()-style dict."""
    n = scenes["x"].shape[0]
    lo, hi = scene_shard(n, rank, world)
    return {k: v[lo:hi] for k, v in scenes.items()}


def reduce_metrics(sum_ade, sum_fde, n, device=None, group=None):
    """All-reduces (sum ADE, sum FDE, n scenes) and returns (mean ADE, mean FDE, total n) as Python floats/ints
    (the reference's final print, train.py:1323-1326).  A rank that owns zero scenes contributes zeros."""
    t = torch.zeros(3, dtype=torch.float64, device=device)
    t[0], t[1], t[2] = float(sum_ade), float(sum_fde), float(n)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    tot = int(round(float(t[2])))
    if tot == 0:
        return float("nan"), float("nan"), 0
    return float(t[0]) / tot, float(t[1]) / tot, tot


class FlatGradBucket:
    """One flat buffer that every trainable parameter's .grad is a view of; `all_reduce_mean()` is the single collective of
    a data-parallel step.  Frozen parameters (the LLM base weights) are not in it, so the exchanged payload is the
    LoRA / Q-Former / encoder / fusion gradients only."""

    def __init__(self, params, dtype=torch.float32):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=dtype, device=dev)
        off = 0
        self.views = []
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            self.views.append(v)
            p.grad = v

    def zero_(self):
        self.flat.zero_()
        for p, v in zip(self.params, self.views):   # optimizers with set_to_none=True drop the views
            p.grad = v

    def all_reduce_mean(self, group=None):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.flat
