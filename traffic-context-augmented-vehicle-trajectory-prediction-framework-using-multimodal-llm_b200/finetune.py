"""Native fine-tune loop (reference scripts/im_kim_train_GRN.py:1019-1041: zero_grad -> forward -> loss.backward() ->
AdamW.step(), wrapped in DistributedDataParallel, train.py:1127-1132).

`FineTuner` keeps every trainable tensor in ONE flat fp32 parameter buffer (the nn.Parameters are re-pointed to views of
it), their gradients in one flat buffer (`distributed.FlatGradBucket`) and the AdamW moments in two more, so a step is:
forward + hand-written backward (train_engine.py) -> ONE all-reduce of the trainable gradients over NCCL (LoRA / Q-Former /
encoder / fusion only; frozen base weights never move) -> ONE fused AdamW launch (tcavp_adamw).  The reference's own loop
(torch.optim.AdamW + DDP) also works unchanged on the same model — this class is the faster equivalent."""
import torch
import torch.distributed as dist

from . import ops
from .distributed import FlatGradBucket


class FineTuner:
    def __init__(self, model, lr=5e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, group=None):
        self.model, self.group = model, group
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.params = [p for _, p in model.trainable_named_parameters()]
        if not self.params:
            raise ValueError("no trainable parameters")
        if any(p.dtype != torch.float32 for p in self.params):
            raise TypeError("trainable parameters must be fp32 (master weights)")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                v = self.flat_p[off:off + p.numel()].view_as(p)
                v.copy_(p)
                p.data = v
                off += p.numel()
        self.bucket = FlatGradBucket(self.params)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.steps = 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    @property
    def payload_bytes(self):
        """Bytes exchanged by the one all-reduce of a step."""
        return self.bucket.flat.numel() * self.bucket.flat.element_size()

    def step(self, x, vision_embs, context_str, lane_polygon_batch, lane_polygon_len, y, norm_stat, input_ids, attention_mask):
        """One optimisation step on this rank's shard of the batch; returns (loss, decoded) like the reference forward."""
        self.bucket.zero_()
        with torch.enable_grad():
            loss, decoded = self.model(x, vision_embs, context_str, lane_polygon_batch, lane_polygon_len, y=y, norm_stat=norm_stat,
                                       input_ids=input_ids, attention_mask=attention_mask)
            loss.backward()
        if self.world > 1:
            dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.steps += 1
        ops.adamw_(self.flat_p, self.bucket.flat, self.exp_avg, self.exp_avg_sq, lr=self.lr, betas=self.betas, eps=self.eps,
                   weight_decay=self.wd, step=self.steps, grad_scale=1.0 / self.world)
        self.model._engine = None        # parameters changed behind torch's version counters: drop the inference packing
        return loss.detach(), decoded
