"""Native fine-tune loop (reference scripts/im_kim_train_GRN.py:1019-1041: zero_grad -> forward -> loss.backward() ->
AdamW.step(), wrapped in DistributedDataParallel, train.py:1127-1132).

`FineTuner` keeps every trainable tensor in ONE flat fp32 parameter buffer (the nn.Parameters are re-pointed to views of
it), their gradients in one flat buffer (`distributed.FlatGradBucket`) and the AdamW moments in two more, so a step is:
forward + hand-written backward (train_engine.py, gradients written straight into their slices of the flat buffer) -> ONE
all-reduce of the trainable gradients over NCCL (LoRA / Q-Former / encoder / fusion only; frozen base weights never move)
-> ONE fused AdamW launch (tcavp_adamw).  Forward + backward of a fixed batch shape are captured once in a CUDA graph and
replayed (one host call instead of ~900 kernel launches per step).  The reference's own loop (torch.optim.AdamW + DDP
+ loss.backward()) also works unchanged on the same model — this class is the faster equivalent."""
import torch
import torch.distributed as dist

from . import ops
from .distributed import FlatGradBucket


class FineTuner:
    def __init__(self, model, lr=5e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, group=None, use_cuda_graph=True, max_graphs=8):
        self.model, self.group = model, group
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        named = model.trainable_named_parameters()
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
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
        self.targets = dict(zip(self.names, self.bucket.views))
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.steps = 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.use_cuda_graph = use_cuda_graph
        # one captured graph (+ its static input buffers and output tensors) per input-shape signature: the reference collate pads the
        # prompt to the longest sequence of each batch (train.py:323-325), so L_text moves from step to step and a single-slot cache would
        # re-capture almost every step.  Least-recently-used entries are dropped beyond `max_graphs` (each holds one step's activations).
        self._graphs = {}
        self.max_graphs = max_graphs
        self.captures = 0
        self.launches_per_step = 0    # libtcavp launches of one forward + backward (captured launches are replayed, not re-counted)
        # train() mode applies the reference's dropout at every site (counter-based masks; the step counter lives on the device and is
        # advanced inside the captured graph).  Ranks draw different masks, as DDP ranks with different RNG offsets do.
        self.dropout_active = model.dropout_active()
        if self.dropout_active:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
            model.set_dropout_seed((torch.initial_seed() + 7919 * rank) & 0x7FFFFFFF, 0)

    @property
    def payload_bytes(self):
        """Bytes exchanged by the one all-reduce of a step."""
        return self.bucket.flat.numel() * self.bucket.flat.element_size()

    # ---- forward + backward into the flat gradient buffer (no autograd graph, no per-tensor accumulation) --------------
    @torch.no_grad()
    def _fwd_bwd(self, inp):
        eng = self.model.train_engine()
        eng.grad_targets = self.targets
        self.bucket.flat.zero_()
        try:
            out = eng.train_forward(**inp)
            grads = eng.train_backward(None)
        finally:
            eng.grad_targets = None
        for name, view in self.targets.items():
            g = grads.get(name)
            if g is not None and g.data_ptr() != view.data_ptr():
                view.copy_(g.reshape(view.shape))
        return out["loss"], out["decoded"]

    def _device_inputs(self, x, vision_embs, lane_polygon_batch, lane_polygon_len, y, norm_stat, input_ids, attention_mask):
        dev = self.flat_p.device
        lens = lane_polygon_len if torch.is_tensor(lane_polygon_len) else torch.tensor(list(lane_polygon_len), dtype=torch.int32)
        ns = norm_stat if torch.is_tensor(norm_stat) else torch.tensor(norm_stat, dtype=torch.float32)
        f32 = lambda t: t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()      # noqa: E731
        vis = vision_embs.to(dev, non_blocking=True)
        if vis.dtype not in (torch.float32, torch.bfloat16):
            vis = vis.float()
        return dict(x=f32(x), vision=vis.contiguous(), polygon=f32(lane_polygon_batch), poly_len=lens.to(device=dev, dtype=torch.int32),
                    input_ids=input_ids.to(device=dev, dtype=torch.int64).contiguous(),
                    attention_mask=attention_mask.to(device=dev, dtype=torch.int64).contiguous(), y=f32(y), norm_stat=f32(ns).view(-1, 4))

    def _run(self, inp):
        if not self.use_cuda_graph:
            n0 = ops.launch_count()
            r = self._fwd_bwd(inp)
            self.launches_per_step = ops.launch_count() - n0
            return r
        sig = tuple((k, tuple(v.shape), v.dtype) for k, v in inp.items())
        ent = self._graphs.pop(sig, None)
        if ent is None:
            # static input buffers + two eager steps on a side stream (lazy packing, cudaFuncSetAttribute, allocator warm-up), then capture
            static = {k: v.clone() for k, v in inp.items()}
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._fwd_bwd(static)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                out = self._fwd_bwd(static)
            ent = (g, static, out, ops.launch_count() - n0)
            self.captures += 1
            while len(self._graphs) >= max(1, self.max_graphs):
                self._graphs.pop(next(iter(self._graphs)))        # dicts keep insertion order: the first key is the least recently used
        self._graphs[sig] = ent                                   # (re-)inserted last = most recently used
        g, static, out, self.launches_per_step = ent
        for k, v in inp.items():
            static[k].copy_(v, non_blocking=True)
        g.replay()
        # the graph's output tensors are overwritten by the next replay: hand out copies
        return out[0].clone(), out[1].clone()

    def all_reduce_only(self):
        """The step's collective on its own (bench.py times it outside the step to report what the exchange costs)."""
        if self.world > 1:
            dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=self.group)

    def step(self, x, vision_embs, context_str, lane_polygon_batch, lane_polygon_len, y, norm_stat, input_ids, attention_mask):
        """One optimisation step on this rank's shard of the batch; returns (loss, decoded) like the reference forward."""
        inp = self._device_inputs(x, vision_embs, lane_polygon_batch, lane_polygon_len, y, norm_stat, input_ids, attention_mask)
        loss, decoded = self._run(inp)
        if self.world > 1:
            dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.steps += 1
        ops.adamw_(self.flat_p, self.bucket.flat, self.exp_avg, self.exp_avg_sq, lr=self.lr, betas=self.betas, eps=self.eps,
                   weight_decay=self.wd, step=self.steps, grad_scale=1.0 / self.world)
        self.model._engine = None        # parameters changed behind torch's version counters: drop the inference packing
        return loss.detach(), decoded
