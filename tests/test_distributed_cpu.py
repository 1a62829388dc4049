"""world_size-2 gloo tests of the N>1 host logic (scene sharding, the final metric all-reduce, the flat trainable-gradient
all-reduce).  CPU only: the data path itself needs no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import tcavp_b200 as T
from tcavp_b200 import distributed as D


def test_scene_shard_partitions_the_range():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [D.scene_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b - a >= d - c >= b - a - 1
    with pytest.raises(ValueError):
        D.scene_shard(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stub_errors(x, y):
    """Stand-in for the network + metric block: "decoded" = last observed position repeated; per-scene (ADE, FDE) in float64."""
    dec = x[:, :, -1:].expand_as(y)
    d = (dec.double() - y.double()).pow(2).sum(1).sqrt()
    return d.mean(1), d[:, -1], dec


class _StubModel(torch.nn.Module):
    """Same predict_with_metrics contract as MultiModalTrajectoryModel (dict with `decoded` and `metrics[2:4]` = sum ADE / sum FDE)."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.calls = 0

    def predict_with_metrics(self, x, vision, polygon, lens, y, ns, ids, mask, max_poly_len=None):
        self.calls += 1
        ade, fde, dec = _stub_errors(x, y)
        m = torch.zeros(8, dtype=torch.float32)
        m[2], m[3] = ade.sum(), fde.sum()
        return {"decoded": dec.contiguous().float(), "metrics": m}


def _stub_batches(s, lo, hi, bs):
    for a in range(lo, hi, bs):
        b = min(a + bs, hi)
        yield {"traj_emb": s["x"][a:b], "target_traj": s["y"][a:b], "vision_emb": s["vision"][a:b], "lane_polygon": s["polygon"][a:b],
               "lane_polygon_len": s["poly_len"][a:b], "norm_stat": s["norm_stat"][a:b], "input_ids": s["input_ids"][a:b],
               "attention_mask": s["attention_mask"][a:b], "first": a}


def test_evaluate_accumulates_and_delivers_decoded_in_order():
    s = T.make_scenes(7, 6, 12, vision_dim=32, l_text=8, vocab=97, seed=3)
    model, seen = _StubModel().train(), []
    res = T.evaluate(model, _stub_batches(s, 0, 7, 3), on_decoded=lambda b, d: seen.append((b["first"], d.clone())))
    ade, fde, dec = _stub_errors(s["x"], s["y"])
    assert model.training and model.calls == 3                       # mode restored; one pass per batch
    assert res["n"] == 7 and abs(res["sum_ade"] - float(ade.sum())) < 1e-3 and abs(res["fde"] - float(fde.mean())) < 1e-4
    assert [f for f, _ in seen] == [0, 3, 6]
    assert torch.equal(torch.cat([d for _, d in seen]), dec.float())
    empty = T.evaluate(model, iter(()))
    assert empty["n"] == 0 and empty["ade"] != empty["ade"]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scenes = T.make_scenes(7, 6, 12, vision_dim=32, l_text=8, vocab=97, seed=3)
        mine = D.shard_scenes(scenes)
        lo, hi = D.scene_shard(7)
        assert mine["x"].shape[0] == hi - lo and mine["poly_len"] == scenes["poly_len"][lo:hi]
        # per-scene "ADE/FDE" stand-ins that are pure functions of the scene, so every sharding gives the same totals
        ade = scenes["y"].abs().sum(dim=(1, 2))
        fde = scenes["y"][:, :, -1].abs().sum(dim=1)
        m_ade, m_fde, n = D.reduce_metrics(ade[lo:hi].sum(), fde[lo:hi].sum(), hi - lo)
        assert n == 7
        assert abs(m_ade - float(ade.mean())) < 1e-5 and abs(m_fde - float(fde.mean())) < 1e-5
        # flat trainable-gradient bucket: frozen params stay out, mean over ranks comes back in every .grad view
        model = T.MultiModalTrajectoryModel(**T.MODEL_PRESETS["tiny"])
        bucket = D.FlatGradBucket(model.parameters())
        n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
        assert bucket.numel == n_train and all("lora_" in k or "llama_model" not in k
                                               for k, p in model.named_parameters() if p.requires_grad)
        for i, p in enumerate(bucket.params):
            p.grad.fill_(float(rank + 1) * (i % 5))
        bucket.all_reduce_mean()
        for i, p in enumerate(bucket.params):
            assert torch.all(p.grad == (1 + world) / 2.0 * (i % 5))
            assert p.grad.data_ptr() >= bucket.flat.data_ptr()
        # the scene-parallel test loop: every rank evaluates its own batches, one all-reduce at the end
        res = T.evaluate(_StubModel(), _stub_batches(scenes, lo, hi, 2))
        want = _stub_errors(scenes["x"], scenes["y"])
        assert res["n"] == 7 and abs(res["ade"] - float(want[0].mean())) < 1e-5 and abs(res["fde"] - float(want[1].mean())) < 1e-5    # fp32 batch sums
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
