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
