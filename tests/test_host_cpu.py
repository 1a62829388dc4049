"""Host-side logic that needs no GPU: state-dict contract, checkpoint key translation, packing, C-ABI exports."""
import ctypes
import re
import os

import pytest
import torch

import tcavp_b200 as T
from conftest import ROOT, load_golden


def test_state_dict_layout_matches_reference():
    for name in ("tiny_b6", "cfg1_b8", "cfg5_b32"):
        fix = load_golden(name)
        sd = T.MultiModalTrajectoryModel(**fix["model_cfg"]).state_dict()
        assert set(sd) == set(fix["state_shapes"]), name
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(fix["state_shapes"][k]), (name, k)


def test_only_lora_and_non_llm_params_trainable():
    m = T.MultiModalTrajectoryModel(**T.MODEL_PRESETS["tiny"])
    for n, p in m.named_parameters():
        if "llama_model" in n:
            assert p.requires_grad == ("lora_" in n), n
        else:
            assert p.requires_grad, n


def test_checkpoint_key_translation():
    cfg = dict(T.MODEL_PRESETS["tiny"])
    m = T.MultiModalTrajectoryModel(**cfg)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    # (a) V2 layout (im_kim_train_GRN.py) + (b) peft<0.7 layout without .base_layer
    old = {}
    for k, v in sd.items():
        k2 = k.replace("mllm.llama_wrapper.llama_model.", "mllm.llama_model.").replace(".base_layer.", ".")
        old[k2] = v.clone() + 1.0
    m.load_state_dict(old, strict=True)
    for k, v in m.state_dict().items():
        torch.testing.assert_close(v, sd[k] + 1.0, msg=k)
    # (c) LoRA checkpoint into a no-LoRA model = reference adjust_state_dict (ablation_study_without_lora.py:1071-1079)
    cfg["use_lora"] = False
    m2 = T.MultiModalTrajectoryModel(**cfg)
    m2.load_state_dict(m.state_dict(), strict=True)
    k = "mllm.llama_wrapper.llama_model.model.layers.0.self_attn.q_proj.weight"
    torch.testing.assert_close(m2.state_dict()[k], m.state_dict()[k.replace("llama_model.model", "llama_model.base_model.model.model").replace("q_proj.weight", "q_proj.base_layer.weight")])


def test_engine_packing_shapes():
    from tcavp_b200.engine import Engine
    m = T.MultiModalTrajectoryModel(**T.MODEL_PRESETS["tiny"])
    e = Engine(m, "bf16")
    c = T.resolve_llama("llama-tiny")
    nqkv = (c["num_attention_heads"] + 2 * c["num_key_value_heads"]) * c["head_dim"]
    assert e.llm["kx"] == 8 and e.llm["layers"][0]["wqkv"].shape == (nqkv, c["hidden_size"] + 8)
    ly = m.mllm.llama_wrapper.causal_lm().model.layers[0]
    # K-extension: q rows carry (alpha/r) B_q in columns [H, H+r), v rows carry (alpha/r) B_v in [H+r, H+2r), k rows zero
    H, r = c["hidden_size"], 4
    nq = c["num_attention_heads"] * c["head_dim"]
    nk = c["num_key_value_heads"] * c["head_dim"]
    w = e.llm["layers"][0]["wqkv"].float()
    assert torch.count_nonzero(w[nq:nq + nk, H:]) == 0
    assert torch.count_nonzero(w[:nq, H + r:]) == 0 and torch.count_nonzero(w[nq + nk:, H:H + r]) == 0
    # gate/up interleave
    gu = e.llm["layers"][0]["wgu"].float()
    torch.testing.assert_close(gu[0::2], ly.mlp.gate_proj.weight.detach().bfloat16().float())
    torch.testing.assert_close(gu[1::2], ly.mlp.up_proj.weight.detach().bfloat16().float())
    # (c,t) -> (t,c) permutation of the lane_fc rows
    C, To = 64, 12
    lf = m.ltsf.decoder.lane_fc.weight.detach()
    torch.testing.assert_close(e.lt["lane_fc"].w.view(To, C, -1)[3, 5], lf.view(C, To, -1)[5, 3])   # small path stays fp32


def test_forward_without_cuda_fails_loudly():
    m = T.MultiModalTrajectoryModel(**T.MODEL_PRESETS["tiny"]).eval()
    s = T.make_scenes(2, 6, 12, vision_dim=32, l_text=8, vocab=97)
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with torch.no_grad(), pytest.raises(Exception, match="CUDA|no CPU fallback"):
        m(s["x"], s["vision"], s["context_str"], s["polygon"], s["poly_len"], input_ids=s["input_ids"], attention_mask=s["attention_mask"])


def test_synthetic_scene_shapes_and_normalisation():
    s = T.make_scenes(5, 15, 25)
    assert s["x"].shape == (5, 2, 15) and s["y"].shape == (5, 2, 25) and s["vision"].shape == (5, 15, 512)
    full = torch.cat([s["x"], s["y"]], dim=2)
    assert float(full.min()) == 0.0 and float(full.max()) == 1.0           # joint min-max normalisation (train.py:164-184)
    assert all(l in T.synthetic.LANE_SIZES for l in s["poly_len"])
    assert (s["attention_mask"].sum(1) >= 96).all()


def test_c_abi_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "tcavp.h")).read()
    declared = set(re.findall(r"\b(tcavp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib_built, name), f"libtcavp.so does not export {name}"
    assert lib_built.tcavp_version() == 100
    import tcavp_b200.lib as L
    assert declared == set(L.EXPORTS)


def test_gemm_tile_order_is_a_panel_major_bijection(lib_built):
    """csrc/gemm.cu: unit_to_tile through its host entry point (no device work): every (row block, n-tile) is visited exactly once, panels
    of `pw` n-tiles are swept one after the other (all row blocks inside a panel before the next panel), n fastest inside a row block."""
    import ctypes
    import itertools
    for tm, tn, pw in itertools.chain([(1, 1, 1), (72, 86, 18), (72, 86, 29), (7, 5, 2), (3, 16, 16), (3, 16, 100), (41, 17, 3), (5, 9, 4)],
                                      ((tm, tn, pw) for tm in (1, 2, 6) for tn in (1, 2, 7, 12) for pw in (1, 2, 5, 12))):
        n = tm * tn
        mg, nt = (ctypes.c_int * n)(), (ctypes.c_int * n)()
        assert lib_built.tcavp_gemm_tile_order(tm, tn, pw, mg, nt) == 0
        tiles = list(zip(mg, nt))
        assert sorted(tiles) == [(a, b) for a in range(tm) for b in range(tn)], (tm, tn, pw)
        if pw >= tn:
            assert tiles == [(a, b) for a in range(tm) for b in range(tn)]
            continue
        panels = [b // pw for _, b in tiles]
        assert panels == sorted(panels), "panels are swept in order"
        for p in set(panels):
            inside = [t for t in tiles if t[1] // pw == p]
            assert inside == sorted(inside), "row-major (n fastest) inside a panel"
    assert lib_built.tcavp_gemm_tile_order(0, 4, 1, None, None) != 0


def test_collate_matches_the_reference_collate_fn():
    """tcavp_b200.custom_collate_fn against the output of the UNMODIFIED reference custom_collate_fn (scripts/train.py:301-347) on
    the same samples (golden minted by oracle/make_collate_golden.py): same keys, values, dtypes; plus the packed extras."""
    import torch
    import tcavp_b200 as T
    g = torch.load(os.path.join(ROOT, "tests", "golden", "collate_b5.pt"), weights_only=False)
    got = T.custom_collate_fn(g["samples"], pin=False)
    want = g["collated"]
    assert set(want) <= set(got)
    for k, v in want.items():
        if torch.is_tensor(v):
            assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
            assert torch.equal(got[k], v), k
        else:
            assert got[k] == v, k
    assert got["lane_polygon_len_t"].tolist() == want["lane_polygon_len"] and got["lane_polygon_len_t"].dtype == torch.int32
    assert torch.allclose(got["norm_stat_t"], torch.tensor(want["norm_stat"], dtype=torch.float32))
    # one arena: every tensor is a 16-byte aligned view of the same buffer, and a (CPU) round trip through to_device keeps the values
    base = got.arena.data_ptr()
    for key, shape, dtype, off, nb in got.plan:
        assert got[key].data_ptr() == base + off and off % 16 == 0
    moved = got.to_device("cpu")
    for k, v in want.items():
        if torch.is_tensor(v):
            assert torch.equal(moved[k], v), k
    # ragged / single-sample batches
    one = T.custom_collate_fn(g["samples"][3:4], pin=False)
    assert one["input_ids"].shape == (1, 12) and one["labels"].min() >= 0
