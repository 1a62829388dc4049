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


def test_mllm_level_load_and_direct_variant():
    """reference scripts/train.py:1137-1138 restores a stage-1 checkpoint with `model.mllm.load_state_dict(sd, strict=True)`: the key
    translation must also run there.  llm_variant="direct" writes the V2 key layout (im_kim_train_GRN.py:444-455) on save."""
    cfg = dict(T.MODEL_PRESETS["tiny"])
    m = T.MultiModalTrajectoryModel(**cfg)
    sd = m.mllm.state_dict()
    v2 = {k.replace("llama_wrapper.llama_model.", "llama_model.").replace(".base_layer.", "."): v.clone() + 2.0 for k, v in sd.items()}
    assert any(k.startswith("llama_model.") for k in v2)
    ref = {k: v.clone() for k, v in sd.items()}
    m.mllm.load_state_dict(v2, strict=True)
    for k, v in m.mllm.state_dict().items():
        torch.testing.assert_close(v, ref[k] + 2.0, msg=k)
    d = T.MultiModalTrajectoryModel(**cfg, llm_variant="direct")
    keys = set(d.state_dict())
    assert not any("llama_wrapper" in k for k in keys) and any(k.startswith("mllm.llama_model.base_model.model.") for k in keys)
    assert {k.replace("mllm.llama_model.", "mllm.llama_wrapper.llama_model.") for k in keys} == set(m.state_dict())
    d.load_state_dict(d.state_dict(), strict=True)            # its own layout round-trips
    m.load_state_dict(d.state_dict(), strict=True)            # and loads into the train.py layout
    with pytest.raises(ValueError):
        T.MultiModalTrajectoryModel(**cfg, llm_variant="nope")
    # a plain (no-peft) backbone checkpoint into a LoRA model: keys gain base_model.model. / .base_layer (adapters stay missing)
    cfg2 = dict(cfg, use_lora=False)
    plain = T.MultiModalTrajectoryModel(**cfg2).state_dict()
    res = m.load_state_dict(plain, strict=False)
    assert not res.unexpected_keys and all("lora_" in k for k in res.missing_keys)


def test_reference_backbone_name_resolves_with_llama3_rope_and_tied_embeddings():
    """The one name the reference hard-codes (scripts/train.py:1347) builds the real geometry: GQA 32/8, llama3 rope scaling, tied
    lm_head / embed_tokens, vocabulary 128256 — constructed on the meta device (1.24 G parameters), with the reference's args dict."""
    from tcavp_b200.config import rope_inv_freq
    args = dict(seq_len=6, out_len=30, individual=True, feature_size=2, d_model=64, lane_polygon_d_model=64, lane_polygon_nhead=4,
                lane_polygon_layers=2, max_polygon_points=64, use_post_mlp=True, post_mlp_hidden_dim=64,
                base_model_name="meta-llama/Llama-3.2-1B", use_lora=True, lora_r=8, lora_alpha=32, lora_dropout=0.1, vision_dim=512,
                q_hidden_size=768, q_nhead=8, q_enc_layers=4, q_dec_layers=4, q_num_query_tokens=16, ltsf_nhead=2, ltsf_dropout=0.1)
    m = T.MultiModalTrajectoryModel(**args, llm_device="meta", llm_param_dtype=torch.bfloat16)
    c = m.mllm.llama_wrapper.config
    assert (c["hidden_size"], c["num_hidden_layers"], c["num_attention_heads"], c["num_key_value_heads"], c["vocab_size"]) == (2048, 16, 32, 8, 128256)
    assert m.llama_hidden_size == 2048 and isinstance(m.mllm.q_proj, torch.nn.Linear)
    lm = m.mllm.llama_wrapper.causal_lm()
    assert lm.lm_head.weight is lm.model.embed_tokens.weight
    sd = m.state_dict()
    pre = "mllm.llama_wrapper.llama_model.base_model.model."
    assert sd[pre + "lm_head.weight"].shape == sd[pre + "model.embed_tokens.weight"].shape == (128256, 2048)
    assert pre + "model.layers.15.self_attn.v_proj.lora_B.default.weight" in sd
    n_llm = sum(p.numel() for n, p in m.named_parameters() if "llama_model" in n and "lora_" not in n)
    assert n_llm == 1235814400                                                # the published parameter count of Llama-3.2-1B
    inv = rope_inv_freq(c)
    plain = 1.0 / (500000.0 ** (torch.arange(0, 64, 2).float() / 64))
    assert torch.equal(inv[:8], plain[:8])                                    # short wavelengths untouched
    torch.testing.assert_close(inv[-1], plain[-1] / 32.0)                     # long wavelengths slowed by `factor`
    assert (inv <= plain).all() and (inv >= plain / 32.0 - 1e-12).all()
    # a tied-embedding checkpoint without lm_head.weight (HF safetensors layout) still loads strictly
    small = T.MultiModalTrajectoryModel(**dict(T.MODEL_PRESETS["tiny"], base_model_name=dict(T.LLAMA_PRESETS["llama-tiny"], tie_word_embeddings=True)))
    ck = {k: v.clone() for k, v in small.state_dict().items() if not k.endswith("lm_head.weight")}
    small.load_state_dict(ck, strict=True)
    with pytest.raises(KeyError):
        T.resolve_llama("no-such/backbone")


def test_collate_survives_dataloader_workers_and_rewrapping():
    """custom_collate_fn inside DataLoader worker processes (no pinning there: CUDA must not initialise in a forked child) and
    through the mapping re-wrap a pinning loader performs: arena + plan are ordinary entries, pack_to_device still works."""
    g = torch.load(os.path.join(ROOT, "tests", "golden", "collate_b5.pt"), weights_only=False)
    dl = torch.utils.data.DataLoader(g["samples"], batch_size=3, shuffle=False, num_workers=2, collate_fn=T.custom_collate_fn)
    batches = list(dl)
    assert [b["traj_emb"].shape[0] for b in batches] == [3, 2]
    want = T.custom_collate_fn(g["samples"][:3], pin=False)
    rewrapped = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batches[0].items()}     # what pin_memory=True hands back
    moved = T.pack_to_device(rewrapped, "cpu")
    for k, v in want.items():
        if torch.is_tensor(v) and not k.startswith("_"):
            assert torch.equal(moved[k], v) and torch.equal(batches[0][k], v), k
    assert moved["lane_polygon_len"] == want["lane_polygon_len"]


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


def test_gpt2_train_engine_packing_and_lora_refresh():
    """GPT-2-arch backbone under the fine-tune engine (host side only): c_attn packed as [W^T | (alpha / r) B] with the LoRA pair re-packed
    from the fp32 masters on every sync, transposed copy for the dX GEMMs, trainable names in HF GPT2LMHeadModel + peft layout."""
    from tcavp_b200.train_engine import DropPlan, TrainEngine
    mc = dict(T.MODEL_PRESETS["tiny"], base_model_name="gpt2-tiny")
    m = T.MultiModalTrajectoryModel(**mc, compute_dtype="fp32")
    te = TrainEngine(m, "fp32")
    te.drop = DropPlan(None, {})
    c = T.resolve_llama("gpt2-tiny")
    H, r = c["hidden_size"], 4
    ca = m.mllm.llama_wrapper.causal_lm().transformer.h[1].attn.c_attn
    with torch.no_grad():
        ca.lora_B["default"].weight.normal_(0, 0.1)
    te.sync_params()
    ly = te.llm["layers"][1]
    assert te.llm["kx"] == 8 and ly["wqkv"].shape == (3 * H, H + 8) and ly["wqkvT"].shape == (H + 8, 3 * H)
    torch.testing.assert_close(ly["wqkv"][:, :H], ca.base_layer.weight.detach().t())                      # Conv1D [in, out] -> [N, K]
    torch.testing.assert_close(ly["wqkv"][:, H:H + r], ca.lora_B["default"].weight.detach() * ca.scaling)
    assert torch.count_nonzero(ly["wqkv"][:, H + r:]) == 0
    torch.testing.assert_close(ly["wqkvT"], ly["wqkv"].t())
    torch.testing.assert_close(ly["a_cat"][:r], ca.lora_A["default"].weight.detach())
    torch.testing.assert_close(ly["a_catT"], ly["a_cat"].t())
    with torch.no_grad():                                            # an optimizer step moves the masters: the next sync re-packs them
        ca.lora_A["default"].weight.add_(1.0)
    te.sync_params()
    torch.testing.assert_close(te.llm["layers"][1]["a_cat"][:r], ca.lora_A["default"].weight.detach())
    names = [n for n, _ in te.params if "lora_" in n]
    assert len(names) == 2 * c["num_hidden_layers"] and all(n.startswith(te.llm_prefix) and ".attn.c_attn.lora_" in n for n in names)
    assert te._dropout_probs()[("llm", "lora_c")] == pytest.approx(mc.get("lora_dropout", 0.1))


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


def test_prompt_cache_reproduces_the_reference_tokenisation_recipe():
    """PromptCache.encode against the recipe of reference scripts/train.py:211-238 on a toy tokenizer: same ids / mask / labels,
    truncation at max_length, and the prompt is tokenised once per distinct text."""
    import torch

    class Tok:                                  # whitespace tokenizer with the HF call signature the reference uses
        calls = 0

        def __call__(self, text, truncation=True, max_length=512, return_tensors="pt", add_special_tokens=False):
            assert truncation and return_tensors == "pt" and add_special_tokens is False
            Tok.calls += 1
            ids = [(sum(map(ord, w)) % 997) + 1 for w in text.split()][:max_length]
            return {"input_ids": torch.tensor([ids], dtype=torch.long), "attention_mask": torch.ones(1, len(ids), dtype=torch.long)}

    tok = Tok()

    def reference_recipe(prompt, answer, max_length):
        p, a = tok(prompt, max_length=max_length), tok(answer, max_length=max_length)
        ids = torch.cat([p["input_ids"], a["input_ids"]], dim=1)
        mask = torch.cat([p["attention_mask"], a["attention_mask"]], dim=1)
        labels = torch.full_like(ids, -100)
        n = p["input_ids"].size(1)
        labels[:, n:] = ids[:, n:]
        return ids[0, :max_length], mask[0, :max_length], labels[0, :max_length]

    prompts = [f"You are analyzing track_id={t} . Describe lane site speed neighbours Answer:" for t in (3, 3, 3, 7, 7)]
    answers = [f"vehicle in lane {i} at speed {10 + i} with {i} neighbours " * (1 + i) for i in range(5)]
    for max_length in (512, 12):
        cache = T.PromptCache(tok, max_length=max_length)
        Tok.calls = 0
        got = [cache.encode(p, a) for p, a in zip(prompts, answers)]
        assert Tok.calls == 5 + 2 and cache.misses == 2 and cache.hits == 3          # 5 answers + 2 distinct prompts
        for g, p, a in zip(got, prompts, answers):
            ids, mask, labels = reference_recipe(p, a, max_length)
            assert torch.equal(g["input_ids"], ids) and torch.equal(g["attention_mask"], mask) and torch.equal(g["labels"], labels)
            assert g["input_ids"].shape[0] <= max_length


def test_trainable_and_lora_only_checkpoints_round_trip():
    cfg = dict(T.MODEL_PRESETS["tiny"])
    a, b = T.MultiModalTrajectoryModel(**cfg), T.MultiModalTrajectoryModel(**cfg)
    T.deterministic_fill_(a.state_dict(), 5)
    T.deterministic_fill_(b.state_dict(), 6)
    tsd = a.trainable_state_dict()
    assert set(tsd) == {n for n, p in a.named_parameters() if p.requires_grad}
    assert not any("llama_model" in k and "lora_" not in k for k in tsd)
    assert sum(v.numel() for v in tsd.values()) < 0.9 * sum(v.numel() for v in a.state_dict().values())
    b.load_trainable_state_dict(tsd)
    for k, v in b.state_dict().items():
        assert torch.equal(v, a.state_dict()[k]) == (k in tsd), k           # trainables copied, frozen backbone untouched
    with pytest.raises(RuntimeError):
        b.load_trainable_state_dict({k: v for k, v in list(tsd.items())[1:]})
    lsd = a.lora_state_dict()
    assert all(k.startswith("base_model.model.model.layers.") and ".default." not in k for k in lsd) and len(lsd) == 2 * 2 * 2
    c = T.MultiModalTrajectoryModel(**cfg)
    c.load_lora_state_dict(lsd)
    for k, v in a.lora_state_dict(peft_format=False).items():
        assert torch.equal(c.state_dict()[k], v)
    with pytest.raises(RuntimeError):
        c.load_lora_state_dict({k: v for k, v in list(lsd.items())[:-1]})
