"""The CPU restatement (oracle/restated.py) against the golden vectors minted from the unmodified reference."""
import pytest
import torch

import tcavp_b200 as T
from conftest import load_golden
from oracle import restated


def _run(fix):
    m = T.MultiModalTrajectoryModel(**fix["model_cfg"])   # parameter container only (CPU); gives the key layout
    sd = m.state_dict()
    assert set(sd) == set(fix["state_shapes"])
    T.deterministic_fill_(sd, fix["weight_seed"])
    for k, v in fix["weight_checksums"].items():
        assert float(sd[k].double().sum()) == pytest.approx(v, rel=1e-12, abs=1e-12), f"weight filler drifted on {k}"
    i = fix["inputs"]
    return restated.forward(sd, fix["model_cfg"], fix["llama_cfg"], i["x"], i["vision"], i["polygon"], i["poly_len"],
                            i["input_ids"], i["attention_mask"], i["y"], i["norm_stat"])


@pytest.mark.parametrize("name", ["tiny_b6", "cfg1_b8", "cfg5_b32", "cfg3l2_b16", "gqa_l2_b32", "llama32_1b_l2_b8", "gpt2_tiny_b6", "gpt2_l2_b8"])
def test_restatement_matches_reference_golden(name):
    fix = load_golden(name)
    o, g = _run(fix), fix["out"]
    # fp32 re-association noise only: the restatement uses the same arithmetic in a different op order
    torch.testing.assert_close(o["poly_emb"], g["poly_emb"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(o["enc"], g["enc"], rtol=1e-4, atol=1e-5)
    n = g["final_hidden_head"].shape[0]
    torch.testing.assert_close(o["final_hidden"][:n], g["final_hidden_head"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(o["final_hidden"].mean(-1), g["final_hidden_rowmean"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(o["decoded"], g["decoded"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(o["loss"], g["loss"], rtol=1e-5, atol=0)
    torch.testing.assert_close(o["ade"], g["ade"], rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(o["fde"], g["fde"], rtol=1e-5, atol=1e-3)


def test_zero_length_polygon_gives_zero_embedding():
    fix = load_golden("tiny_b6")
    assert fix["inputs"]["poly_len"][0] == 0
    o = _run(fix)
    assert torch.count_nonzero(o["poly_emb"][0]) == 0
    assert torch.isfinite(o["poly_emb"]).all()


def test_peft_shim_layout_and_math():
    from oracle import peft_shim
    base = torch.nn.Linear(16, 12, bias=False)
    l = peft_shim.LoraLinear(base, r=4, alpha=16, dropout=0.0)
    torch.nn.init.normal_(l.lora_B["default"].weight)
    x = torch.randn(3, 16)
    want = x @ base.weight.t() + (x @ l.lora_A["default"].weight.t()) @ l.lora_B["default"].weight.t() * 4.0
    torch.testing.assert_close(l(x), want)
    assert set(l.state_dict()) == {"base_layer.weight", "lora_A.default.weight", "lora_B.default.weight"}


def _check_against_compressed(got, want, rtol, atol, key):
    if "full" in want:
        torch.testing.assert_close(got.float().cpu(), want["full"], rtol=rtol, atol=atol, msg=lambda m: f"{key}: {m}")
        return
    g2 = got.float().cpu().reshape(got.shape[0], -1)
    assert tuple(got.shape) == tuple(want["shape"]), key
    torch.testing.assert_close(g2[:4], want["head"], rtol=rtol, atol=atol, msg=lambda m: f"{key} head: {m}")
    scale = float(want["rowsum"].abs().max()) + float(want["colsum"].abs().max()) + 1e-12
    torch.testing.assert_close(g2.double().sum(1).float(), want["rowsum"], rtol=rtol, atol=atol + 1e-4 * scale, msg=lambda m: f"{key} rowsum: {m}")
    torch.testing.assert_close(g2.double().sum(0).float(), want["colsum"], rtol=rtol, atol=atol + 1e-4 * scale, msg=lambda m: f"{key} colsum: {m}")


@pytest.mark.parametrize("name", ["tiny_b5_grads", "cfg1_b3_grads", "cfg3l2_b4_grads", "gpt2_tiny_b5_grads", "gpt2_l2_b3_grads"])
def test_restated_gradients_match_reference_autograd(name):
    """Pins the fine-tune oracle: autograd through oracle/restated.py == gradients of the unmodified reference model."""
    fix = load_golden(name)
    m = T.MultiModalTrajectoryModel(**fix["model_cfg"])
    sd = m.state_dict()
    T.deterministic_fill_(sd, fix["weight_seed"])
    i = fix["inputs"]
    loss, decoded, grads = restated.loss_and_grads(sd, fix["model_cfg"], fix["llama_cfg"], i["x"], i["vision"], i["polygon"], i["poly_len"],
                                                   i["input_ids"], i["attention_mask"], i["y"], i["norm_stat"])
    torch.testing.assert_close(loss, fix["loss"], rtol=1e-5, atol=0)
    assert set(grads) == set(fix["grads"]) and len(grads) == fix["n_trainable"]
    assert {n for n, p in m.named_parameters() if p.requires_grad} == set(grads)
    for k, want in fix["grads"].items():
        ref_scale = float((want["full"] if "full" in want else want["head"]).abs().max()) + 1e-8
        _check_against_compressed(grads[k], want, rtol=2e-3, atol=2e-4 * ref_scale + 1e-6, key=k)


@pytest.mark.parametrize("name", ["tiny_b5_grads_drop", "cfg1_b3_grads_drop", "gpt2_tiny_b5_grads_drop", "gpt2_l2_b3_grads_drop"])
def test_restated_train_mode_dropout_matches_reference(name):
    """Train mode: the restatement with oracle/dropout.py's masks at its dropout sites == the UNMODIFIED reference in train() mode with
    the same masks substituted for torch's RNG (fixture minted through patch_reference_dropout, which also checks that the reference
    makes exactly the dropout calls reference_call_sequence lists, in that order)."""
    from oracle import dropout as OD
    fix = load_golden(name)
    m = T.MultiModalTrajectoryModel(**fix["model_cfg"])
    sd = m.state_dict()
    T.deterministic_fill_(sd, fix["weight_seed"])
    i = fix["inputs"]
    orc = OD.DropOracle(fix["dropout"]["seed"], fix["dropout"]["step"], OD.default_probs(fix["model_cfg"], llama_cfg=fix["llama_cfg"]))
    with restated.dropout(orc):
        loss, decoded, grads = restated.loss_and_grads(sd, fix["model_cfg"], fix["llama_cfg"], i["x"], i["vision"], i["polygon"], i["poly_len"],
                                                       i["input_ids"], i["attention_mask"], i["y"], i["norm_stat"])
    assert restated.DROP is None
    torch.testing.assert_close(loss, fix["loss"], rtol=1e-5, atol=0)
    torch.testing.assert_close(decoded, fix["decoded"], rtol=1e-4, atol=1e-5)
    lc = fix["llama_cfg"]
    assert len(orc.used) == len(OD.reference_call_sequence(fix["model_cfg"], lc["num_hidden_layers"], arch=lc.get("arch", "llama"), llama_cfg=lc))
    for k, want in fix["grads"].items():
        ref_scale = float((want["full"] if "full" in want else want["head"]).abs().max()) + 1e-8
        _check_against_compressed(grads[k], want, rtol=2e-3, atol=2e-4 * ref_scale + 1e-6, key=k)
    # and it is a different function from the eval-mode one
    eval_loss = load_golden(name.replace("_drop", ""))["loss"]
    assert abs(float(loss) - float(eval_loss)) > 1e-4 * float(eval_loss)


def test_dropout_mask_function_and_site_table():
    """The host restatement of the mask function: keep rate, independence across sites / steps, 64-bit indices; and the site-id table
    of the product (train_engine.py) is the one the oracle uses."""
    import numpy as np
    from oracle import dropout as OD
    from tcavp_b200 import ops, train_engine as TE
    assert TE.MOD == OD.MOD and TE.KIND == OD.KIND
    assert TE.site_id("qdec", 3, "ca_attn") == OD.site_id("qdec", 3, "ca_attn") == (3 << 20) | (3 << 8) | 4
    assert ops.drop_threshold(0.1) == OD.drop_threshold(0.1) == 429496730 and ops.drop_threshold(0.0) == 0
    n = 1 << 20
    a = OD.keep_mask(5, 0, OD.site_id("llm", 0, "lora_q"), 0.1, n)
    assert abs(a.mean() - 0.9) < 2e-3
    b = OD.keep_mask(5, 0, OD.site_id("llm", 0, "lora_v"), 0.1, n)
    c = OD.keep_mask(5, 1, OD.site_id("llm", 0, "lora_q"), 0.1, n)
    for other in (b, c):
        assert abs((a & other).mean() - 0.81) < 3e-3                      # independent draws
    assert OD.keep_mask(5, 0, 7, 0.0, 1000).all() and not OD.keep_mask(5, 0, 7, 0.999999, 1000).all()
    assert np.array_equal(a[:4096], OD.keep_mask(5, 0, OD.site_id("llm", 0, "lora_q"), 0.1, 4096))     # a pure function of the index
    # lag-1 / row-structured correlations of consecutive indices stay at the noise floor
    x = a.astype(np.float64) - a.mean()
    for lag in (1, 2, 64, 768, 4096):
        assert abs((x[:-lag] * x[lag:]).mean() / x.var()) < 5e-3, lag


@pytest.mark.parametrize("name", ["tiny_b5_stage1", "cfg1_b2_stage1", "gpt2_tiny_b5_stage1"])
def test_restated_stage1_objective_matches_reference(name):
    """Pins the stage-1 (CausalLM) oracle: restated.causal_lm_loss and its autograd gradients == `outputs.loss` of the unmodified reference
    classes around HF's LlamaForCausalLM (reference scripts/check_generation.py:131-151, train.py:533-547) and its gradients."""
    fix = load_golden(name)
    m = T.MultiModalTrajectoryModel(**fix["model_cfg"])
    sd = m.state_dict()
    T.deterministic_fill_(sd, fix["weight_seed"])
    i = fix["inputs"]
    assert int((i["labels"] != -100).sum()) == fix["n_tokens"] > 0
    loss, grads = restated.stage1_loss_and_grads(sd, fix["model_cfg"], fix["llama_cfg"], i["vision"], i["input_ids"], i["attention_mask"], i["labels"])
    torch.testing.assert_close(loss, fix["loss"], rtol=1e-5, atol=0)
    assert set(grads) == set(fix["grads"]) and len(grads) == fix["n_trainable"]
    for k, want in fix["grads"].items():
        ref = want["full"] if "full" in want else want["head"]
        scale = float(ref.abs().max()) + 1e-8
        _check_against_compressed(grads[k], want, rtol=2e-3, atol=2e-4 * scale + 1e-7, key=k)
