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


@pytest.mark.parametrize("name", ["tiny_b6", "cfg1_b8", "cfg5_b32", "cfg3l2_b16", "gqa_l2_b32", "llama32_1b_l2_b8"])
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


@pytest.mark.parametrize("name", ["tiny_b5_grads", "cfg1_b3_grads", "cfg3l2_b4_grads"])
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
