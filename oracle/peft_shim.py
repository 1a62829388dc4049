"""TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything under oracle/.

Restatement of the slice of `peft` the reference touches (reference scripts/train.py:22, 432-440:
`LoraConfig(r, lora_alpha, lora_dropout, bias="none", task_type=CAUSAL_LM)` + `get_peft_model`).
peft is NOT installed in this image and the reference pins no version, so the published semantics of
peft >= 0.7 `lora.Linear` are restated here ("parity unpinned" at this boundary: no reference test or
golden vector exists for it — see DESIGN.md):

    y = base_layer(x) + lora_B(lora_A(dropout(x))) * (lora_alpha / r)

  * target modules for model_type "llama": q_proj, v_proj (peft's default mapping; explicit in reference
    modify_scripts/modify_train.py:518);
  * lora_A ~ kaiming_uniform(a=sqrt(5)), lora_B = 0, both bias-free;
  * every non-LoRA parameter gets requires_grad=False;
  * the wrapper nests the HF model as `base_model.model`, so state-dict keys read
    `base_model.model.model.layers.N.self_attn.q_proj.{base_layer.weight, lora_A.default.weight,
    lora_B.default.weight}` — the grammar reference ablation_study_without_lora.py:1071-1079 strips.
"""
import math
import types

import torch
import torch.nn as nn

DEFAULT_TARGETS = {"llama": ("q_proj", "v_proj"), "gpt2": ("c_attn",)}


class TaskType:
    CAUSAL_LM = "CAUSAL_LM"


class LoraConfig:
    def __init__(self, r=8, lora_alpha=8, lora_dropout=0.0, bias="none", task_type=None, target_modules=None, **_):
        self.r, self.lora_alpha, self.lora_dropout = r, lora_alpha, lora_dropout
        self.bias, self.task_type, self.target_modules = bias, task_type, target_modules


class LoraLinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int, alpha: float, dropout: float):
        super().__init__()
        self.base_layer = base
        self.lora_dropout = nn.ModuleDict({"default": nn.Dropout(dropout) if dropout > 0 else nn.Identity()})
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)
        self.scaling = alpha / r

    def forward(self, x):
        y = self.base_layer(x)
        return y + self.lora_B["default"](self.lora_A["default"](self.lora_dropout["default"](x))) * self.scaling


class LoraConv1D(nn.Module):
    """peft lora.Linear around transformers' Conv1D (peft sets fan_in_fan_out=True for GPT-2's c_attn): the base layer keeps its
    [in, out] weight and bias, the adapters are ordinary bias-free Linears —  y = base(x) + lora_B(lora_A(dropout(x))) * alpha / r."""

    def __init__(self, base, r: int, alpha: float, dropout: float):
        super().__init__()
        nx, nf = base.weight.shape
        self.base_layer = base
        self.lora_dropout = nn.ModuleDict({"default": nn.Dropout(dropout) if dropout > 0 else nn.Identity()})
        self.lora_A = nn.ModuleDict({"default": nn.Linear(nx, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, nf, bias=False)})
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B["default"].weight)
        self.scaling = alpha / r

    def forward(self, x):
        y = self.base_layer(x)
        return y + self.lora_B["default"](self.lora_A["default"](self.lora_dropout["default"](x))) * self.scaling


class _LoraModel(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model


class PeftModelForCausalLM(nn.Module):
    def __init__(self, model, cfg: LoraConfig):
        super().__init__()
        targets = cfg.target_modules or DEFAULT_TARGETS[getattr(model.config, "model_type", "llama")]
        for p in model.parameters():
            p.requires_grad_(False)
        for parent in list(model.modules()):
            for name, child in list(parent.named_children()):
                if name in targets and isinstance(child, nn.Linear):
                    setattr(parent, name, LoraLinear(child, cfg.r, cfg.lora_alpha, cfg.lora_dropout))
                elif name in targets and type(child).__name__ == "Conv1D":
                    setattr(parent, name, LoraConv1D(child, cfg.r, cfg.lora_alpha, cfg.lora_dropout))
        self.base_model = _LoraModel(model)
        self.config = model.config
        self.peft_config = {"default": cfg}

    def forward(self, *a, **k):
        return self.base_model.model(*a, **k)

    def get_input_embeddings(self):
        return self.base_model.model.get_input_embeddings()

    def generate(self, *a, **k):
        return self.base_model.model.generate(*a, **k)


def get_peft_model(model, cfg):
    return PeftModelForCausalLM(model, cfg)


def as_module():
    m = types.ModuleType("peft")
    m.LoraConfig, m.TaskType, m.get_peft_model = LoraConfig, TaskType, get_peft_model
    m.PeftModelForCausalLM = PeftModelForCausalLM
    return m
