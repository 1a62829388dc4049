"""TEST INFRASTRUCTURE — not product code (see oracle/README.md).

Imports the UNMODIFIED reference script from /root/reference (read-only, exists only in the authoring
container — never on the GPU box) so that (a) the plain-tensor restatement in oracle/restated.py can be
validated against the real reference classes and (b) golden vectors can be minted (oracle/make_golden.py).

Three shims are needed because the image lacks the reference's unpinned third-party deps:
  1. `matplotlib` / `matplotlib.pyplot` stubs (viz only, reference scripts/train.py:17-19);
  2. a `peft` module (oracle/peft_shim.py restates lora.Linear + the key layout);
  3. `AutoModelForCausalLM.from_pretrained` -> `LlamaForCausalLM(LlamaConfig(**llama_cfg))` (no hub, no
     network) and `AutoTokenizer.from_pretrained` -> a stub (only pad/eos are touched when input_ids are
     given, reference scripts/train.py:500-502).
"""
import importlib.util
import os
import sys
import types

import torch

from . import peft_shim

REFERENCE_ROOT = os.environ.get("TCAVP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "scripts", "train.py"))


class _Tok:
    pad_token = None
    eos_token = "</s>"

    def __call__(self, *a, **k):
        raise RuntimeError("oracle tokenizer stub: pass input_ids/attention_mask explicitly")


def hf_llama_config(llama_cfg: dict):
    from transformers import LlamaConfig
    c = dict(llama_cfg)
    kw = dict(vocab_size=c["vocab_size"], hidden_size=c["hidden_size"], intermediate_size=c["intermediate_size"],
              num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
              num_key_value_heads=c.get("num_key_value_heads", c["num_attention_heads"]),
              head_dim=c.get("head_dim", c["hidden_size"] // c["num_attention_heads"]),
              rms_norm_eps=c.get("rms_norm_eps", 1e-6), max_position_embeddings=c.get("max_position_embeddings", 2048),
              tie_word_embeddings=bool(c.get("tie_word_embeddings", False)), attention_bias=False, mlp_bias=False)
    theta = float(c.get("rope_theta", 10000.0))
    rs = c.get("rope_scaling")
    if rs:       # e.g. Llama-3.2-1B: {"rope_type": "llama3", factor, low_freq_factor, high_freq_factor, original_max_position_embeddings}
        kw["rope_parameters"] = dict(rs, rope_theta=theta)
    cfg = LlamaConfig(**kw)
    # transformers 5.x keeps theta inside rope_parameters; older versions as an attribute
    if getattr(cfg, "rope_parameters", None) is not None:
        cfg.rope_parameters["rope_theta"] = theta
    else:
        cfg.rope_theta = theta
        if rs:
            cfg.rope_scaling = dict(rs)
    return cfg


def load_reference(script: str = "scripts/train.py", llama_cfg: dict = None):
    """Executes the reference script as a module (main() is __main__-guarded) and returns it."""
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_ROOT} is not present (expected on the GPU box): use the golden fixtures")
    import transformers
    from transformers import LlamaForCausalLM

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot
    sys.modules["peft"] = peft_shim.as_module()

    cfg_holder = {"cfg": llama_cfg}

    def _from_pretrained(cls, name, **kw):
        c = cfg_holder["cfg"]
        if c.get("arch") == "gpt2":          # what AutoModelForCausalLM resolves a model_type "gpt2" checkpoint to
            from transformers import GPT2Config, GPT2LMHeadModel
            return GPT2LMHeadModel(GPT2Config(vocab_size=c["vocab_size"], n_positions=c["n_positions"], n_embd=c["hidden_size"],
                                              n_layer=c["num_hidden_layers"], n_head=c["num_attention_heads"], n_inner=c["intermediate_size"],
                                              activation_function="gelu_new", layer_norm_epsilon=c.get("layer_norm_epsilon", 1e-5),
                                              resid_pdrop=c.get("resid_pdrop", 0.0), embd_pdrop=c.get("embd_pdrop", 0.0),
                                              attn_pdrop=c.get("attn_pdrop", 0.0)))
        return LlamaForCausalLM(hf_llama_config(c))

    transformers.AutoModelForCausalLM.from_pretrained = classmethod(_from_pretrained)
    transformers.AutoTokenizer.from_pretrained = classmethod(lambda cls, name, **kw: _Tok())

    path = os.path.join(REFERENCE_ROOT, script)
    modname = "tcavp_ref_" + os.path.basename(script).replace(".py", "")
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    # the reference sets MASTER_PORT at import (train.py:24-25); keep the caller's env intact
    saved = os.environ.get("MASTER_PORT")
    spec.loader.exec_module(mod)
    if saved is None:
        os.environ.pop("MASTER_PORT", None)
    else:
        os.environ["MASTER_PORT"] = saved
    mod._tcavp_cfg_holder = cfg_holder
    return mod


def build_reference_model(mod, model_cfg: dict, llama_cfg: dict):
    """`model_cfg` = ctor kwargs of MultiModalTrajectoryModel (reference scripts/train.py:848-872)."""
    mod._tcavp_cfg_holder["cfg"] = llama_cfg
    torch.backends.mha.set_fastpath_enabled(False)  # pin nn.Transformer* to the slow (non-nested) path
    kw = dict(model_cfg)
    kw.setdefault("base_model_name", "synthetic-llama")
    model = mod.MultiModalTrajectoryModel(**kw)
    return model.eval()
