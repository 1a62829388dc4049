"""tcavp_b200 — B200-native (sm_100a) forward hot path of the Traffic-Context-Augmented Vehicle Trajectory
Prediction model.  Public surface mirrors the reference (reference scripts/train.py:847-964)."""
from .config import LLAMA_PRESETS, MODEL_PRESETS, resolve_llama  # noqa: F401
from .synthetic import make_scenes  # noqa: F401
from .weights import deterministic_fill_  # noqa: F401
from .model import MultiModalTrajectoryModel  # noqa: F401,E402
from . import ops  # noqa: F401,E402
from . import distributed  # noqa: F401,E402
from .finetune import FineTuner  # noqa: F401,E402
from .collate import custom_collate_fn, pack_to_device, ScenePack  # noqa: F401,E402
from .evaluate import evaluate  # noqa: F401,E402
from .prompt_cache import PromptCache  # noqa: F401,E402
from .generate import generate_batch, generate_ids  # noqa: F401,E402
