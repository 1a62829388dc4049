import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


@pytest.fixture(scope="session")
def lib_built():
    import tcavp_b200.lib as L
    L.build()
    return L.load()


def build_filled_model(fix, compute_dtype, device):
    import tcavp_b200 as T
    m = T.MultiModalTrajectoryModel(**fix["model_cfg"], compute_dtype=compute_dtype)
    sd = m.state_dict()
    T.deterministic_fill_(sd, fix["weight_seed"])
    m.load_state_dict(sd, strict=True)
    return m.to(device).eval()
