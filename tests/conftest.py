import os
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def unet_weights():
    """Seeded random-init SD-1.5-architecture UNet weights (fp32, CPU), shared by oracle and engine."""
    from eta_inversion_b200 import synthetic as syn
    return syn.random_state_dict(syn.unet_param_spec(), 0)


@pytest.fixture(scope="session")
def engine_fp32(unet_weights):
    from eta_inversion_b200.engine import UNetEngine
    return UNetEngine(unet_weights, dtype=torch.float32, max_batch=4)


@pytest.fixture(scope="session")
def engine_fp16(unet_weights):
    from eta_inversion_b200.engine import UNetEngine
    return UNetEngine(unet_weights, dtype=torch.float16, max_batch=4)
