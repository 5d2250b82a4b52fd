import os
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def kodak():
    """(imgs f32 [24,3,224,224] in [0,1], scores f32 [24,196]) - the reference's test-time inputs."""
    import numpy as np
    z = np.load(GOLDEN / "kodak_224.npz")
    imgs = torch.from_numpy(z["imgs"]).permute(0, 3, 1, 2).float() / 255.0
    scores = torch.load(GOLDEN / "kodak_scores.pt")
    return imgs, scores


@pytest.fixture(scope="session")
def cuda_dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
