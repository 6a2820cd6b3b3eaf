import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    import torch

    # parity runs compare against an fp32 CPU oracle: keep library convs / matmuls in true fp32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    path = os.path.join(REPO, "tests", "golden", "reference_modules.npz")
    with np.load(path) as data:
        return {k: data[k] for k in data.files}
