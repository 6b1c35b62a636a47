import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (sm_100) GPU; run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
    arrays = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return meta, arrays


@pytest.fixture(scope="session")
def golden_cases():
    return {n: load_golden(n) for n in ("tiny_32x64_b2", "tiny_128x256_b1")}


def rel_err(a, b):
    """max|a-b| / max|b| -- the fp32-path metric of BASELINE.json's north_star."""
    a = a.detach().float().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float32)
    b = b.detach().float().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float32)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
