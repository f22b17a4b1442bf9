import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:   # tests/scenes.py
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """-> (arrays dict of torch tensors, state_dict of torch tensors)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    arrays, sd = {}, {}
    for k in z.files:
        if z[k].dtype.kind in "US":  # lists of names stay lists of str
            arrays[k] = [str(v) for v in z[k]]
            continue
        t = torch.from_numpy(z[k])
        if k.startswith("sd::"):
            sd[k[4:]] = t
        else:
            arrays[k] = t
    return arrays, sd


@pytest.fixture
def golden():
    return load_golden
