import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name)))

    return load


@pytest.fixture(scope="session")
def synth_sd():
    from rgrg_b200 import synth

    return synth.make_state_dict(0)


@pytest.fixture(scope="session")
def lm_sd():
    """LM + selection-head weights only: RNG-only tensors, bit-identical on every host, seconds to build."""
    from rgrg_b200 import synth

    return synth.make_partial_state_dict(0, ("heads", "lm"))
