import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a machine without a device: skip them loudly unless selected.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_positions():
    from p3achygo_b200._lib import GO_FEATURES_DTYPE
    z = np.load(os.path.join(GOLDEN, "positions.npz"))
    d = {k: z[k] for k in z.files}
    d["feats"] = np.ascontiguousarray(d["feats"]).view(GO_FEATURES_DTYPE).reshape(-1)
    return d


@pytest.fixture(scope="session")
def known_answers():
    z = np.load(os.path.join(GOLDEN, "known_answers.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def weight_dir(tmp_path_factory):
    """Seeded synthetic weight files, one per config, created on demand."""
    from p3achygo_b200 import weights as W
    d = tmp_path_factory.mktemp("weights")
    cache = {}

    def get(name: str, seed: int = 0):
        key = (name, seed)
        if key not in cache:
            path = os.path.join(d, f"{name}_{seed}.p3w")
            cfg = W.config_from_str(name)
            tensors = W.synthetic_weights(cfg, seed)
            W.save_weights(path, cfg, tensors)
            cache[key] = (path, cfg, tensors)
        return cache[key]

    return get
