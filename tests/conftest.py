import importlib
import json
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
    # GPU tests are skipped (not failed) when collected on a machine without a device
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("dsvt-ai-trt_b200")


@pytest.fixture(scope="session")
def cfgs(pkg):
    return pkg.config


@pytest.fixture(scope="session")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.fixture(scope="session")
def frame0():
    """The reference's sample cloud data/bin/000000.bin (committed fixture; see tools/make_golden.py)."""
    return np.load(os.path.join(GOLDEN, "frame_000000.npz"))["points"]


@pytest.fixture(scope="session")
def attention_case():
    d = dict(np.load(os.path.join(GOLDEN, "attention_case.npz")))
    d["w_in"] = d["w_in"].astype(np.float32)
    d["w_out"] = d["w_out"].astype(np.float32)
    return d


def pad_points(points, cap):
    out = np.zeros((cap, 4), np.float32)
    out[: len(points)] = points[:cap]
    return out
