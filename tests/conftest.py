"""Shared pytest configuration: the ``gpu`` marker and fixture helpers."""
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
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_nn_cfg1():
    return load_golden("nn_cfg1.npz")


@pytest.fixture(scope="session")
def golden_nn_small():
    return load_golden("nn_small.npz")


@pytest.fixture(scope="session")
def golden_fm():
    return load_golden("fm_pair_ico3.npz")


@pytest.fixture(scope="session")
def golden_zo():
    return load_golden("zoomout_ico3.npz")


@pytest.fixture(scope="session")
def golden_extras():
    return load_golden("extras_ico3.npz")


@pytest.fixture(scope="session")
def golden_full():
    """Reference outputs at BASELINE size (icosphere(4): 2562 vertices, K = 200, k = 100); the eigenbases are stored as
    float32 and used as float64(float32(.)) by the reference run that minted the file and by everything under test."""
    g = load_golden("fm_full_ico4.npz")
    g["Phi1"] = g.pop("Phi1_f32").astype(np.float64)
    g["Phi2"] = g.pop("Phi2_f32").astype(np.float64)
    return g
