import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_bitboard():
    with open(os.path.join(GOLDEN, "bitboard.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_games():
    return dict(np.load(os.path.join(GOLDEN, "ref_games.npz")))


@pytest.fixture(scope="session")
def golden_moves65():
    return dict(np.load(os.path.join(GOLDEN, "ref_moves65.npz")))


@pytest.fixture(scope="session")
def golden_mcts():
    return dict(np.load(os.path.join(GOLDEN, "mcts_ref.npz")))


@pytest.fixture(scope="session")
def golden_selfplay():
    return dict(np.load(os.path.join(GOLDEN, "selfplay_ref.npz")))


@pytest.fixture(scope="session")
def golden_net():
    return dict(np.load(os.path.join(GOLDEN, "net_ref.npz")))


@pytest.fixture(scope="session")
def golden_net_bulk():
    return dict(np.load(os.path.join(GOLDEN, "net_bulk_ref.npz")))


@pytest.fixture(scope="session")
def golden_symmetry():
    return dict(np.load(os.path.join(GOLDEN, "symmetry_ref.npz")))


@pytest.fixture(scope="session")
def ctx():
    import othello_reinforcement_learning_test_b200 as pkg
    return pkg.Context.default(0)
