"""CPU suite: the C-ABI library loads and exports every symbol include/othello_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "othello_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oth_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from othello_reinforcement_learning_test_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # and the Python binding covers the same set
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_sample_record_layout():
    from othello_reinforcement_learning_test_b200 import _lib
    dt = _lib.SAMPLE_DTYPE
    assert dt.itemsize == 168
    assert dt.fields["visits"][1] == 32 and dt.fields["value"][1] == 30 and dt.fields["game"][1] == 24
    assert ctypes.sizeof(_lib.SelfPlayConfig) == 56


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import othello_reinforcement_learning_test_b200 as pkg
    with pytest.raises(pkg.OthelloB200Error):
        pkg.Context(0)
    b = pkg.OthelloBitboard()          # constructing is free of device work ...
    with pytest.raises(pkg.OthelloB200Error):
        b.get_legal_moves()            # ... but every rule evaluation needs the GPU


def test_product_package_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "othello_reinforcement_learning_test_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "ref_rules.h" not in src and "libref_rules" not in src, f
