"""oracle/refload.py -- TEST INFRASTRUCTURE ONLY.

Loads the REFERENCE ITSELF for pinning the oracle:

* ``ref_bitboard_class()`` -- the reference's compiled Cython ``OthelloBitboard``
  from ``oracle/_ref`` (built by ``oracle/build.py``; travels to the GPU box as a
  prebuilt .so), imported under its own module name ``src.cython.bitboard``.
* ``reference_python()`` -- the reference's pure-Python modules (MCTS, network,
  batched self-play, trainer, arena, players), unmodified: imported straight from
  /root/reference in the build container (``oracle/gen_golden.py`` writes the
  fixtures in ``tests/golden/`` with them), or from the sourceless byte-code that
  ``oracle/build.py`` compiled into ``oracle/_ref`` (``bytecode=True``; the only
  form that exists on the GPU box, where it is bench.py's ``--impl reference`` arm
  and the caller side of the drop-in tests).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

from . import build as _build


def ref_bitboard_class():
    """The reference's own OthelloBitboard (Cython), or None if not built."""
    path = _build.build_ref()
    if path is None:
        return None
    name = "src.cython.bitboard"
    if name in sys.modules and getattr(sys.modules[name], "__file__", None) == path:
        return sys.modules[name].OthelloBitboard
    # The extension was cythonized as src.cython.bitboard; give it that package context.
    for pkg in ("src", "src.cython"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []          # namespace-like
            sys.modules[pkg] = m
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    sys.modules["src.cython"].bitboard = mod
    return mod.OthelloBitboard


def reference_sources_available() -> bool:
    return os.path.isdir(os.path.join(_build.REFERENCE_ROOT, "src", "mcts"))


def reference_available() -> bool:
    """True when the reference's Python can be imported here: sources (build container) or byte-code (anywhere)."""
    return reference_sources_available() or _build.build_ref_python() is not None


def reference_root(bytecode: bool | None = None) -> str:
    """sys.path entry under which the reference's `src` package lives.  bytecode=None: sources if present."""
    if bytecode is None:
        bytecode = not reference_sources_available()
    if not bytecode:
        if not reference_sources_available():
            raise RuntimeError("reference sources not present (build container only); use bytecode=True")
        return _build.REFERENCE_ROOT
    root = _build.build_ref_python()
    if root is None:
        raise RuntimeError("oracle/_ref holds no byte-compiled reference (run `python -m oracle.build` in the build container)")
    return root


def purge_reference_modules(keep_bitboard: bool = True) -> None:
    """Forget every imported `src.*` module (reference or drop-in shim) so the next import resolves afresh."""
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        if keep_bitboard and k == "src.cython.bitboard":
            continue
        del sys.modules[k]


def reference_python(bytecode: bool | None = None):
    """Import the reference's Python hot-path modules (unmodified) from /root/reference or from oracle/_ref.

    Returns a namespace with MCTS, MCTSNode, BatchMCTS, ParallelSelfPlayWorker,
    SelfPlayWorker, OthelloResNet, OthelloBitboard, ReplayBuffer, Arena, players, AlphaZeroTrainer.
    """
    Board = ref_bitboard_class()
    if Board is None:
        raise RuntimeError("oracle/_ref is not built")
    root = reference_root(bytecode)
    # make the real 'src' package resolvable while keeping the compiled bitboard we just loaded
    bb = sys.modules["src.cython.bitboard"]
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        if k != "src.cython.bitboard":
            del sys.modules[k]
    if root not in sys.path:
        sys.path.insert(0, root)
    import src  # noqa: F401  (the reference's package)
    import src.cython as _c
    sys.modules["src.cython.bitboard"] = bb
    _c.bitboard = bb
    ns = types.SimpleNamespace()
    ns.OthelloBitboard = bb.OthelloBitboard
    ns.MCTSNode = importlib.import_module("src.mcts.node").MCTSNode
    ns.MCTS = importlib.import_module("src.mcts.mcts").MCTS
    psp = importlib.import_module("src.train.parallel_self_play")
    ns.BatchMCTS = psp.BatchMCTS
    ns.ParallelSelfPlayWorker = psp.ParallelSelfPlayWorker
    ns.SelfPlayWorker = importlib.import_module("src.train.self_play").SelfPlayWorker
    ns.OthelloResNet = importlib.import_module("src.model.net").OthelloResNet
    ns.ReplayBuffer = importlib.import_module("src.train.buffer").ReplayBuffer
    arena = importlib.import_module("src.eval.arena")
    players = importlib.import_module("src.eval.players")
    ns.Arena, ns.evaluate_player = arena.Arena, arena.evaluate_player
    ns.GreedyPlayer, ns.RandomPlayer, ns.MCTSPlayer = players.GreedyPlayer, players.RandomPlayer, players.MCTSPlayer
    ns.root = root
    return ns
