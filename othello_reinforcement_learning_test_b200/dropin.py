"""Make the reference's own import paths resolve to this package.

The reference's callers import by absolute path (`from src.cython.bitboard import
OthelloBitboard` in src/eval/arena.py:10, src/eval/players.py:17, main.py:13;
`from src.mcts.mcts import MCTS` in players.py:145; `from src.train.parallel_self_play import
create_parallel_self_play_worker` in main.py:112).  `install()` pre-seeds `sys.modules` with
modules of those names backed by this package, so arena.py / players.py / self_play.py /
trainer.py run unchanged on top of the CUDA engine:

    import othello_reinforcement_learning_test_b200.dropin as dropin
    dropin.install()                 # before the reference modules are imported
    from src.eval.arena import Arena # reference code, now driving the B200 kernels

Modules of the reference that are NOT on the hot path (trainer, buffer, arena, players, GUI,
web) are left alone: if the reference tree is on sys.path they import from there.
"""
from __future__ import annotations

import importlib
import sys
import types

_REPLACED = {
    "src.cython.bitboard": ("othello_reinforcement_learning_test_b200.bitboard", ["OthelloBitboard"]),
    "src.mcts.mcts": ("othello_reinforcement_learning_test_b200.mcts", ["MCTS"]),
    "src.model.net": ("othello_reinforcement_learning_test_b200.net", ["OthelloResNet", "create_model"]),
    "src.train.parallel_self_play": ("othello_reinforcement_learning_test_b200.self_play",
                                     ["ParallelSelfPlayWorker", "BatchMCTS", "create_parallel_self_play_worker"]),
    "src.train.self_play": ("othello_reinforcement_learning_test_b200.self_play",
                            ["SelfPlayWorker", "augment_data_with_symmetries"]),
}


def _ensure_package(name: str) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        try:
            mod = importlib.import_module(name)          # the reference's real package, if importable
        except Exception:
            mod = types.ModuleType(name)
            mod.__path__ = []                             # namespace stand-in
            sys.modules[name] = mod
    return mod


_OPTIONAL = {
    # the rows SURVEY.md 8(f) marks "next": opt in with install(replay_buffer=True, arena=True)
    "replay_buffer": {"src.train.buffer": ("othello_reinforcement_learning_test_b200.buffer", ["ReplayBuffer", "PrioritizedReplayBuffer"])},
    "arena": {"src.eval.arena": ("othello_reinforcement_learning_test_b200.arena", ["MatchResult", "Arena", "BatchArena", "evaluate_player"]),
              "src.eval.players": ("othello_reinforcement_learning_test_b200.arena", ["RandomPlayer", "GreedyPlayer", "MCTSPlayer"])},
}


def install(force: bool = True, replay_buffer: bool = False, arena: bool = False, only: list | None = None) -> None:
    """Route the hot-path module names of the reference to this package.

    only=[...]: route just these module names (e.g. ["src.cython.bitboard", "src.mcts.mcts", "src.model.net"] keeps the
    reference's own self-play workers, which then drive this package's board and search).

    replay_buffer=True also routes `src.train.buffer` (trainer.py imports ReplayBuffer from there) to the
    device-resident buffer; arena=True routes `src.eval.arena` / `src.eval.players` to the batched arena
    (its `Arena` name is `BatchArena`; players take and return whole batches as well as single boards)."""
    extra = {}
    if replay_buffer:
        extra.update(_OPTIONAL["replay_buffer"])
    if arena:
        extra.update(_OPTIONAL["arena"])
    table = {**_REPLACED, **extra}
    if only is not None:
        unknown = [n for n in only if n not in table]
        if unknown:
            raise ValueError(f"not a routable reference module: {unknown}")
        table = {n: table[n] for n in only}
    _install_table(table, force)


def _install_table(table, force):
    from . import mcts as _mcts, self_play as _sp   # noqa: F401
    _sp.BatchMCTS = _mcts.BatchMCTS                  # the reference keeps BatchMCTS in parallel_self_play.py
    for name, (target, names) in table.items():
        if not force and name in sys.modules:
            continue
        parts = name.split(".")
        for i in range(1, len(parts)):
            _ensure_package(".".join(parts[:i]))
        src_mod = importlib.import_module(target)
        shim = types.ModuleType(name)
        shim.__doc__ = f"B200 drop-in for the reference module {name} (backed by {target})"
        for n in names:
            setattr(shim, n, getattr(src_mod, n))
        shim.__b200_dropin__ = True
        sys.modules[name] = shim
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], shim)


def uninstall() -> None:
    for name in list(_REPLACED) + [k for t in _OPTIONAL.values() for k in t]:
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__b200_dropin__", False):
            del sys.modules[name]
