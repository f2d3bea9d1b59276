"""othello_reinforcement_learning_test_b200 -- B200-native (sm_100a) self-play hot path.

The reference's API surface for the path bitboard -> MCTS -> ResNet leaf evaluation
(SURVEY.md section 8), re-implemented as hand-written CUDA behind a C ABI
(include/othello_b200.h, csrc/):

    OthelloBitboard, BoardBatch          src/cython/bitboard.pyx
    MCTS, BatchMCTS                      src/mcts/mcts.py, src/train/parallel_self_play.py
    OthelloResNet, InferenceNet          src/model/net.py
    SelfPlayWorker, ParallelSelfPlayWorker, create_parallel_self_play_worker
                                         src/train/self_play.py, src/train/parallel_self_play.py
    ReplayBuffer                         src/train/buffer.py (device-resident, packed samples)
    BatchArena, Random/Greedy/MCTSPlayer src/eval/arena.py, src/eval/players.py (all games of a match in flight together)
    dropin.install()                     makes `from src.cython.bitboard import OthelloBitboard`
                                         etc. resolve to this package

Importing this package loads libothello_b200.so; there is no CPU fallback.
"""
from . import _lib
from ._lib import Context, OthelloB200Error

_lib.load()   # fail loudly, right here, if the CUDA library is missing

from .bitboard import BoardBatch, OthelloBitboard   # noqa: E402
from .mcts import MCTS, BatchMCTS   # noqa: E402
from .net import InferenceNet, OthelloResNet, create_model   # noqa: E402
from .buffer import PrioritizedReplayBuffer, ReplayBuffer   # noqa: E402
from .arena import Arena, BatchArena, GreedyPlayer, MatchResult, MCTSPlayer, RandomPlayer, evaluate_player   # noqa: E402
from .self_play import (ParallelSelfPlayWorker, SelfPlayWorker, augment_data_with_symmetries,   # noqa: E402
                        create_parallel_self_play_worker)

__all__ = [
    "Context", "OthelloB200Error", "OthelloBitboard", "BoardBatch", "MCTS", "BatchMCTS", "OthelloResNet",
    "InferenceNet", "create_model", "SelfPlayWorker", "ParallelSelfPlayWorker", "create_parallel_self_play_worker",
    "augment_data_with_symmetries", "ReplayBuffer", "PrioritizedReplayBuffer", "Arena", "BatchArena", "MatchResult", "RandomPlayer",
    "GreedyPlayer", "MCTSPlayer", "evaluate_player",
]
__version__ = "0.1.0"
