"""Batched arena: the reference's evaluation matches, all games advancing together on the GPU.

Mirrors `src/eval/arena.py` (`MatchResult` :14-51, `Arena.play_game` :68-162, `play_matches` :164-202,
`evaluate_player` :235-271) and the players of `src/eval/players.py` (Random :49-67, Greedy :70-113,
MCTS :116-157).  The reference plays `num_games` sequentially with per-move Python calls; here every game of
a match is a column of SoA board arrays, and per ply one batched call per player type chooses all moves
(`oth_choose_random`, `oth_choose_greedy`, or one lock-step MCTS search over all boards to move).

Result conventions are the reference's, quirks included: `winner_color = board.get_winner()` is taken from the
side to move at the terminal position but read as "black" (arena.py:124-148), scores are
`get_stone_counts()` at the end read as (black, white), `num_moves` counts passes.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass
from typing import List

import numpy as np

from . import _lib
from . import bitboard as bb
from ._lib import MEM_HOST, Context, check, ptr
from .mcts import MCTS


@dataclass
class MatchResult:
    """arena.py:14-51"""
    player1_name: str
    player2_name: str
    winner: int
    player1_score: int
    player2_score: int
    num_moves: int
    duration: float

    def __str__(self) -> str:
        res = f"{self.player1_name} wins" if self.winner == 1 else (f"{self.player2_name} wins" if self.winner == -1 else "Draw")
        return (f"{res} | {self.player1_name}: {self.player1_score} - {self.player2_name}: {self.player2_score} | "
                f"Moves: {self.num_moves} | Time: {self.duration:.2f}s")


class BatchPlayer:
    """A player that moves in many games at once: get_actions(self_b, opp_b, move_count) -> int32 actions."""
    name = "Player"

    def reset(self):
        pass

    def get_actions(self, self_b: np.ndarray, opp_b: np.ndarray, move_count: np.ndarray, game_ids: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    # single-board form, so the same object also works with the reference's own Arena
    def get_action(self, board) -> int:
        s = np.array([board.self_board], np.uint64); o = np.array([board.opp_board], np.uint64)
        return int(self.get_actions(s, o, np.array([board.move_count], np.int32), np.zeros(1, np.int64))[0])


class RandomPlayer(BatchPlayer):
    """players.py:49-67 (own RNG: mix64 of seed, game, ply)"""

    def __init__(self, name: str = "Random", seed: int = 0, ctx: Context | None = None):
        self.name, self.seed, self.ctx = name, seed, ctx

    def get_actions(self, self_b, opp_b, move_count, game_ids):
        ctx = self.ctx or Context.default()
        salt = (game_ids.astype(np.uint64) << np.uint64(16)) ^ move_count.astype(np.uint64)
        out = np.empty(self_b.size, np.int32)
        check(ctx.lib.oth_choose_random(ctx.handle, ptr(self_b), ptr(opp_b), ptr(salt), int(self.seed) & (2**64 - 1), ptr(out),
                                        self_b.size, MEM_HOST))
        return out


class GreedyPlayer(BatchPlayer):
    """players.py:70-113, scoring rule as written there"""

    def __init__(self, name: str = "Greedy", ctx: Context | None = None):
        self.name, self.ctx = name, ctx

    def get_actions(self, self_b, opp_b, move_count, game_ids):
        ctx = self.ctx or Context.default()
        out = np.empty(self_b.size, np.int32)
        mc = np.ascontiguousarray(move_count, np.int32)
        check(ctx.lib.oth_choose_greedy(ctx.handle, ptr(self_b), ptr(opp_b), ptr(mc), ptr(out), self_b.size, MEM_HOST))
        return out


class MCTSPlayer(BatchPlayer):
    """players.py:116-157: get_best_action with c_puct 1.0, no noise, temperature 0 -- for all boards at once."""

    def __init__(self, model, device, num_simulations: int = 50, name: str = "MCTS-AI", **mcts_kwargs):
        self.name, self.model, self.device, self.num_simulations = name, model, device, num_simulations
        self.mcts = MCTS(model=model, device=device, c_puct=1.0, **mcts_kwargs)

    @classmethod
    def from_checkpoint(cls, checkpoint_path: str, device, num_simulations: int = 50):
        """players.py:159-222: network shape read off the checkpoint's state_dict (filters from the stem's weight,
        blocks from the highest `res_blocks.N`), falling back to its `config`, then to 10 x 128."""
        import torch
        from .net import OthelloResNet
        checkpoint = torch.load(checkpoint_path, map_location="cpu")
        config = checkpoint.get("config", {})
        sd = checkpoint["model_state_dict"]
        num_filters = sd["conv_block.conv.weight"].shape[0] if "conv_block.conv.weight" in sd else config.get("num_filters", 128)
        num_blocks = max((int(k.split(".")[1]) + 1 for k in sd if k.startswith("res_blocks.")), default=0)
        if num_blocks == 0:
            num_blocks = config.get("num_blocks", 10)
        print(f"Detected model config: blocks={num_blocks}, filters={num_filters}")
        model = OthelloResNet(num_blocks=num_blocks, num_filters=num_filters)
        model.load_state_dict(sd)
        model.eval()
        return cls(model=model, device=device, num_simulations=num_simulations, name=f"MCTS-AI-{num_simulations}sim")

    def get_actions(self, self_b, opp_b, move_count, game_ids):
        legal = bb.legal_moves(self_b, opp_b, self.mcts._context())
        if self.num_simulations < 1:                                   # mcts.py:278-279: first legal move
            return np.array([64 if m == 0 else (int(m) & -int(m)).bit_length() - 1 for m in legal], np.int32)
        vis, _, _ = self.mcts.search_arrays(self_b, opp_b, self.num_simulations)
        # temperature 0: one-hot at the first most-visited child (node.py:171-174), then first max over legal (mcts.py:286-296)
        masked = np.where(vis > 0, vis, 0)
        act = masked.argmax(axis=1).astype(np.int32)
        none = masked.max(axis=1) == 0                                  # zero simulations worth of visits: first child
        if none.any():
            first = np.array([64 if m == 0 else (int(m) & -int(m)).bit_length() - 1 for m in legal], np.int32)
            act[none] = first[none]
        return act


class BatchArena:
    """`Arena` (arena.py:54-233) with all games of a match in flight together."""

    def __init__(self, verbose: bool = False, ctx: Context | None = None):
        self.verbose = verbose
        self.ctx = ctx or Context.default()

    def play_game(self, player1: BatchPlayer, player2: BatchPlayer, starting_player: int = 1) -> MatchResult:
        """arena.py:68-162: one game; starting_player = 1 (player1 moves first) or -1."""
        result = self._play(player1, player2, np.array([1 if starting_player == 1 else -1]))[0]
        if self.verbose:
            print(f"\n{result}\n")
        return result

    def play_matches(self, player1: BatchPlayer, player2: BatchPlayer, num_games: int = 10,
                     alternate_colors: bool = True) -> List[MatchResult]:
        """arena.py:164-202"""
        n = int(num_games)
        starting = np.where((np.arange(n) % 2 == 0) | (not alternate_colors), 1, -1)        # arena.py:187-190
        results = self._play(player1, player2, starting)
        if self.verbose:
            self._print_summary(results, player1.name, player2.name)
        return results

    def _print_summary(self, results: List[MatchResult], player1_name: str, player2_name: str) -> None:
        """arena.py:204-233"""
        total = len(results)
        w1 = sum(1 for r in results if r.winner == 1); w2 = sum(1 for r in results if r.winner == -1)
        print("\n" + "=" * 70 + "\nMatch Summary\n" + "=" * 70)
        print(f"\nTotal Games: {total}")
        print(f"{player1_name}: {w1} wins ({w1 / total * 100 if total else 0:.1f}%)")
        print(f"{player2_name}: {w2} wins ({w2 / total * 100 if total else 0:.1f}%)")
        print(f"Draws: {total - w1 - w2}")
        print(f"\nAverage Moves: {sum(r.num_moves for r in results) / total if total else 0:.1f}")
        print(f"Average Duration: {sum(r.duration for r in results) / total if total else 0:.2f}s")
        print("=" * 70 + "\n")

    def _play(self, player1: BatchPlayer, player2: BatchPlayer, starting: np.ndarray) -> List[MatchResult]:
        """All games of `starting` (one entry per game: +1 = player1 moves first) in flight together."""
        n = int(starting.size)
        t0 = time.time()
        player1.reset(); player2.reset()
        s = np.full(n, bb.START_SELF, np.uint64); o = np.full(n, bb.START_OPP, np.uint64); mc = np.zeros(n, np.int32)
        to_move = starting.copy()                       # +1: player1 moves, -1: player2 moves (arena.py:91-100)
        live = np.ones(n, bool)
        gid = np.arange(n, dtype=np.int64)
        term, _, _ = bb.terminal_winner(s, o, self.ctx)
        live &= term == 0
        while live.any():
            for who, player in ((1, player1), (-1, player2)):
                sel = np.flatnonzero(live & (to_move == who))
                if sel.size == 0:
                    continue
                acts = player.get_actions(s[sel].copy(), o[sel].copy(), mc[sel].copy(), gid[sel])
                s2, o2, m2 = s[sel].copy(), o[sel].copy(), mc[sel].copy()
                ok = bb.make_move(s2, o2, m2, acts, self.ctx)          # rejected moves leave the board as is (arena.py:116)
                s[sel], o[sel], mc[sel] = s2, o2, m2
                # arena.py:118-119 swaps the players after every call, accepted or not; an illegal move would spin
                if not ok.all():
                    raise RuntimeError(f"{player.name} returned an illegal move")
            to_move[live] = -to_move[live]
            term, _, _ = bb.terminal_winner(s, o, self.ctx)
            live &= term == 0
        _, win, counts = bb.terminal_winner(s, o, self.ctx)
        dur = (time.time() - t0) / max(n, 1)
        results = []
        for i in range(n):
            wc = int(win[i])                                            # read as "black's" result (arena.py:124)
            black, white = int(counts[i, 0]), int(counts[i, 1])         # get_stone_counts() read as (black, white)
            if starting[i] == 1:
                w, p1, p2 = wc, black, white                            # arena.py:128-138
            else:
                w, p1, p2 = -wc, white, black                           # arena.py:139-147
            results.append(MatchResult(player1.name, player2.name, w, p1, p2, int(mc[i]), dur))
        return results


Arena = BatchArena      # the reference's class name (src/eval/arena.py:54)


def evaluate_player(player: BatchPlayer, opponent: BatchPlayer, num_games: int = 10, verbose: bool = False) -> dict:
    """arena.py:235-271"""
    results = BatchArena(verbose=verbose).play_matches(player, opponent, num_games=num_games)
    wins = sum(1 for r in results if r.winner == 1)
    return {"win_rate": wins / num_games if num_games > 0 else 0,
            "avg_score": sum(r.player1_score for r in results) / num_games if num_games > 0 else 0,
            "avg_moves": sum(r.num_moves for r in results) / num_games if num_games > 0 else 0,
            "results": results}
