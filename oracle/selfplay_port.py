"""oracle/selfplay_port.py -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

CPU port of the reference's batched self-play (ParallelSelfPlayWorker._execute_batch,
src/train/parallel_self_play.py:324-407): N games in lock-step, one batched fp32 network call per
simulation step (BatchMCTS.batch_predict, :53-78), tree operations by the C oracle
(oracle/ref_rules.c), move choice with numpy's RNG exactly as the reference does (:379-382).

Used by bench.py as the `cpu_baseline` / `--impl reference` arm (kind "port": the reference's own
Python cannot travel to the GPU box; this port does the same work per game with less interpreter
overhead, so it is a conservative -- i.e. fast -- stand-in) and by the tests as the checker for
whole-campaign statistics.  Never imported by the product package.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import cref, net_oracle


class CpuSelfPlay:
    def __init__(self, state_dict=None, num_simulations=50, temperature_threshold=15, num_parallel_games=16,
                 c_puct=1.0, seed=0, threads=None):
        self.sd = state_dict
        self.sims, self.threshold, self.parallel, self.c_puct = num_simulations, temperature_threshold, num_parallel_games, c_puct
        self.rng = np.random.default_rng(seed)
        if threads:
            torch.set_num_threads(int(threads))
        self.threads = torch.get_num_threads()
        self.search = cref.BatchSearch(num_parallel_games, num_simulations, c_puct)
        self.n_evals = 0
        self.plies = 0

    def _evaluate(self, ls, lo):
        self.n_evals += int(ls.size)
        if self.sd is None:
            return cref.hashnet_batch(ls, lo)
        x = net_oracle.boards_to_tensor(ls, lo)                      # get_tensor_input x N (:64-66)
        logp, v = net_oracle.forward_fp32(self.sd, x)                # model(batch) in fp32 eval mode (:72-74)
        return torch.exp(logp).numpy(), v.numpy().reshape(-1)        # exp(log_probs) (:75)

    def play(self, n_games: int, max_plies: int | None = None, time_budget_s: float | None = None):
        """Lock-step batch of `n_games` (<= num_parallel_games).  Stops early after `max_plies` plies
        per game or when the time budget is used up (bounded sample for the benchmark).
        Returns dict(samples=[(self, opp, visits, player)], values, finished, plies, seconds)."""
        t0 = time.perf_counter()
        s = np.full(n_games, cref.START_SELF, np.uint64); o = np.full(n_games, cref.START_OPP, np.uint64)
        mc = np.zeros(n_games, np.int32)
        live = np.ones(n_games, bool)
        hist = [[] for _ in range(n_games)]
        winners = np.zeros(n_games, np.int64)
        ply = 0
        while live.any():
            if max_plies is not None and ply >= max_plies:
                break
            if time_budget_s is not None and time.perf_counter() - t0 > time_budget_s:
                break
            idx = np.flatnonzero(live)
            vis, _ = self.search.search(s[idx], o[idx], self.sims, self._evaluate)
            for j, g in enumerate(idx):
                v = vis[j]
                pol = v.astype(np.float32) / np.float32(v.sum())                              # T = 1 policy (:164-168)
                if mc[g] < self.threshold:
                    a = int(self.rng.choice(65, p=pol.astype(np.float64) / pol.astype(np.float64).sum()))
                else:
                    a = int(np.argmax(pol))
                hist[g].append((int(s[g]), int(o[g]), v.copy(), 1 if mc[g] % 2 == 0 else -1))
                ok, s2, o2, m2 = cref.make_move(int(s[g]), int(o[g]), int(mc[g]), a)
                assert ok
                s[g], o[g], mc[g] = s2, o2, m2
                self.plies += 1
                if cref.is_terminal(s2, o2):
                    live[g] = False
                    winners[g] = cref.winner(s2, o2)
            ply += 1
        finished = ~live
        values = [[float(winners[g] * pl) for (_, _, _, pl) in hist[g]] if finished[g] else None for g in range(n_games)]
        return {"samples": hist, "values": values, "finished": finished, "plies": int(sum(len(h) for h in hist)),
                "seconds": time.perf_counter() - t0}


def measure_games_per_second(state_dict, num_simulations=50, c_puct=1.0, temperature_threshold=15,
                             num_parallel_games=16, time_budget_s=20.0, threads=None, mean_plies_per_game=60.5, seed=0):
    """Bounded CPU sample -> games/s.  Plays `num_parallel_games` lock-step games from the start position
    for about `time_budget_s` seconds, counts plies (each ply = one full 1+sims-evaluation search per game)
    and converts with the mean game length."""
    sp = CpuSelfPlay(state_dict, num_simulations, temperature_threshold, num_parallel_games, c_puct, seed, threads)
    r = sp.play(num_parallel_games, time_budget_s=time_budget_s)
    plies_per_s = r["plies"] / r["seconds"]
    return {"games_per_s": plies_per_s / mean_plies_per_game, "plies": r["plies"], "seconds": r["seconds"],
            "nn_evals": sp.n_evals, "threads": sp.threads,
            "sample": f"{num_parallel_games} lock-step games x {r['plies'] // max(num_parallel_games, 1)} plies "
                      f"({r['plies']} searches of 1+{num_simulations} evals) in {r['seconds']:.1f}s; "
                      f"games/s = plies/s / {mean_plies_per_game} plies per game"}
