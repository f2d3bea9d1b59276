"""oracle/ref_arm.py -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Times the UNMODIFIED REFERENCE on the host cores for bench.py's `--impl reference` arm and its
`cpu_baseline` leg: the reference's own `ParallelSelfPlayWorker._execute_batch`
(src/train/parallel_self_play.py:324-407) with its own `BatchMCTS`, `MCTSNode`, `OthelloResNet`
(PyTorch CPU, fp32) and compiled Cython `OthelloBitboard`, imported from `oracle/_ref` (the
Cython module and sourceless byte-code that `oracle/build.py` compiled from the reference
sources in place; nothing of this package is on that path).

A full batch of 16 games takes the reference tens of seconds on a host CPU, a benchmark step
must be a BOUNDED sample: the worker is handed a `board_class` that subclasses the reference's
board and reports `is_terminal()` once the step's time budget is over.  Only the top-level game
boards are of that class (`copy()` returns plain reference boards, bitboard.pyx:324-335), so
every search inside the budget is exactly the reference's; the batch simply stops after the ply
in flight.  Work done is counted in plies (one ply = one full search of 1 + num_simulations
network evaluations per game) and converted to games with the measured mean game length.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import refload

MEAN_PLIES = 60.5          # mean length of a self-play game (reference: 60.2 random playouts, ~61 self-play)


class ReferenceArm:
    def __init__(self, num_blocks=10, num_filters=128, num_simulations=50, c_puct=1.0, temperature_threshold=15,
                 num_parallel_games=16, threads=None, seed=42):
        self.ref = refload.reference_python(bytecode=True)          # the same byte-code here and on the GPU box
        self.threads = int(threads or os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        torch.manual_seed(seed)                                     # BASELINE configs: random-init weights, seed 42
        np.random.seed(seed)
        self.model = self.ref.OthelloResNet(num_blocks=num_blocks, num_filters=num_filters).eval()
        Board = self.ref.OthelloBitboard
        arm = self

        class BudgetBoard(Board):
            """The reference's board; additionally reports 'terminal' when the step's time budget is used up."""

            def is_terminal(self):
                return time.perf_counter() >= arm.deadline or Board.is_terminal(self)

        self.board_class = BudgetBoard
        self.parallel = int(num_parallel_games)
        self.worker = self.ref.ParallelSelfPlayWorker(
            board_class=BudgetBoard, model=self.model, device=torch.device("cpu"), num_simulations=num_simulations,
            temperature_threshold=temperature_threshold, num_parallel_games=self.parallel, c_puct=c_puct,
            dirichlet_alpha=0.3, dirichlet_epsilon=0.25)
        self.sims = int(num_simulations)
        self.deadline = float("inf")

    def state_dict(self):
        return self.model.state_dict()

    def step(self, time_budget_s: float) -> dict:
        """One bounded sample: a lock-step batch of `num_parallel_games` games from the start position, played by the
        reference until the budget is over (or the games end)."""
        t0 = time.perf_counter()
        self.deadline = t0 + float(time_budget_s)
        data = self.worker._execute_batch(self.parallel, True)      # parallel_self_play.py:324-407, unmodified
        dt = time.perf_counter() - t0
        plies = len(data)
        return {"plies": plies, "seconds": dt, "games_per_s": plies / dt / MEAN_PLIES, "threads": self.threads,
                "sample": f"reference ParallelSelfPlayWorker._execute_batch: {self.parallel} lock-step games x "
                          f"{plies // max(self.parallel, 1)} plies ({plies} searches of 1+{self.sims} evals, PyTorch CPU fp32, "
                          f"{self.threads} threads) in {dt:.1f}s; games/s = plies/s / {MEAN_PLIES} plies per game"}
