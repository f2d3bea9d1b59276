"""Kernel-only run of k_playouts (BASELINE config 1) -- the command ncu captures for the bitboard kernel evidence."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import othello_reinforcement_learning_test_b200 as pkg
from othello_reinforcement_learning_test_b200 import bitboard as bb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
ctx = pkg.Context.default(0)
bb.random_playouts(1 << 18, seed=1, ctx=ctx)
t0 = time.perf_counter()
r = bb.random_playouts(n, seed=2, ctx=ctx)
dt = time.perf_counter() - t0
print(json.dumps({"kernel": "k_playouts", "games": n, "games_per_s": n / dt, "mean_plies": r["total_plies"] / n}))
