"""Kernel-only timing of k_net_tc on a resident batch (CUDA events on the library's stream).
Also the command ncu captures for the tensor-pipe evidence (a full, representative launch)."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import othello_reinforcement_learning_test_b200 as pkg
from othello_reinforcement_learning_test_b200.net import InferenceNet, OthelloResNet

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=148 * 4 * 32)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--blocks", type=int, default=10)
ap.add_argument("--filters", type=int, default=128)
ap.add_argument("--engine", default=None, help="tcgen05 | tcgen05_pair | simt (default: the library's choice)")
a = ap.parse_args()
ctx = pkg.Context.default(0)
torch.manual_seed(42)
net = InferenceNet.from_module(OthelloResNet(a.blocks, a.filters).eval(), ctx, engine=a.engine)
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_games.npz"))
live = np.flatnonzero(g["terminal"] == 0)
idx = live[np.arange(a.n) % live.size]
S = torch.from_numpy(g["self_b"][idx].view(np.int64)).cuda(); O = torch.from_numpy(g["opp_b"][idx].view(np.int64)).cuda()
stream = torch.cuda.ExternalStream(ctx.stream)
for _ in range(3):
    net.forward(S, O, out="priors")
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(a.reps):
    net.forward(S, O, out="priors")
e1.record(stream); e1.synchronize()
ms = e0.elapsed_time(e1) / a.reps
F, B = a.filters, a.blocks
flop = 2 * (64 * 27 * F + B * 2 * 64 * 9 * F * F + 64 * 2 * F + 128 * 65 + 64 * F + 64 * 256 + 256)
print(json.dumps({"kernel": "k_net_tc", "engine": a.engine or "default", "positions": a.n, "ms_per_launch": ms, "tflops": a.n * flop / ms / 1e9,
                  "positions_per_s": a.n / ms * 1e3}))
