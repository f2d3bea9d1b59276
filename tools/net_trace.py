"""Per-layer timeline of CTA 0 of k_net_tc (clock64 stamps): where does a layer's time go?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import othello_reinforcement_learning_test_b200 as pkg
from othello_reinforcement_learning_test_b200.net import InferenceNet, OthelloResNet

nb, nf = int(sys.argv[1]) if len(sys.argv) > 1 else 10, int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = int(sys.argv[3]) if len(sys.argv) > 3 else 592
ctx = pkg.Context.default(0)
torch.manual_seed(42)
net = InferenceNet.from_module(OthelloResNet(nb, nf).eval(), ctx, engine=(sys.argv[4] if len(sys.argv) > 4 else None))
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_games.npz"))
S = np.ascontiguousarray(g["self_b"][:n]); O = np.ascontiguousarray(g["opp_b"][:n])
L = 1 + 2 * nb
tr = np.zeros(L * 8, np.uint64)
pkg._lib.check(ctx.lib.oth_debug_net_trace(net.handle, S.ctypes.data, O.ctypes.data, n, tr.ctypes.data, tr.size))
tr = tr.reshape(L, 8).astype(np.int64)
t0 = tr[0, 0]
print(f"{nb}x{nf}, n={n}; cycles relative to the first MMA issue of CTA 0")
print("layer  mma_start  mma_issue_done  epi0_start  epi0_end  epi1_start  epi1_end | issue  mma->epi0  epi0  epi1  layer_total")
for l in range(L):
    a = tr[l] - t0
    nxt = (tr[l + 1, 0] - t0) if l + 1 < L else a[5]
    print(f"{l:3d} {a[0]:10d} {a[1]:10d} {a[2]:10d} {a[3]:10d} {a[4]:10d} {a[5]:10d} | {a[1]-a[0]:6d} {a[2]-a[0]:6d} {a[3]-a[2]:6d} {a[5]-a[4]:6d} {nxt-a[0]:7d}")
print("total cycles:", tr[L - 1, 5] - t0)
