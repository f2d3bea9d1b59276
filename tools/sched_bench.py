"""Small-campaign throughput of one schedule: python tools/sched_bench.py --games 100 --schedule async [--sims 50] [--blocks 10 --filters 128]
Prints one JSON line (device-timed campaigns, CUDA events inside oth_selfplay_run)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import othello_reinforcement_learning_test_b200 as pkg
from othello_reinforcement_learning_test_b200.net import OthelloResNet

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=100)
ap.add_argument("--schedule", default="auto")
ap.add_argument("--sims", type=int, default=50)
ap.add_argument("--blocks", type=int, default=10)
ap.add_argument("--filters", type=int, default=128)
ap.add_argument("--c-puct", type=float, default=1.0)
ap.add_argument("--threshold", type=int, default=15)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-cache", action="store_true")
ap.add_argument("--tag", default="")
ap.add_argument("--no-timing", action="store_true", help="no per-kernel CUDA events (their records sit between the launches)")
a = ap.parse_args()
ctx = pkg.Context.default(0)
torch.manual_seed(42)
model = OthelloResNet(a.blocks, a.filters).eval()
w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, torch.device("cuda"), num_simulations=a.sims, temperature_threshold=a.threshold,
                               num_parallel_games=16, c_puct=a.c_puct, concurrent_games=a.games, seed=4242, verbose=False,
                               eval_cache=not a.no_cache, schedule=a.schedule, ctx=ctx)
net = w.batch_mcts._native_net()
eng = w._get_engine(a.games, True)
eng.play(net.handle, a.games)
ctx.timing_enable(not a.no_timing)
ms = ticks = launches = pos = evals = 0
for _ in range(a.reps):
    eng.play(net.handle, a.games)
    st = eng.last_stats
    ms += st["device_ms"]; ticks += st["network_launches"]; launches += st["kernel_launches"]; pos += st["nn_positions"]; evals += eng.last_n_evals
t = ctx.timing_read()
print(json.dumps({"tag": a.tag, "games": a.games, "schedule": eng.last_stats["schedule"], "net": f"{a.blocks}x{a.filters}", "sims": a.sims,
                  "games_per_s": round(a.games * a.reps / (ms / 1e3), 1), "ms_per_campaign": round(ms / a.reps, 1),
                  "network_launches": ticks // a.reps, "us_per_network_launch_step": round(1e3 * ms / ticks, 1),
                  "net_us_per_launch": round(1e3 * t["net"][0] / max(t["net"][1], 1), 1) if not a.no_timing else None,
                  "net_share": round(t["net"][0] / ms, 3), "tree_share": round(t["tree"][0] / ms, 3), "kernel_events": not a.no_timing, "positions_per_expansion": round(pos / evals, 3),
                  "env": {k: v for k, v in os.environ.items() if k.startswith("OTH_")}}))
