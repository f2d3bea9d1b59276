#!/usr/bin/env python
"""bench.py -- self-play games/s @50 sims/move, 10x128 ResNet (BASELINE.json metric), on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one self-play campaign: G concurrent games per GPU (default 303104 = 2048 x 148 SMs) started
from the initial position and played to completion with the device-resident engine
(ParallelSelfPlayWorker / oth_selfplay_run): per ply one search of 1 + 50 leaf evaluations per game
(select -> tcgen05 ResNet -> expand/backup), move choice, trajectory recording, labelling.
Every timed step is measured twice, on the SAME campaign:
  value : games finished by all ranks / device time of the campaigns alone (CUDA events on the engine's
          stream around oth_selfplay_run; weights and buffers resident), max over ranks;
  e2e   : the same campaigns through the reference-facing API with HOST buffers: weights re-folded and
          uploaded from the torch module (the trainer changes them every iteration), packed trajectories
          fetched into pinned host memory (+ NCCL weight broadcast before / trajectory all-gather into the
          replay buffer after when N > 1, timed separately in `collectives`);
  roofline : bf16 tensor-core roofline of the dominant kernel (k_net_tc), algorithmic FLOPs of the
          USEFUL leaf evaluations / summed CUDA-event time of its launches inside the timed region;
  legs  : (N = 1) short secondary campaigns: BASELINE configs 2, 3 as written (100 games), 4, both schedules,
          and the decomposition with evaluation cache and search sharing off;
  cpu_baseline : the UNMODIFIED reference (its own ParallelSelfPlayWorker / BatchMCTS / OthelloResNet /
          Cython bitboard, byte-compiled into oracle/_ref) on all host threads, bounded sample, N = 1 only.
`--impl reference` times that reference alone (rank 0), same metric / unit / config; it imports nothing
of the product package.
Warm-up steps are smaller campaigns (--warmup-games) through the same engine, kernels and e2e path, so that
`--steps 20 --warmup 5` ends within the driver's window.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  NCCL prints its version banner with printf on fd 1 (NCCL_DEBUG_FILE does not
# catch it), so the real stdout is set aside here and fd 1 is pointed at stderr for everything else in the process.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_RESULT_OUT = None


def isolate_stdout() -> None:
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "self-play games/s @50 sims/move 10x128 ResNet"
UNIT = "games/s"
MEAN_PLIES = 60.5            # measured mean game length of self-play games (reference: 60.2 random, 61 self-play)


def flops_per_position(blocks: int, F: int) -> int:
    """2 x MAC of one forward pass."""
    return 2 * (64 * 27 * F + blocks * 2 * 64 * 9 * F * F + 64 * 2 * F + 128 * 65 + 64 * F + 64 * 256 + 256)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained", 1387.2), "bf16_burst": d.get("bf16_tflops", 1660.3),
                "hbm_gbs": d.get("hbm_gbs", 6555.8), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [c for c, p in zip(sm, pw) if p > 250.0] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


class _PortArm:
    """Stand-in when oracle/_ref did not travel: the CPU port of the reference's batched self-play (oracle/selfplay_port.py:
    C tree + fp32 torch network).  It does the reference's work per game with less interpreter overhead (kind "port")."""
    kind = "port"

    def __init__(self, args, threads):
        import torch
        from oracle import net_oracle, selfplay_port
        self.args, self.threads, self.port = args, threads, selfplay_port
        torch.manual_seed(42)
        self.sd = net_oracle.make_state_dict(args.blocks, args.filters, 42)      # synthetic weights of the same shape: the cost is the same

    def step(self, budget_s):
        a = self.args
        r = self.port.measure_games_per_second(self.sd, num_simulations=a.sims, c_puct=a.c_puct, temperature_threshold=a.temp_threshold,
                                               num_parallel_games=16, time_budget_s=budget_s, threads=self.threads,
                                               mean_plies_per_game=MEAN_PLIES, seed=1)
        return r


def reference_arm(args, threads=None):
    """The UNMODIFIED reference on the host cores (oracle/ref_arm.py: its own ParallelSelfPlayWorker / BatchMCTS /
    OthelloResNet / Cython bitboard from oracle/_ref).  Nothing of the product package is imported on this path."""
    from oracle import ref_arm
    try:
        arm = ref_arm.ReferenceArm(num_blocks=args.blocks, num_filters=args.filters, num_simulations=args.sims, c_puct=args.c_puct,
                                   temperature_threshold=args.temp_threshold, num_parallel_games=16, threads=threads, seed=42)
        arm.kind = "reference"
        return arm
    except Exception as e:
        print(f"reference arm: oracle/_ref unusable ({e}); falling back to the CPU port", file=sys.stderr)
        return _PortArm(args, threads or os.cpu_count() or 1)


def cpu_playout_baseline():
    """BASELINE config 1 on the host: the reference's own Cython bitboard in benchmark.py's loop (1 thread, from
    oracle/_ref when it travelled) and the C restatement on 1 and on all threads."""
    import random as _r
    from oracle import cref, refload
    out = {}
    t0 = time.perf_counter(); cref.random_playouts(20000, 1, threads=1); dt = time.perf_counter() - t0
    out["c_restatement_1_thread_games_per_s"] = 20000 / dt
    n = 200000
    t0 = time.perf_counter(); cref.random_playouts(n, 2, threads=0); dt = time.perf_counter() - t0
    out["c_restatement_all_threads_games_per_s"] = n / dt
    out["threads"] = os.cpu_count()
    try:
        Board = refload.ref_bitboard_class()
        if Board is not None:
            def play():                                   # benchmark.py:18-40
                b = Board()
                while not b.is_terminal():
                    lm = b.get_legal_moves()
                    b.make_move(64 if lm == [64] else _r.choice(lm))
            for _ in range(50):
                play()
            t0 = time.perf_counter()
            for _ in range(1000):
                play()
            out["reference_cython_1_thread_games_per_s"] = 1000 / (time.perf_counter() - t0)
    except Exception as e:      # the compiled reference is a convenience, never a requirement
        out["reference_cython_error"] = str(e)[:120]
    return out


_MODELS = {}


def build_model(blocks, filters):
    import torch
    from othello_reinforcement_learning_test_b200.net import OthelloResNet
    key = (blocks, filters)
    if key not in _MODELS:
        torch.manual_seed(42)                       # BASELINE configs: random-init weights, seed 42
        _MODELS[key] = OthelloResNet(blocks, filters).eval()
    return _MODELS[key]


def config_dict(args, world):
    return {"workload": f"default_8x8 self-play: {args.blocks}x{args.filters} ResNet, {args.sims} sims/move, c_puct {args.c_puct:g}, "
                        f"temperature threshold {args.temp_threshold}, Dirichlet noise on; one step = one campaign of {args.games} "
                        f"concurrent games per GPU played from the start position to completion",
            "games_per_step_per_gpu": args.games, "sims_per_move": args.sims, "parallelism": f"games sharded over {world} GPU(s), no "
            "collective inside the move loop", "weights": "random init, torch.manual_seed(42)",
            "l2": "256 MiB scratch buffer written between timed steps (L2 flush)", "engine": args.engine,
            "eval_cache": not args.no_eval_cache, "search_sharing": not args.no_sharing, "schedule": args.schedule,
            "warmup": f"warm-up steps are campaigns of {min(args.warmup_games, args.games)} games through the same engine, kernels "
                      "and end-to-end path (the timed steps are full campaigns); every campaign draws its moves from its own seed"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    arm = reference_arm(args, threads)
    per_step, sample = [], ""
    for i in range(args.warmup + args.steps):
        r = arm.step(args.cpu_step_seconds)
        if i >= args.warmup:
            per_step.append(r)
        sample = r["sample"]
    plies = sum(r["plies"] for r in per_step); secs = sum(r["seconds"] for r in per_step)
    value = plies / secs / MEAN_PLIES
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * secs / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": arm.kind,
                             "sample": f"each step: {sample}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def weight_upload_bytes(nb, nf):
    """Bytes oth_net_load_weights moves host -> device for one re-fold: bf16 UMMA tiles + fp32 validation copy + biases/heads."""
    return (9 * 16 * nf + 2 * nb * 9 * nf * nf) * 2 + (9 * 8 * nf + 2 * nb * 9 * nf * nf) * 4 + 4 * (
        (1 + 2 * nb) * nf + 3 * nf + 3 + 128 * 65 + 65 + 64 * 256 + 513)


def run_leg(pkg, ctx, name, *, blocks=10, filters=128, sims=50, c_puct=1.0, thr=15, games=4096, schedule="auto", eval_cache=True,
            share=True, min_seconds=1.5, max_campaigns=4, what=""):
    """A short secondary measurement: one warm-up campaign, then campaigns until `min_seconds` of device time."""
    import torch
    model = build_model(blocks, filters)
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, torch.device("cuda", ctx.device), num_simulations=sims,
                                   temperature_threshold=thr, num_parallel_games=16, c_puct=c_puct, concurrent_games=games,
                                   seed=4242, verbose=False, eval_cache=eval_cache, share_searches=share, schedule=schedule, ctx=ctx)
    net = w.batch_mcts._native_net()
    eng = w._get_engine(games, True)
    eng.play(net.handle, games)
    # throughput: campaigns WITHOUT the per-kernel CUDA events (8 event records per network launch sit between the kernels
    # and cost a 100-game campaign ~15 %); then one more campaign with them for the kernel shares
    tot = {"ms": 0.0, "games": 0, "samples": 0, "evals": 0, "pos": 0, "hits": 0, "dups": 0, "searches": 0, "ticks": 0, "launches": 0}
    n = 0
    while n < max_campaigns and (n == 0 or tot["ms"] < 1000.0 * min_seconds):
        ns = eng.play(net.handle, games)
        st = eng.last_stats
        tot["ms"] += st["device_ms"]; tot["games"] += games; tot["samples"] += ns; tot["evals"] += eng.last_n_evals
        tot["pos"] += st["nn_positions"]; tot["hits"] += st["cache_hits"]; tot["dups"] += st["same_step_duplicates"]
        tot["searches"] += st["searches_run"]; tot["ticks"] += st["network_launches"]; tot["launches"] += st["kernel_launches"]
        n += 1
    ctx.timing_enable(True)
    eng.play(net.handle, games)
    ev_ms, ev_pos = eng.last_stats["device_ms"], eng.last_stats["nn_positions"]
    timing = ctx.timing_read()
    ctx.timing_enable(False)
    sched = eng.last_stats["schedule"]
    w._engine.close(); w._engine = None
    fpp = flops_per_position(blocks, filters)
    net_ms = timing["net"][0]
    return {"what": what, "net": f"{blocks}x{filters}", "sims": sims, "c_puct": c_puct, "temperature_threshold": thr,
            "games_per_campaign": games, "campaigns": n, "schedule": sched, "eval_cache": eval_cache, "search_sharing": share,
            "games_per_s": tot["games"] / (tot["ms"] / 1e3), "ms_per_campaign": tot["ms"] / n,
            "expansions_per_game": tot["evals"] / tot["games"], "samples_per_game": tot["samples"] / tot["games"],
            "network_positions_per_expansion": tot["pos"] / max(tot["evals"], 1),
            "searches_run_per_requested": tot["searches"] / max(tot["samples"], 1),
            "network_launches_per_campaign": tot["ticks"] / n,
            "kernel_launches_per_campaign": tot["launches"] / n,
            "network_tflops": ev_pos * fpp / (net_ms / 1e3) / 1e12 if net_ms > 0 else None,
            "network_us_per_launch": 1e3 * net_ms / max(timing["net"][1], 1),
            "network_share_of_campaign": net_ms / ev_ms if ev_ms else None,
            "tree_kernels_share_of_campaign": timing["tree"][0] / ev_ms if ev_ms else None,
            "shares_from": "one extra campaign with per-kernel CUDA events (which slow it down; games_per_s is measured without them)"}


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    t_start = time.time()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200 import dist as odist

    ctx = pkg.Context.default(local)
    # torch (NCCL included) works on the library's stream: collectives, copies and kernels are ordered without host syncs
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    torch.cuda.set_stream(stream)
    model = build_model(args.blocks, args.filters)
    if world > 1:
        odist.broadcast_weights(model, src=0)
    worker = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, dev, num_simulations=args.sims,
                                        temperature_threshold=args.temp_threshold, num_parallel_games=16, c_puct=args.c_puct, dirichlet_alpha=0.3,
                                        dirichlet_epsilon=0.25, concurrent_games=args.games, engine=args.engine,
                                        seed=1000 + rank, verbose=False, eval_cache=not args.no_eval_cache,
                                        share_searches=not args.no_sharing, schedule=args.schedule, ctx=ctx)
    G = args.games
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    net = worker.batch_mcts._native_net()
    engine = worker._get_engine(G, True)
    engine._pinned_out(G * 66)          # page-locked result buffer allocated once, outside the timed region (setup, like cudaMalloc)
    replay = pkg.ReplayBuffer(max_size=int(G * 64 * world), ctx=ctx) if world > 1 else None
    gatherer = odist.DeviceSampleGather(dev) if world > 1 else None
    if gatherer is not None:
        gatherer.reserve(G * 64)        # receive buffer allocated once, outside the timed region
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1, b0, b1, g0, g1, s0 = ev(), ev(), ev(), ev(), ev(), ev(), ev()
    acc = {"e2e_ms": 0.0, "dev_ms": 0.0, "bcast_ms": 0.0, "gather_ms": 0.0, "skew_ms": 0.0, "samples": 0, "evals": 0, "h2d": 0, "d2h": 0,
           "pos": 0, "hits": 0, "dups": 0, "coll": 0, "searches": 0, "ticks": 0, "gather_bytes": 0}

    def step(episodes, timed):
        """One campaign the way the trainer's iteration sees it: new weights in, packed trajectories out (host buffers)."""
        flush_l2()
        e0.record(stream)
        if world > 1:
            b0.record(stream)
            nb_bytes = odist.broadcast_weights(model, src=0)             # NCCL broadcast of the new weights (one flat buffer)
            b1.record(stream)
        net.sync_from(model, force=True)                                 # fold BN + bf16 pack + H2D
        smp = worker.execute_episodes_packed(episodes, add_dirichlet_noise=True, reuse_buffer=True)   # campaign + D2H into pinned host memory
        if world > 1:                                                    # trajectories of every rank into the replay buffer,
            s0.record(stream)                                            # NCCL all-gather device to device.  Ranks play different
            dist.barrier()                                               # games and finish at different times: that wait is timed
            g0.record(stream)                                            # apart from the collective itself
            dptr, cnt = worker._engine.samples_device()
            segments = gatherer.gather(dptr, cnt, G * 128, episodes=episodes)     # straight from the engine's buffer
            replay.clear()
            total_cnt = 0
            for seg, c in segments:
                replay.add_device(seg, c)
                total_cnt += c
            g1.record(stream)
        e1.record(stream)
        e1.synchronize()
        if not timed:
            return
        st = worker.last_stats
        acc["e2e_ms"] += e0.elapsed_time(e1); acc["dev_ms"] += st["device_ms"]
        acc["samples"] += int(smp.size); acc["evals"] += int(st["nn_evals"])
        acc["pos"] += st["nn_positions"]; acc["hits"] += st["cache_hits"]; acc["dups"] += st["same_step_duplicates"]
        acc["coll"] += st["hash_collisions"]; acc["searches"] += st["searches_run"]; acc["ticks"] += st["network_launches"]
        acc["h2d"] += weight_upload_bytes(args.blocks, args.filters)
        acc["d2h"] += int(smp.size) * 168 + 64 * 130
        if world > 1:
            acc["bcast_ms"] += b0.elapsed_time(b1); acc["gather_ms"] += g0.elapsed_time(g1); acc["skew_ms"] += s0.elapsed_time(g0)
            acc["gather_bytes"] += int(total_cnt) * 168

    for _ in range(args.warmup):
        step(min(args.warmup_games, G), False)

    # ---------------- timed: K campaigns; device interval and end-to-end interval of the SAME campaigns ----------------
    sampler = ClockSampler(local)
    barrier()
    ctx.timing_enable(True)
    sampler.start()
    launches0 = ctx.launch_count
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        step(G, True)
    barrier()
    wall_s = time.perf_counter() - t_wall
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    timing = ctx.timing_read()
    ctx.timing_enable(False)
    ms = torch.tensor([acc["dev_ms"], acc["e2e_ms"], acc["bcast_ms"], acc["gather_ms"], acc["skew_ms"]], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(acc["samples"]), float(acc["evals"]), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    max_ms, max_e2e_ms = float(ms[0].item()), float(ms[1].item())
    games_total = G * args.steps * world
    value = games_total / (max_ms / 1000.0)
    e2e_value = games_total / (max_e2e_ms / 1000.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = measured_peaks()
    fpp = flops_per_position(args.blocks, args.filters)
    net_ms, net_launches = timing["net"]
    useful_evals = float(acc["pos"])           # positions the network really evaluated on this rank (compacted, de-duplicated)
    achieved = useful_evals * fpp / (net_ms / 1000.0) / 1e12 if net_ms > 0 else 0.0
    traffic, traffic_src = None, None          # dram read+write bytes per launch of the kernel, from a committed ncu --set full capture
    for prof in ("r02_prof_net_tc_fullbatch.txt", "r01_prof_net_tc_fullbatch.txt"):
        try:
            rd = wr = None
            for ln in open(os.path.join(ROOT, "profiles", prof)):
                if "dram__bytes_read.sum " in ln and "Mbyte" in ln:
                    rd = float(ln.split()[2]) * 1e6 if ln.split()[1] == "Mbyte" else float(ln.split()[1]) * 1e6
                if "dram__bytes_write.sum " in ln and ("Mbyte" in ln or "Kbyte" in ln):
                    f = ln.split(); unit = 1e6 if "Mbyte" in ln else 1e3
                    wr = (float(f[2]) if f[1] in ("Mbyte", "Kbyte") else float(f[1])) * unit
            if rd is not None:
                traffic = int(rd + (wr or 0.0)); traffic_src = f"profiles/{prof} (ncu --set full of one 18,944-position launch; not measured in this run)"
                break
        except (OSError, ValueError, IndexError):
            continue
    roof = {"bound": "tensor", "kernel": "k_net_tc" if args.engine != "simt" else "k_net_simt", "achieved": achieved,
            "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peaks["source"] + ", bf16 sustained (kernel timed inside a long step)",
            "flop_per_position": fpp, "useful_positions": int(useful_evals), "kernel_ms_total": net_ms, "kernel_launches": int(net_launches),
            "kernel_share_of_step": net_ms / acc["dev_ms"] if acc["dev_ms"] else None,
            "tree_kernels_ms_total": timing["tree"][0], "move_kernels_ms_total": timing["move"][0]}

    # ---- secondary rooflines: tree kernels and random playouts against the HBM roofline ----
    tree_ms = timing["tree"][0]
    sims_run = float(acc["searches"]) * (args.sims + 1)    # simulations that really walked a tree (shared searches run once)
    tree_bytes = sims_run * 750.0                          # ~0.75 KB of tree traffic per simulation (select + expand + backup)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config_dict(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": acc["h2d"] // args.steps, "d2h_bytes_per_step": acc["d2h"] // args.steps,
                    "ms_per_step": max_e2e_ms / args.steps,
                    "what": "ParallelSelfPlayWorker.execute_episodes_packed with host buffers: weights re-folded and uploaded from the torch "
                            "module, campaign, packed records fetched into pinned host memory" + (
                                "; plus NCCL weight broadcast before and trajectory all-gather into the replay buffer after" if world > 1 else "")},
            "gpu_launches": int(tot[2].item()), "clocks": clocks, "roofline": roof,
            "expansions_per_game": float(tot[1].item()) / games_total, "samples_per_game": float(tot[0].item()) / games_total,
            "schedule": worker.last_stats.get("schedule"),
            "eval_cache": {"enabled": not args.no_eval_cache, "rank0_expansions": int(acc["evals"]), "rank0_network_positions": int(acc["pos"]),
                           "rank0_cache_hits": int(acc["hits"]), "rank0_same_step_duplicates": int(acc["dups"]),
                           "rank0_hash_collisions": int(acc["coll"]), "rank0_searches_run": int(acc["searches"]),
                           "rank0_searches_requested": int(acc["samples"]), "rank0_network_launches": int(acc["ticks"]),
                           "network_positions_per_expansion": acc["pos"] / max(acc["evals"], 1),
                           "note": "result-transparent: cached / shared outputs are bit-identical to re-evaluation; the cache is emptied at "
                                   "the start of every campaign.  Sharing exists because all games of a campaign start together: see "
                                   "legs.decomposition for the same engine with cache and sharing off"},
            "wall_s_timed_region": wall_s}
    if world > 1:
        line["collectives"] = {"broadcast_ms_per_step": float(ms[2].item()) / args.steps, "all_gather_ms_per_step": float(ms[3].item()) / args.steps,
                               "all_gather_bytes_per_rank_per_step": acc["gather_bytes"] // args.steps,
                               "rank_skew_wait_ms_per_step": float(ms[4].item()) / args.steps,
                               "share_of_e2e_step": (float(ms[2].item()) + float(ms[3].item())) / max_e2e_ms,
                               "what": "max over ranks; broadcast = one flat fp32 buffer of every parameter and buffer (NCCL), all-gather = "
                                       "168-byte records device to device (NCCL, straight from the engine's buffer) + copy into the device "
                                       "replay ring; rank_skew_wait = the barrier before the all-gather (ranks play different games and "
                                       "finish at different times; the longest wait is the fastest rank's)"}
    secondary = {"tree_kernels": {"bound": "hbm (latency in practice)", "algorithmic_bytes": tree_bytes, "ms": tree_ms,
                                  "simulations_run": sims_run,
                                  "achieved": tree_bytes / (tree_ms / 1e3) / 1e9 if tree_ms else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                  "frac": (tree_bytes / (tree_ms / 1e3) / 1e9 / peaks["hbm_gbs"]) if tree_ms else None,
                                  "note": "on EXECUTED simulations: searches_run x (sims + 1) x 750 B; pointer chasing, one dependent 24-byte-record "
                                          "load per tree level -- bound by load latency, not bandwidth (profiles/r02_prof_tree.txt)"}}
    line["secondary_rooflines"] = secondary

    if world == 1:
        # release the big engine before the small legs
        worker._engine.close(); worker._engine = None
        del flush
        torch.cuda.empty_cache()
        # ---------------- secondary metric: random playouts (BASELINE config 1) ----------------
        from othello_reinforcement_learning_test_b200 import bitboard as bb
        n_po = 1 << 24
        bb.random_playouts(1 << 20, seed=1, ctx=ctx)
        e0.record(stream)
        po = bb.random_playouts(n_po, seed=2, ctx=ctx)
        e1.record(stream); e1.synchronize()
        po_ms = e0.elapsed_time(e1)
        playout_bytes = n_po * 20.0                        # 16 B final board + 4 B ply count per game; state lives in registers
        secondary["random_playout"] = {"bound": "int-pipe (nominally hbm)", "algorithmic_bytes": playout_bytes, "ms": po_ms,
                                       "achieved": playout_bytes / (po_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                       "frac": playout_bytes / (po_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                                       "plies_per_s": po["total_plies"] / (po_ms / 1e3),
                                       "note": "whole game in registers: ~36k integer ops per game, 20 B of HBM traffic; the bound is the INT pipe"}
        line["random_playout"] = {"games_per_s": n_po / (po_ms / 1000.0), "games": n_po, "mean_plies": po["total_plies"] / n_po,
                                  "note": "BASELINE config 1, one game per thread in registers"}
        if not args.no_legs:
            line["legs"] = run_legs(args, pkg, ctx, t_start)
            rs = line["legs"].get("reference_signature_e2e_4096_games", {})
            if "reference_signature" in rs:       # the list-of-tuples route next to the packed one, where a reader looks for e2e
                line["e2e"]["reference_signature_4096_games"] = {
                    "list_api_games_per_s": rs["reference_signature"]["games_per_s"], "packed_api_games_per_s": rs["packed"]["games_per_s"],
                    "what": "host wall clock of execute_episodes() -> ReplayBuffer.add(list) -> sample(256) vs execute_episodes_packed() -> "
                            "add_from_worker -> sample_torch(256) on a 4,096-game campaign (legs.reference_signature_e2e_4096_games)"}
        if not args.no_cpu_baseline:
            try:
                arm = reference_arm(args)
                r = arm.step(args.cpu_seconds)
                line["cpu_baseline"] = {"value": r["games_per_s"], "unit": UNIT, "cores": r["threads"], "kind": arm.kind, "sample": r["sample"]}
                line["random_playout"]["cpu"] = cpu_playout_baseline()
            except Exception as e:          # the baseline must never take the measured line down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {str(e)[:160]}"}
        else:
            line["cpu_baseline"] = None
    else:
        line["cpu_baseline"] = None
    line["wall_s_total"] = time.time() - t_start
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_legs(args, pkg, ctx, t_start):
    """Short secondary measurements on one GPU (N = 1 only): every BASELINE config, both schedules at small campaigns, and the
    decomposition of the headline (cache and search sharing off).  Each leg is a few seconds; legs are skipped, and say so,
    once the run is older than --leg-deadline seconds."""
    strong = dict(sims=100, c_puct=1.5, thr=20)
    plan = [
        ("config3_literal_100_games", dict(games=100, what="default_8x8.yaml as written: 100 games per iteration, 10x128, 50 sims")),
        ("config3_literal_100_games_lockstep", dict(games=100, schedule="lockstep", what="same, lock-step schedule (round-1 engine) for comparison")),
        ("config3_4096_games", dict(games=4096, what="10x128, 50 sims, 4,096 concurrent games")),
        ("config3_4096_games_lockstep", dict(games=4096, schedule="lockstep", what="same, lock-step schedule for comparison")),
        ("config4_strong_4096_games", dict(games=4096, what="strong_8x8.yaml: 100 sims, c_puct 1.5, threshold 20, 4,096 concurrent games", **strong)),
        ("config4_strong_37888_games", dict(games=37888, max_campaigns=1, min_seconds=0.0,
                                            what="strong_8x8.yaml search settings, 37,888 concurrent games (one campaign)", **strong)),
        ("config2_debug_5x64_100_games", dict(blocks=5, filters=64, games=100, what="debug_6x6.yaml: 5x64 net, 50 sims, 100 games")),
        ("config2_debug_5x64_4096_games", dict(blocks=5, filters=64, games=4096, what="debug_6x6.yaml net, 4,096 concurrent games")),
        ("decomposition_no_cache_no_sharing_4736_games", dict(games=4736, eval_cache=False, share=False, schedule="lockstep", max_campaigns=1,
                                                             min_seconds=0.0,
                                                             what="10x128, 50 sims with the evaluation cache AND search sharing off: every expansion is a "
                                                                  "network evaluation (kernel speed without result-transparent sharing; 4,736 games = 8 "
                                                                  "positions per SM per launch, enough to fill the tensor pipes)")),
    ]
    out = {}
    for name, kw in plan:
        if time.time() - t_start > args.leg_deadline:
            out[name] = {"skipped": f"run older than --leg-deadline {args.leg_deadline:.0f}s"}
            continue
        try:
            out[name] = run_leg(pkg, ctx, name, **kw)
        except Exception as e:          # a leg must never take the headline down with it
            out[name] = {"error": str(e)[:200]}
    try:
        out["single_board_path"] = leg_single_board(pkg, ctx)
    except Exception as e:
        out["single_board_path"] = {"error": str(e)[:200]}
    if time.time() - t_start <= args.leg_deadline:
        try:
            out["reference_signature_e2e_4096_games"] = leg_reference_signature(args, pkg, ctx)
        except Exception as e:
            out["reference_signature_e2e_4096_games"] = {"error": str(e)[:200]}
    return out


def leg_single_board(pkg, ctx, games=20):
    """The reference's per-move board calls (arena.py:106-119: is_terminal, get_legal_moves, make_move, then get_winner /
    get_stone_counts at the end) on OthelloBitboard: one oth_board_step launch per move, everything else from its result.
    Next to it the compiled reference board (Cython, oracle/_ref) doing the same loop on this host."""
    def loop(Board):
        n = 0
        t0 = time.perf_counter()
        for g in range(games):
            b = Board()
            while not b.is_terminal():
                lm = b.get_legal_moves()
                b.make_move(lm[(n + g) % len(lm)])
                n += 1
            b.get_winner(); b.get_stone_counts()
        return 1e6 * (time.perf_counter() - t0) / n, n
    loop(pkg.OthelloBitboard)
    l0 = ctx.launch_count
    us, moves = loop(pkg.OthelloBitboard)
    out = {"us_per_move": us, "moves": moves, "launches_per_move": (ctx.launch_count - l0) / moves,
           "what": "is_terminal + get_legal_moves + make_move per move through the reference-shaped class; one one-warp launch, result in a "
                   "mapped page-locked mailbox (no allocation, no memcpy)"}
    b = pkg.OthelloBitboard()
    t0 = time.perf_counter()
    for i in range(2000):
        b.self_board = b.self_board            # drops the cached answers: every query is a launch
        b.get_legal_moves_bits()
    out["us_per_uncached_query"] = 1e6 * (time.perf_counter() - t0) / 2000
    try:
        from oracle import refload
        Ref = refload.ref_bitboard_class()
        if Ref is not None:
            out["reference_cython_us_per_move"] = loop(Ref)[0]
    except Exception as e:
        out["reference_cython_error"] = str(e)[:120]
    return out


def leg_reference_signature(args, pkg, ctx, games=4096):
    """trainer.py:180-185,264-269 as the reference's trainer calls them: execute_episodes() -> list[(f32[3,8,8], f32[65], float)]
    -> ReplayBuffer.add(list) -> sample(256), next to the packed path (execute_episodes_packed -> add_from_worker -> sample_torch)."""
    import torch
    model = build_model(args.blocks, args.filters)
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, torch.device("cuda", ctx.device), num_simulations=args.sims,
                                   temperature_threshold=args.temp_threshold, num_parallel_games=16, c_puct=args.c_puct,
                                   concurrent_games=games, seed=77, verbose=False, ctx=ctx)
    w.execute_episodes_packed(games)                                     # warm-up
    buf = pkg.ReplayBuffer(max_size=games * 70, ctx=ctx)
    t0 = time.perf_counter()
    data = w.execute_episodes(num_episodes=games, add_dirichlet_noise=True)
    t1 = time.perf_counter()
    buf.add(data)
    t2 = time.perf_counter()
    st, po, va = buf.sample(256)
    t3 = time.perf_counter()
    n_samples = len(data)
    del data
    buf.clear()
    t4 = time.perf_counter()
    w.execute_episodes_packed(games, reuse_buffer=True)
    t5 = time.perf_counter()
    buf.add_from_worker(w)
    ctx.sync()
    t6 = time.perf_counter()
    buf.sample_torch(256)
    torch.cuda.synchronize()
    t7 = time.perf_counter()
    w._engine.close(); w._engine = None
    return {"games": games, "samples": n_samples,
            "reference_signature": {"games_per_s": games / (t3 - t0), "execute_episodes_s": t1 - t0, "replay_add_list_s": t2 - t1, "sample_256_s": t3 - t2,
                                    "what": "execute_episodes() -> list of (f32[3,8,8], f32[65], float) tuples -> ReplayBuffer.add(list) -> sample(256) -> numpy"},
            "packed": {"games_per_s": games / (t7 - t4), "execute_episodes_packed_s": t5 - t4, "add_from_worker_s": t6 - t5, "sample_torch_256_s": t7 - t6,
                       "what": "execute_episodes_packed() -> ReplayBuffer.add_from_worker (device to device) -> sample_torch(256) -> CUDA tensors"},
            "note": "host wall clock; the difference is the expansion of 168-byte records into float planes and Python tuples, "
                    "which the trainer can skip (INTEGRATION.md)"}


def main():
    isolate_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=303104, help="concurrent games per GPU = games per step per GPU (2048 x 148 SMs)")
    ap.add_argument("--sims", type=int, default=50)
    ap.add_argument("--c-puct", type=float, default=1.0, help="default_8x8.yaml: 1.0; strong_8x8.yaml: 1.5")
    ap.add_argument("--temp-threshold", type=int, default=15, help="default_8x8.yaml: 15; strong_8x8.yaml: 20")
    ap.add_argument("--blocks", type=int, default=10)
    ap.add_argument("--filters", type=int, default=128)
    ap.add_argument("--engine", default="tcgen05", choices=["tcgen05", "tcgen05_pair", "simt"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="time budget of the cpu_baseline sample")
    ap.add_argument("--cpu-step-seconds", type=float, default=8.0, help="--impl reference: time budget per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval-cache", action="store_true", help="evaluate every leaf with the network (no position cache / dedup)")
    ap.add_argument("--no-sharing", action="store_true", help="every slot runs its own search (identical roots do not share one)")
    ap.add_argument("--schedule", default="auto", choices=["auto", "lockstep", "async"],
                    help="lockstep: 1 + sims network launches per ply; async: run-until-miss; auto: async up to 8,192 slots")
    ap.add_argument("--warmup-games", type=int, default=18944, help="games per warm-up campaign (the timed steps are full campaigns)")
    ap.add_argument("--no-legs", action="store_true", help="skip the secondary legs (BASELINE configs 2/3-literal/4, decomposition)")
    ap.add_argument("--leg-deadline", type=float, default=600.0, help="secondary legs are skipped once the run is older than this (s)")
    args = ap.parse_args()
    if args.warmup < 3:
        print(f"note: --warmup {args.warmup} < 3 (timing hygiene wants >= 3)", file=sys.stderr)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
