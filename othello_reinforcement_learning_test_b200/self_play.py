"""Self-play workers with the reference's signatures, backed by the device-resident engine.

* `ParallelSelfPlayWorker` / `create_parallel_self_play_worker` mirror
  `src/train/parallel_self_play.py:217-434`: `execute_episodes(num_episodes,
  add_dirichlet_noise) -> list[(float32[3,8,8], float32[65], float)]`.  One call = one GPU
  campaign (csrc/selfplay.cu): boards, trees, leaf evaluation, move choice, trajectories and
  labels all stay in HBM; only the packed samples (168 B each) come back.
* `SelfPlayWorker` mirrors `src/train/self_play.py:27-164` (the serial worker: it stores the
  one-hot policy after the temperature threshold and samples moves with numpy's RNG on the
  host); its episodes advance in lock-step through `mcts.search_batch`.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import List, Tuple

import numpy as np

from . import _lib
from ._lib import MEM_HOST, Context, SelfPlayConfig, check, ptr
from .mcts import BatchMCTS, MCTS

Sample = Tuple[np.ndarray, np.ndarray, float]

_BIT_ORDER = "little"


def planes_from_bits(self_b: np.ndarray, opp_b: np.ndarray, legal: np.ndarray) -> np.ndarray:
    """(self, opp, legal) words -> float32 [n,3,8,8] (the layout of get_tensor_input).  Pure format
    expansion of values the GPU produced; no rules are evaluated here."""
    words = np.stack([self_b, opp_b, legal], axis=1).astype("<u8", copy=False)          # [n,3]
    bits = np.unpackbits(words.view(np.uint8).reshape(-1, 3, 8), axis=2, bitorder=_BIT_ORDER)
    return bits.reshape(-1, 3, 8, 8).astype(np.float32)


def samples_to_arrays(samples: np.ndarray, sort: bool = True):
    """Packed oth_sample records -> (states f32[n,3,8,8], policies f32[n,65], values f64[n]) in (episode, ply)
    order: the three arrays the reference's list of tuples is made of, built with a handful of vectorised numpy
    calls (no per-sample Python)."""
    if sort and samples.size:
        samples = samples[np.lexsort((samples["ply"], samples["game"]))]
    states = planes_from_bits(samples["self_b"], samples["opp_b"], samples["legal"])
    counts = samples["visits"].astype(np.float32)
    totals = counts.sum(axis=1, keepdims=True, dtype=np.float32)
    policies = np.divide(counts, totals, out=np.zeros_like(counts), where=totals > 0)   # node.py:177-180
    return states, policies, samples["value"].astype(np.float64)


def samples_to_training_data(samples: np.ndarray) -> List[Sample]:
    """Packed oth_sample records -> the reference's list of (state, policy, value), ordered by
    (episode, ply) like parallel_self_play.py:399-405.  The tuples hold views into three big arrays
    (`zip` over their rows); the only per-sample Python work is creating the tuple itself."""
    if samples.size == 0:
        return []
    states, policies, values = samples_to_arrays(samples)
    return list(zip(states, policies, values.tolist()))


class SelfPlayEngine:
    """Owner of an oth_selfplay handle."""

    def __init__(self, ctx: Context, num_simulations: int, temperature_threshold: int, concurrent_games: int,
                 c_puct: float, dirichlet_alpha: float, dirichlet_epsilon: float, add_dirichlet_noise: bool,
                 flags: int, seed: int, schedule: int = _lib.SCHEDULE_AUTO):
        self.ctx = ctx
        self.cfg = SelfPlayConfig(int(num_simulations), int(temperature_threshold), int(bool(add_dirichlet_noise)),
                                  int(concurrent_games), float(c_puct), float(dirichlet_alpha), float(dirichlet_epsilon),
                                  int(flags), int(schedule), int(seed) & (2**64 - 1))
        h = C.c_void_p()
        check(ctx.lib.oth_selfplay_create(ctx.handle, C.byref(self.cfg), C.byref(h)))
        self.handle = h
        self.last_n_evals = 0

    def _pinned_out(self, n: int) -> np.ndarray:
        """Grow-only page-locked host buffer for the campaign's records (device->host at PCIe speed, no page faults)."""
        import torch
        need = max(int(n), 1) * _lib.SAMPLE_DTYPE.itemsize
        buf = getattr(self, "_pinned", None)
        if buf is None or buf.numel() < need:
            self._pinned = buf = torch.empty(int(need * 1.1) + 4096, dtype=torch.uint8, pin_memory=True)
        return buf.numpy()[:int(n) * _lib.SAMPLE_DTYPE.itemsize].view(_lib.SAMPLE_DTYPE)

    def set_seed(self, seed: int) -> None:
        """New move-sampling seed for the next campaign (the handle and its buffers stay)."""
        check(self.ctx.lib.oth_selfplay_set_seed(self.handle, int(seed) & (2**64 - 1)))

    def play(self, net_handle, num_episodes: int) -> int:
        """One campaign on the device; the records stay there.  Returns the number of samples."""
        ns, ne = C.c_int64(0), C.c_int64(0)
        check(self.ctx.lib.oth_selfplay_run(self.handle, net_handle, int(num_episodes), C.byref(ns), C.byref(ne)))
        self.last_n_evals = int(ne.value)
        st = (C.c_uint64 * 5)()
        check(self.ctx.lib.oth_selfplay_stats(self.handle, st))
        tm = (C.c_double * 4)()
        check(self.ctx.lib.oth_selfplay_timing(self.handle, tm))
        self.last_stats = {"nn_positions": int(st[0]), "cache_hits": int(st[1]), "same_step_duplicates": int(st[2]),
                           "hash_collisions": int(st[3]), "searches_run": int(st[4]), "device_ms": float(tm[0]),
                           "network_launches": int(tm[1]), "schedule": {1: "lockstep", 2: "async"}.get(int(tm[2]), "?"),
                           "kernel_launches": int(tm[3])}
        self.last_n_samples = int(ns.value)
        return self.last_n_samples

    def fetch(self, reuse_buffer: bool = False) -> np.ndarray:
        """Device -> host copy of the last campaign's records (reuse_buffer=True: into the engine's pinned buffer)."""
        n = self.last_n_samples
        out = self._pinned_out(n) if reuse_buffer else np.empty(n, _lib.SAMPLE_DTYPE)
        check(self.ctx.lib.oth_selfplay_fetch(self.handle, ptr(out), out.size, MEM_HOST))
        return out

    def run(self, net_handle, num_episodes: int, reuse_buffer: bool = False) -> np.ndarray:
        """reuse_buffer=True returns a view into the engine's pinned buffer, valid until the next run()."""
        self.play(net_handle, num_episodes)
        return self.fetch(reuse_buffer)

    def samples_device(self):
        p, n = C.c_void_p(), C.c_int64(0)
        check(self.ctx.lib.oth_selfplay_samples_device(self.handle, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.oth_selfplay_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ParallelSelfPlayWorker:
    """Drop-in for `ParallelSelfPlayWorker` (parallel_self_play.py:217-407)."""

    DEFAULT_CONCURRENT_GAMES = 4096

    def __init__(self, board_class, model, device, num_simulations: int = 25, temperature_threshold: int = 15,
                 num_parallel_games: int = 8, c_puct: float = 1.0, dirichlet_alpha: float = 0.3,
                 dirichlet_epsilon: float = 0.25, *, concurrent_games: int | None = None, evaluator: str = "auto",
                 winner_black: bool = False, root_n_sum: bool = False, q_canonical: bool = False,
                 engine: str | None = None, seed: int | None = None, verbose: bool = True, eval_cache: bool = True,
                 share_searches: bool = True, schedule: str = "auto", ctx: Context | None = None):
        self.board_class = board_class
        self.num_simulations = num_simulations
        self.temperature_threshold = temperature_threshold
        self.num_parallel_games = num_parallel_games
        # `num_parallel_games` was sized for a 6 GB laptop GPU; games are independent, so the engine
        # keeps as many slots as the campaign can fill (capped) unless told otherwise.
        self.concurrent_games = concurrent_games
        self.batch_mcts = BatchMCTS(model=model, device=device, c_puct=c_puct, dirichlet_alpha=dirichlet_alpha,
                                    dirichlet_epsilon=dirichlet_epsilon, evaluator=evaluator, root_n_sum=root_n_sum,
                                    q_canonical=q_canonical, engine=engine, eval_cache=eval_cache, ctx=ctx)
        self.winner_black = winner_black
        self.share_searches = share_searches      # identical root positions run one search (same results)
        # "lockstep": 1 + sims network launches per ply; "async": run-until-miss (one launch per cache miss of the
        # slowest slot); "auto": async up to 8,192 slots when the cache is on.  Same records either way.
        self.schedule = {"auto": _lib.SCHEDULE_AUTO, "lockstep": _lib.SCHEDULE_LOCKSTEP, "async": _lib.SCHEDULE_ASYNC}[schedule]
        self.seed = seed
        self.verbose = verbose
        self._engine: SelfPlayEngine | None = None
        self._engine_key = None
        self.last_stats: dict = {}

    def _slots_for(self, num_episodes: int) -> int:
        # an explicit `concurrent_games` is the engine's size, whatever a campaign asks for (a shorter campaign leaves
        # slots idle; the trees, staging, cache and pinned buffer are allocated once)
        if self.concurrent_games:
            return max(1, int(self.concurrent_games))
        want = max(self.num_parallel_games, min(num_episodes, self.DEFAULT_CONCURRENT_GAMES))
        return max(1, min(int(want), max(num_episodes, 1)))

    def _get_engine(self, num_episodes: int, add_noise: bool) -> SelfPlayEngine:
        m = self.batch_mcts
        flags = m._flags() | (_lib.FLAG_WINNER_BLACK if self.winner_black else 0) | (
            0 if self.share_searches else _lib.FLAG_NO_SEARCH_SHARING)
        # The seed is NOT part of the key: trees, staging, the cache and the pinned buffer are kept from campaign to
        # campaign.  seed=None draws one from numpy when the engine is made (the reference consumes np.random too);
        # every later campaign on the handle derives its own stream from it, so no two campaigns replay the same draws.
        key = (self._slots_for(num_episodes), self.num_simulations, self.temperature_threshold, m.c_puct,
               m.dirichlet_alpha, m.dirichlet_epsilon, bool(add_noise), flags, self.schedule)
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            seed = self.seed if self.seed is not None else int(np.random.randint(0, 2**31 - 1))
            self._engine = SelfPlayEngine(m._context(), self.num_simulations, self.temperature_threshold, key[0],
                                          m.c_puct, m.dirichlet_alpha, m.dirichlet_epsilon, add_noise, flags, seed, self.schedule)
            self._engine_key = key
        return self._engine

    def execute_episodes_packed(self, num_episodes: int, add_dirichlet_noise: bool = True, reuse_buffer: bool = False) -> np.ndarray:
        """Same campaign, packed records (structured array of oth_sample).  reuse_buffer=True: the result is a view
        into a page-locked buffer owned by the worker (no 3 GB pageable allocation per campaign), valid until the next call."""
        m = self.batch_mcts
        if m.evaluator == "external":
            raise _lib.OthelloB200Error(
                "the device-resident campaign needs an OthelloResNet-shaped model (or the hash-net test evaluator); "
                "use SelfPlayWorker for arbitrary callables")
        t0 = time.time()
        eng = self._get_engine(num_episodes, add_dirichlet_noise)
        net = m._native_net().handle if m.evaluator == "native" else None
        samples = eng.run(net, num_episodes, reuse_buffer)
        dt = time.time() - t0
        self.last_stats = {"episodes": num_episodes, "samples": int(samples.size), "seconds": dt,
                           "nn_evals": eng.last_n_evals, "concurrent_games": eng.cfg.concurrent_games, **eng.last_stats}
        return samples

    def execute_episodes(self, num_episodes: int, add_dirichlet_noise: bool = True) -> List[Sample]:
        """parallel_self_play.py:282-322"""
        samples = self.execute_episodes_packed(num_episodes, add_dirichlet_noise)
        data = samples_to_training_data(samples)
        if self.verbose:
            s = self.last_stats
            print(f"  Self-Play: {num_episodes}/{num_episodes} (100.0%) | {len(data):,} samples | "
                  f"{s['seconds']:.1f}s | {num_episodes / max(s['seconds'], 1e-9):.1f} games/s")
        return data

    def _execute_batch(self, batch_size: int, add_dirichlet_noise: bool) -> List[Sample]:
        """parallel_self_play.py:324-407 (one lock-step batch)."""
        return samples_to_training_data(self.execute_episodes_packed(batch_size, add_dirichlet_noise))


def create_parallel_self_play_worker(config: dict, model, device, **kwargs) -> ParallelSelfPlayWorker:
    """parallel_self_play.py:410-434: same YAML keys, same defaults."""
    from .bitboard import OthelloBitboard
    mc, sp = config.get("mcts", {}), config.get("self_play", {})
    return ParallelSelfPlayWorker(
        board_class=OthelloBitboard, model=model, device=device,
        num_simulations=mc.get("num_simulations", 25),
        temperature_threshold=sp.get("temperature_threshold", 15),
        num_parallel_games=sp.get("num_parallel_games", 8),
        c_puct=mc.get("c_puct", 1.0), dirichlet_alpha=mc.get("dirichlet_alpha", 0.3),
        dirichlet_epsilon=mc.get("dirichlet_epsilon", 0.25), **kwargs)


class SelfPlayWorker:
    """Drop-in for `SelfPlayWorker` (self_play.py:27-164)."""

    def __init__(self, board_class, mcts, num_simulations: int = 25, temperature_threshold: int = 15):
        self.board_class = board_class
        self.mcts = mcts
        self.num_simulations = num_simulations
        self.temperature_threshold = temperature_threshold

    def _play(self, n_games: int, add_dirichlet_noise: bool) -> List[List[Sample]]:
        boards = []
        for _ in range(n_games):
            b = self.board_class(); b.reset(); boards.append(b)
        hist: List[list] = [[] for _ in range(n_games)]
        moves = [0] * n_games
        live = [i for i in range(n_games) if not boards[i].is_terminal()]
        batched = hasattr(self.mcts, "search_batch")
        while live:
            # the stored policy uses the move's own temperature (one-hot after the threshold), self_play.py:87-98
            temps = [1.0 if moves[i] < self.temperature_threshold else 0.0 for i in live]
            results = {}
            for T in sorted(set(temps)):
                idx = [i for i, t in zip(live, temps) if t == T]
                if batched:
                    res = self.mcts.search_batch([boards[i] for i in idx], self.num_simulations, T, add_dirichlet_noise)
                else:
                    res = [self.mcts.search(boards[i], num_simulations=self.num_simulations, temperature=T,
                                            add_dirichlet_noise=add_dirichlet_noise) for i in idx]
                results.update(dict(zip(idx, res)))
            nxt = []
            for i, T in zip(live, temps):
                policy = results[i][0]
                player = 1 if moves[i] % 2 == 0 else -1                          # self_play.py:83
                hist[i].append((boards[i].get_tensor_input().copy(), policy.copy(), player))
                action = int(np.argmax(policy)) if T == 0 else int(np.random.choice(len(policy), p=policy))
                boards[i].make_move(action)
                moves[i] += 1
                if not boards[i].is_terminal():
                    nxt.append(i)
            live = nxt
        out = []
        for i in range(n_games):
            winner = boards[i].get_winner()                                      # self_play.py:120
            out.append([(s, p, float(winner * pl)) for s, p, pl in hist[i]])
        return out

    def execute_episode(self, add_dirichlet_noise: bool = True) -> List[Sample]:
        """self_play.py:52-135"""
        return self._play(1, add_dirichlet_noise)[0]

    def execute_episodes(self, num_episodes: int, add_dirichlet_noise: bool = True) -> List[Sample]:
        """self_play.py:137-164 (episodes advance together; each is an independent game)."""
        data: List[Sample] = []
        for ep in self._play(num_episodes, add_dirichlet_noise):
            data.extend(ep)
        return data


def augment_data_with_symmetries(training_data: List[Sample], board_class, full: bool = False) -> List[Sample]:
    """self_play.py:166-212.  The reference's function is a stub that returns the input unchanged;
    that is the default here too.  `full=True` returns the 8 dihedral images of every sample."""
    if not full:
        return list(training_data)
    out: List[Sample] = []
    for state, policy, value in training_data:
        grid = np.asarray(policy)[:64].reshape(8, 8)
        for k in range(4):
            rs = np.rot90(state, k, axes=(1, 2)); rg = np.rot90(grid, k)
            for flip in (False, True):
                s2 = np.flip(rs, axis=2) if flip else rs
                g2 = np.flip(rg, axis=1) if flip else rg
                p2 = np.zeros(65, np.float32); p2[:64] = g2.reshape(-1); p2[64] = policy[64]
                out.append((np.ascontiguousarray(s2), p2, value))
    return out
