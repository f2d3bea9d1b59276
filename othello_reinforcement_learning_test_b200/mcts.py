"""MCTS / BatchMCTS with the reference's signatures, running on the GPU-resident tree.

Mirrors `src/mcts/mcts.py` (class MCTS, :18-362) and `BatchMCTS`
(`src/train/parallel_self_play.py:32-214`): same constructor arguments and public
attributes, `search(board, num_simulations, temperature, add_dirichlet_noise) ->
(np.float32[65], float)`, `get_action_probs`, `get_best_action`, `get_action_evaluations`,
`search_batch`, `batch_predict`.

Select / expand / backup run in csrc/search.cu (a group of 8 lanes per game, four games per warp).  The leaf evaluator is
  * the native tcgen05 network when `model` is an OthelloResNet-shaped torch module
    (weights are re-read whenever the module's tensors change), or
  * any other callable `model(x[B,3,8,8]) -> (log_probs, value)` through the
    external-evaluator entry points (tree on the GPU, network wherever the model lives), or
  * the built-in integer test evaluator (`evaluator="hashnet"`), used by the parity tests.

Default flags reproduce the reference exactly (root visit count stays 0 so the root is
an arg-max over Q and Dirichlet noise is inert; child Q is not negated) -- see SURVEY.md 0.3.
`root_n_sum=True` / `q_canonical=True` switch to the textbook AlphaZero behaviour.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import bitboard as bb
from ._lib import MEM_HOST, Context, check, ptr


class TreeSearch:
    """Thin owner of an oth_search handle (batched, host-array API)."""

    def __init__(self, ctx: Context, capacity: int, max_sims: int):
        self.ctx, self.capacity, self.max_sims = ctx, int(capacity), int(max_sims)
        h = C.c_void_p()
        check(ctx.lib.oth_search_create(ctx.handle, self.capacity, self.max_sims, C.byref(h)))
        self.handle = h
        self.n = 0

    def configure(self, c_puct, alpha, eps, flags):
        check(self.ctx.lib.oth_search_configure(self.handle, float(c_puct), float(alpha), float(eps), int(flags)))

    def begin(self, self_b: np.ndarray, opp_b: np.ndarray):
        self.n = int(self_b.size)
        check(self.ctx.lib.oth_search_begin(self.handle, ptr(self_b), ptr(opp_b), self.n, MEM_HOST))

    def collect(self):
        ls = np.empty(self.n, np.uint64); lo = np.empty(self.n, np.uint64); need = np.empty(self.n, np.uint8)
        check(self.ctx.lib.oth_search_collect(self.handle, ptr(ls), ptr(lo), ptr(need), MEM_HOST))
        return ls, lo, need

    def apply(self, probs: np.ndarray, value: np.ndarray):
        probs = np.ascontiguousarray(probs, np.float32); value = np.ascontiguousarray(value, np.float32)
        assert probs.shape == (self.n, 65) and value.shape == (self.n,)
        check(self.ctx.lib.oth_search_apply(self.handle, ptr(probs), ptr(value), MEM_HOST))

    def run(self, net_handle, sims: int, add_noise: bool, seed: int, inflight: int = 1):
        if inflight > 1:
            check(self.ctx.lib.oth_search_run_waves(self.handle, net_handle, int(sims), int(inflight), int(bool(add_noise)),
                                                    int(seed) & (2**64 - 1)))
        else:
            check(self.ctx.lib.oth_search_run(self.handle, net_handle, int(sims), int(bool(add_noise)), int(seed) & (2**64 - 1)))

    def results(self):
        v = np.empty((self.n, 65), np.int32); q = np.empty((self.n, 65), np.float64); e = np.empty(self.n, np.int32)
        check(self.ctx.lib.oth_search_results(self.handle, ptr(v), ptr(q), ptr(e), MEM_HOST))
        return v, q, e

    def stats(self) -> dict:
        v = (C.c_uint64 * 4)()
        check(self.ctx.lib.oth_search_stats(self.handle, v))
        return {"nn_positions": int(v[0]), "cache_hits": int(v[1]), "same_step_duplicates": int(v[2]), "hash_collisions": int(v[3])}

    def policy(self, temperature: float) -> np.ndarray:
        p = np.empty((self.n, 65), np.float32)
        check(self.ctx.lib.oth_search_policy(self.handle, float(temperature), ptr(p), MEM_HOST))
        return p

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.oth_search_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _boards_to_arrays(boards):
    s = np.fromiter((int(b.self_board) for b in boards), np.uint64, len(boards))
    o = np.fromiter((int(b.opp_board) for b in boards), np.uint64, len(boards))
    return s, o


class MCTS:
    """Drop-in for `src.mcts.mcts.MCTS` (mcts.py:18-47)."""

    def __init__(self, model, device, c_puct: float = 1.0, dirichlet_alpha: float = 0.3,
                 dirichlet_epsilon: float = 0.25, *, evaluator: str = "auto", root_n_sum: bool = False,
                 q_canonical: bool = False, engine: str | None = None, eval_cache: bool = False,
                 inflight: int = 1, ctx: Context | None = None):
        self.model = model
        self.device = device
        self.c_puct = c_puct
        self.dirichlet_alpha = dirichlet_alpha
        self.dirichlet_epsilon = dirichlet_epsilon
        self.root_n_sum, self.q_canonical = root_n_sum, q_canonical
        self.eval_cache = eval_cache          # position-keyed evaluation cache (native evaluator only; same results)
        self.inflight = int(inflight)         # > 1: K simulations per game per launch with virtual loss (deviates from the reference)
        self._ctx = ctx
        self._engine = engine
        self._tree: TreeSearch | None = None
        self._net = None
        if evaluator == "auto":
            sd = getattr(model, "state_dict", None)
            from .net import is_resnet_state_dict
            if model is None:
                evaluator = "hashnet"
            elif callable(sd) and is_resnet_state_dict(model.state_dict()):
                evaluator = "native"
            else:
                evaluator = "external"
        if evaluator not in ("native", "external", "hashnet"):
            raise ValueError(f"unknown evaluator {evaluator!r}")
        self.evaluator = evaluator
        self.last_n_evals = None

    # -- plumbing --------------------------------------------------------------------------------
    def _context(self) -> Context:
        if self._ctx is None:
            self._ctx = Context.default(_lib.device_index(self.device))
        return self._ctx

    def _flags(self) -> int:
        f = 0
        if self.root_n_sum:
            f |= _lib.FLAG_ROOT_N_SUM
        if self.q_canonical:
            f |= _lib.FLAG_Q_CANONICAL
        if self.evaluator == "hashnet":
            f |= _lib.FLAG_EVAL_HASHNET
        if self.eval_cache and self.evaluator == "native":
            f |= _lib.FLAG_EVAL_CACHE
        return f

    def _tree_for(self, n: int, sims: int) -> TreeSearch:
        t = self._tree
        if t is None or t.capacity < n or t.max_sims < sims:
            if t is not None:
                t.close()
            cap = max(n, 1)
            t = self._tree = TreeSearch(self._context(), cap, max(sims, 1))
        t.configure(self.c_puct, self.dirichlet_alpha, self.dirichlet_epsilon, self._flags())
        return t

    def _native_net(self):
        from .net import InferenceNet
        if self._net is None:
            self._net = InferenceNet.from_module(self.model, self._context(), self._engine)
        elif self._net.sync_from(self.model) and self._tree is not None:
            check(self._context().lib.oth_search_invalidate_cache(self._tree.handle))   # new weights: cached outputs are stale
        return self._net

    def _call_model(self, leaf_self, leaf_opp):
        """mcts.py:174-208 for a batch: planes -> model -> (exp(log_probs), value)."""
        import torch
        planes = bb.tensor_input(leaf_self, leaf_opp, self._context())
        x = torch.from_numpy(planes).float().to(self.device)
        self.model.eval()
        with torch.no_grad():
            logp, v = self.model(x)
            probs = torch.exp(logp).cpu().numpy()
            val = v.reshape(-1).cpu().numpy()
        return probs.astype(np.float32, copy=False), val.astype(np.float32, copy=False)

    def _run(self, self_b, opp_b, num_simulations: int, add_noise: bool) -> TreeSearch:
        n = int(self_b.size)
        t = self._tree_for(n, num_simulations)
        t.begin(self_b, opp_b)
        seed = int(np.random.randint(0, 2**31 - 1)) if add_noise else 0    # the reference draws from np.random here
        if self.evaluator == "external":
            for _ in range(num_simulations + 1):                            # root evaluation + simulations
                ls, lo, need = t.collect()
                if need.any():
                    probs, val = self._call_model(ls, lo)
                else:
                    probs = np.zeros((n, 65), np.float32); val = np.zeros(n, np.float32)
                t.apply(probs, val)
        else:
            net = self._native_net().handle if self.evaluator == "native" else None
            t.run(net, num_simulations, add_noise, seed, self.inflight)
        return t

    # -- the reference's public API ----------------------------------------------------------------
    def search(self, board, num_simulations: int, temperature: float = 1.0, add_dirichlet_noise: bool = False) -> tuple:
        """mcts.py:49-98 -> (policy float32[65], root value).  The root value is the root's
        Q, which the reference never updates: 0.0."""
        res = self.search_batch([board], num_simulations, temperature, add_dirichlet_noise)
        return res[0]

    def search_batch(self, boards, num_simulations: int, temperature: float = 1.0,
                     add_dirichlet_noise: bool = False) -> list:
        """parallel_self_play.py:80-170: one independent search per board, lock-step on the GPU."""
        if len(boards) == 0:
            return []
        s, o = _boards_to_arrays(boards)
        t = self._run(s, o, num_simulations, add_dirichlet_noise)
        pol = t.policy(temperature)
        _, _, nev = t.results()
        self.last_n_evals = nev
        return [(pol[i].copy(), 0.0) for i in range(len(boards))]

    def search_arrays(self, self_b, opp_b, num_simulations: int, add_dirichlet_noise: bool = False):
        """Batched search on raw board words -> (visits int32[n,65], q float64[n,65], n_evals int32[n])."""
        s = np.ascontiguousarray(self_b, np.uint64); o = np.ascontiguousarray(opp_b, np.uint64)
        t = self._run(s, o, num_simulations, add_dirichlet_noise)
        return t.results()

    def get_action_probs(self, board, num_simulations: int, temperature: float = 1.0,
                         add_dirichlet_noise: bool = False) -> np.ndarray:
        """mcts.py:230-255"""
        return self.search(board, num_simulations, temperature, add_dirichlet_noise)[0]

    def get_best_action(self, board, num_simulations: int) -> int:
        """mcts.py:257-296"""
        legal = board.get_legal_moves()
        if len(legal) == 0:
            return 64
        if num_simulations < 1:
            return legal[0]
        policy, _ = self.search(board, num_simulations=num_simulations, temperature=0.0, add_dirichlet_noise=False)
        best, best_p = legal[0], policy[legal[0]]
        for a in legal:
            if policy[a] > best_p:
                best, best_p = a, policy[a]
        return best

    def get_action_evaluations(self, board, num_simulations: int) -> np.ndarray:
        """mcts.py:298-362: int32[65], 0..100 from the root children's Q."""
        out = np.zeros(65, dtype=np.int32)
        legal = board.get_legal_moves()
        if len(legal) == 0 or num_simulations < 1:
            return out
        s, o = _boards_to_arrays([board])
        t = self._run(s, o, num_simulations, False)
        _, q, _ = t.results()
        for a in legal:
            out[a] = max(0, min(100, int((q[0, a] + 1.0) * 50.0)))
        return out


class BatchMCTS(MCTS):
    """Drop-in for `BatchMCTS` (parallel_self_play.py:32-214)."""

    def batch_predict(self, boards):
        """parallel_self_play.py:53-78 -> (policies float32[n,65] probabilities, values float32[n,1])."""
        if len(boards) == 0:
            return np.array([]), np.array([])
        s, o = _boards_to_arrays(boards)
        if self.evaluator == "native":
            pol, val = self._native_net().forward(s, o, out="probs")
            return pol, val.reshape(-1, 1)
        if self.evaluator == "external":
            pol, val = self._call_model(s, o)
            return pol, val.reshape(-1, 1)
        raise RuntimeError("batch_predict needs a model")
