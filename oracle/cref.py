"""oracle/cref.py -- TEST INFRASTRUCTURE ONLY.

numpy/ctypes view of the C restatement in ``ref_rules.c``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")

EVAL_FN = C.CFUNCTYPE(None, C.c_uint64, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


class MctsCfg(C.Structure):
    _fields_ = [("c_puct", C.c_double), ("num_simulations", C.c_int),
                ("root_n_sum", C.c_int), ("q_canonical", C.c_int)]


class MctsResult(C.Structure):
    _fields_ = [("visits", C.c_int32 * 65), ("q", C.c_double * 65), ("n_children", C.c_int),
                ("n_evals", C.c_int), ("max_depth", C.c_int), ("is_child", C.c_int * 65)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build_oracle()
    L = C.CDLL(path)
    L.ref_flips.restype = C.c_uint64
    L.ref_flips.argtypes = [C.c_int, C.c_uint64, C.c_uint64]
    L.ref_legal.restype = C.c_uint64
    L.ref_legal.argtypes = [C.c_uint64, C.c_uint64]
    L.ref_is_terminal.argtypes = [C.c_uint64, C.c_uint64]
    L.ref_winner.argtypes = [C.c_uint64, C.c_uint64]
    L.ref_perft.restype = C.c_uint64
    L.ref_perft.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
    L.ref_make_move.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int),
                                C.POINTER(C.c_int), C.c_int]
    L.ref_legal_batch.argtypes = [_u64p, _u64p, _u64p, C.c_int64]
    L.ref_flips_batch.argtypes = [_u64p, _u64p, _i32p, _u64p, C.c_int64]
    L.ref_make_move_batch.argtypes = [_u64p, _u64p, _i32p, _i32p, _u8p, C.c_int64]
    L.ref_terminal_winner_batch.argtypes = [_u64p, _u64p, _u8p, _i8p, C.c_int64]
    L.ref_tensor_input_batch.argtypes = [_u64p, _u64p, _f32p, C.c_int64]
    L.ref_random_playouts.restype = C.c_int64
    L.ref_random_playouts.argtypes = [C.c_int64, C.c_uint64, C.c_int, _i64p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_hashnet.argtypes = [C.c_uint64, C.c_uint64, _f32p, _f32p]
    L.ref_mcts_search.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(MctsCfg), EVAL_FN, C.c_void_p,
                                  C.POINTER(MctsResult)]
    L.ref_expand_priors.argtypes = [_f32p, _i32p, C.c_int, _f32p]
    L.ref_policy_from_visits.argtypes = [_i32p, _i32p, C.c_double, _f32p]
    L.ref_mcts_search_hashnet_batch.argtypes = [_u64p, _u64p, C.c_int64, C.c_double, C.c_int, C.c_int,
                                                _i32p, C.c_void_p]
    L.ref_batch_create.restype = C.c_void_p
    L.ref_batch_create.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
    L.ref_batch_destroy.argtypes = [C.c_void_p]
    L.ref_batch_begin.argtypes = [C.c_void_p, _u64p, _u64p, C.c_int]
    L.ref_batch_collect.argtypes = [C.c_void_p, C.c_int, _u64p, _u64p, _u8p]
    L.ref_batch_apply.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
    L.ref_batch_visits.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_void_p]
    _lib = L
    return L


START_SELF = (1 << 28) | (1 << 35)
START_OPP = (1 << 27) | (1 << 36)


def legal(s: int, o: int) -> int:
    return int(lib().ref_legal(s, o))


def flips(pos: int, s: int, o: int) -> int:
    return int(lib().ref_flips(pos, s, o))


def is_terminal(s: int, o: int) -> bool:
    return bool(lib().ref_is_terminal(s, o))


def winner(s: int, o: int) -> int:
    return int(lib().ref_winner(s, o))


def perft(s: int, o: int, depth: int) -> int:
    return int(lib().ref_perft(s, o, depth))


def make_move(s: int, o: int, move_count: int, pos: int):
    """-> (ok, self, opp, move_count)"""
    a, b, m, p = C.c_uint64(s), C.c_uint64(o), C.c_int(move_count), C.c_int(0)
    ok = lib().ref_make_move(C.byref(a), C.byref(b), C.byref(m), C.byref(p), pos)
    return bool(ok), int(a.value), int(b.value), int(m.value)


def legal_list(s: int, o: int):
    m = legal(s, o)
    return [i for i in range(64) if (m >> i) & 1] or [64]


def legal_batch(s: np.ndarray, o: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(s, np.uint64); o = np.ascontiguousarray(o, np.uint64)
    out = np.empty_like(s)
    lib().ref_legal_batch(s, o, out, s.size)
    return out


def flips_batch(s, o, pos) -> np.ndarray:
    s = np.ascontiguousarray(s, np.uint64); o = np.ascontiguousarray(o, np.uint64)
    pos = np.ascontiguousarray(pos, np.int32)
    out = np.empty_like(s)
    lib().ref_flips_batch(s, o, pos, out, s.size)
    return out


def make_move_batch(s, o, mc, action):
    """-> (self, opp, move_count, ok) as fresh arrays"""
    s = np.array(s, np.uint64); o = np.array(o, np.uint64); mc = np.array(mc, np.int32)
    action = np.ascontiguousarray(action, np.int32)
    ok = np.empty(s.size, np.uint8)
    lib().ref_make_move_batch(s, o, mc, action, ok, s.size)
    return s, o, mc, ok


def terminal_winner_batch(s, o):
    s = np.ascontiguousarray(s, np.uint64); o = np.ascontiguousarray(o, np.uint64)
    t = np.empty(s.size, np.uint8); w = np.empty(s.size, np.int8)
    lib().ref_terminal_winner_batch(s, o, t, w, s.size)
    return t, w


def tensor_input_batch(s, o) -> np.ndarray:
    s = np.ascontiguousarray(s, np.uint64); o = np.ascontiguousarray(o, np.uint64)
    out = np.empty((s.size, 3, 8, 8), np.float32)
    lib().ref_tensor_input_batch(s, o, out.reshape(-1), s.size)
    return out


def random_playouts(n_games: int, seed: int, threads: int = 0, want_finals: bool = False):
    """-> dict(total_plies, winner_hist[-1,0,+1], final_self, final_opp, plies)"""
    hist = np.zeros(3, np.int64)
    fs = fo = pl = None
    a = b = c = None
    if want_finals:
        fs = np.empty(n_games, np.uint64); fo = np.empty(n_games, np.uint64); pl = np.empty(n_games, np.int32)
        a, b, c = fs.ctypes.data, fo.ctypes.data, pl.ctypes.data
    total = lib().ref_random_playouts(n_games, seed, threads, hist, a, b, c)
    return {"total_plies": int(total), "winner_hist": hist, "final_self": fs, "final_opp": fo, "plies": pl}


def hashnet(s: int, o: int):
    p = np.empty(65, np.float32); v = np.empty(1, np.float32)
    lib().ref_hashnet(s, o, p, v)
    return p, float(v[0])


def hashnet_batch(s, o):
    s = np.ascontiguousarray(s, np.uint64); o = np.ascontiguousarray(o, np.uint64)
    P = np.empty((s.size, 65), np.float32); V = np.empty(s.size, np.float32)
    v = np.empty(1, np.float32)
    L = lib()
    for i in range(s.size):
        L.ref_hashnet(int(s[i]), int(o[i]), P[i], v)
        V[i] = v[0]
    return P, V


def expand_priors(probs: np.ndarray, legal_actions) -> np.ndarray:
    probs = np.ascontiguousarray(probs, np.float32)
    la = np.ascontiguousarray(legal_actions, np.int32)
    out = np.empty(65, np.float32)
    lib().ref_expand_priors(probs, la, la.size, out)
    return out


def policy_from_visits(visits, is_child, temperature: float) -> np.ndarray:
    out = np.empty(65, np.float32)
    lib().ref_policy_from_visits(np.ascontiguousarray(visits, np.int32),
                                 np.ascontiguousarray(is_child, np.int32), float(temperature), out)
    return out


def mcts_search(s: int, o: int, num_simulations: int, c_puct: float = 1.0, evaluator=None,
                root_n_sum: bool = False, q_canonical: bool = False):
    """Run the C restatement of MCTS.search.

    ``evaluator(self, opp) -> (probs f32[65], value float)``; default = hash-net.
    Returns dict(visits i32[65], q f64[65], is_child, n_evals, max_depth).
    """
    L = lib()
    if evaluator is None:
        def evaluator(a, b):
            return hashnet(a, b)

    def _cb(a, b, p_ptr, v_ptr, _user):
        p, v = evaluator(int(a), int(b))
        p = np.asarray(p, np.float32)
        for i in range(65):
            p_ptr[i] = float(p[i])
        v_ptr[0] = float(np.float32(v))

    cb = EVAL_FN(_cb)
    cfg = MctsCfg(float(c_puct), int(num_simulations), int(root_n_sum), int(q_canonical))
    res = MctsResult()
    rc = L.ref_mcts_search(s, o, C.byref(cfg), cb, None, C.byref(res))
    if rc != 0:
        raise RuntimeError("ref_mcts_search failed")
    return {
        "visits": np.array(res.visits[:], np.int32),
        "q": np.array(res.q[:], np.float64),
        "is_child": np.array(res.is_child[:], np.int32),
        "n_evals": int(res.n_evals),
        "max_depth": int(res.max_depth),
    }


def mcts_search_hashnet_batch(s, o, num_simulations: int, c_puct: float = 1.0, threads: int = 0):
    s = np.ascontiguousarray(s, np.uint64); o = np.ascontiguousarray(o, np.uint64)
    visits = np.zeros((s.size, 65), np.int32)
    nev = np.zeros(s.size, np.int32)
    rc = lib().ref_mcts_search_hashnet_batch(s, o, s.size, float(c_puct), int(num_simulations), threads,
                                             visits.reshape(-1), nev.ctypes.data)
    if rc != 0:
        raise RuntimeError("ref_mcts_search_hashnet_batch failed")
    return visits, nev


class BatchSearch:
    """Lock-step batch of oracle searches with a batched external evaluator
    (restates BatchMCTS.search_batch, src/train/parallel_self_play.py:80-197)."""

    def __init__(self, n_games: int, max_simulations: int, c_puct: float = 1.0,
                 root_n_sum: bool = False, q_canonical: bool = False):
        self.n_cap = int(n_games)
        self.h = lib().ref_batch_create(self.n_cap, float(c_puct), int(max_simulations), int(root_n_sum), int(q_canonical))
        if not self.h:
            raise MemoryError("ref_batch_create")
        self.n = 0

    def search(self, s, o, num_simulations: int, evaluate):
        """evaluate(leaf_self, leaf_opp) -> (probs [k,65] float32, value [k] float32) for the k needed leaves.
        Returns (visits int32 [n,65], n_evals int32 [n])."""
        L = lib()
        s = np.ascontiguousarray(s, np.uint64); o = np.ascontiguousarray(o, np.uint64)
        n = self.n = int(s.size)
        if L.ref_batch_begin(self.h, s, o, n) != 0:
            raise RuntimeError("ref_batch_begin failed")
        ls = np.empty(n, np.uint64); lo = np.empty(n, np.uint64); need = np.empty(n, np.uint8)
        probs = np.zeros((n, 65), np.float32); val = np.zeros(n, np.float32)
        for _ in range(num_simulations + 1):
            k = L.ref_batch_collect(self.h, n, ls, lo, need)
            if k:
                idx = np.flatnonzero(need)
                p, v = evaluate(ls[idx], lo[idx])
                probs[idx] = p; val[idx] = np.asarray(v, np.float32).reshape(-1)
            if L.ref_batch_apply(self.h, n, probs.reshape(-1), val) != 0:
                raise RuntimeError("ref_batch_apply failed")
        vis = np.zeros((n, 65), np.int32); nev = np.zeros(n, np.int32)
        L.ref_batch_visits(self.h, n, vis.reshape(-1), nev.ctypes.data)
        return vis, nev

    def close(self):
        if self.h:
            lib().ref_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
