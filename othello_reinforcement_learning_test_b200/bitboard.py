"""OthelloBitboard / BoardBatch -- the reference's board API on top of the CUDA kernels.

`OthelloBitboard` mirrors the Cython class of the reference (src/cython/bitboard.pyx,
surface in src/cython/bitboard.pxd:25-48): same attributes (`self_board`, `opp_board`,
`move_count`, `passed`), same methods, same return conventions (make_move returns a
bool and never raises; get_legal_moves never returns an empty list, `[64]` means pass).
Every rule evaluation is done by libothello_b200.so on the GPU (REF rules, bit-exact with
the reference, including its behaviour on the A/H files -- see csrc/bitboard.cuh); there
is no CPU rule engine in this package.

`BoardBatch` is the batched SoA form the hot path uses: thousands of games, one kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, Context, check, ptr

START_SELF = (1 << 28) | (1 << 35)   # bitboard.pyx:60-66
START_OPP = (1 << 27) | (1 << 36)
PASS = 64


# ---------------------------------------------------------------------------------------------
# batched operations (numpy host arrays or torch CUDA tensors)
# ---------------------------------------------------------------------------------------------

def _mem_of(*arrays) -> int:
    kinds = set()
    for a in arrays:
        if a is None:
            continue
        kinds.add(MEM_DEVICE if (hasattr(a, "is_cuda") and a.is_cuda) else MEM_HOST)
    if len(kinds) > 1:
        raise ValueError("mix of host and device arrays")
    return kinds.pop() if kinds else MEM_HOST


def _like(a, dtype_np, shape, torch_dtype=None):
    if hasattr(a, "is_cuda") and a.is_cuda:
        import torch
        return torch.empty(shape, dtype=torch_dtype, device=a.device)
    return np.empty(shape, dtype_np)


def _u64(a):
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a, np.uint64)
    if hasattr(a, "is_cuda"):
        import torch
        # torch stores the words as int64 (same bits)
        assert a.dtype in (torch.int64, torch.uint64), "board tensors must be 64-bit integers"
        return a.contiguous()
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def legal_moves(self_b, opp_b, ctx: Context | None = None):
    """get_legal_moves_bits for a batch (bitboard.pyx:187-193)."""
    ctx = ctx or Context.default()
    s, o = _u64(self_b), _u64(opp_b)
    import torch  # noqa: F401  (only needed for device tensors)
    out = _like(s, np.uint64, s.shape, getattr(s, "dtype", None))
    n = int(np.prod(s.shape))
    with _lib.torch_order(ctx, s, o, out):
        check(ctx.lib.oth_legal_moves(ctx.handle, ptr(s), ptr(o), ptr(out), n, _mem_of(s, o)))
    return out


def flips(self_b, opp_b, pos, ctx: Context | None = None):
    """_get_flip_bits for a batch (bitboard.pyx:116-133)."""
    ctx = ctx or Context.default()
    s, o = _u64(self_b), _u64(opp_b)
    p = np.ascontiguousarray(pos, np.int32) if not hasattr(pos, "is_cuda") else pos.contiguous()
    out = _like(s, np.uint64, s.shape, getattr(s, "dtype", None))
    with _lib.torch_order(ctx, s, o, p, out):
        check(ctx.lib.oth_flips(ctx.handle, ptr(s), ptr(o), ptr(p), ptr(out), int(np.prod(s.shape)), _mem_of(s, o, p)))
    return out


def make_move(self_b, opp_b, move_count, action, ctx: Context | None = None):
    """make_move for a batch, IN PLACE on the given arrays; returns ok (uint8)."""
    ctx = ctx or Context.default()
    mem = _mem_of(self_b, opp_b, move_count, action)
    n = int(np.prod(self_b.shape))
    if mem == MEM_HOST:
        assert self_b.dtype == np.uint64 and opp_b.dtype == np.uint64
        action = np.ascontiguousarray(action, np.int32)
        ok = np.empty(n, np.uint8)
    else:
        import torch
        ok = torch.empty(n, dtype=torch.uint8, device=self_b.device)
    with _lib.torch_order(ctx, self_b, opp_b, move_count, action, ok):
        check(ctx.lib.oth_make_move(ctx.handle, ptr(self_b), ptr(opp_b), ptr(move_count), ptr(action), ptr(ok), n, mem))
    return ok


def terminal_winner(self_b, opp_b, ctx: Context | None = None):
    """(is_terminal uint8, get_winner int8, stone counts int32 [n,2]) for a batch."""
    ctx = ctx or Context.default()
    s, o = _u64(self_b), _u64(opp_b)
    n = int(np.prod(s.shape))
    mem = _mem_of(s, o)
    if mem == MEM_HOST:
        t = np.empty(n, np.uint8); w = np.empty(n, np.int8); c = np.empty((n, 2), np.int32)
    else:
        import torch
        t = torch.empty(n, dtype=torch.uint8, device=s.device); w = torch.empty(n, dtype=torch.int8, device=s.device)
        c = torch.empty((n, 2), dtype=torch.int32, device=s.device)
    with _lib.torch_order(ctx, s, o, t, w, c):
        check(ctx.lib.oth_terminal_winner(ctx.handle, ptr(s), ptr(o), ptr(t), ptr(w), ptr(c), n, mem))
    return t, w, c


def tensor_input(self_b, opp_b, ctx: Context | None = None):
    """get_tensor_input for a batch: float32 [n,3,8,8] (bitboard.pyx:300-323)."""
    ctx = ctx or Context.default()
    s, o = _u64(self_b), _u64(opp_b)
    n = int(np.prod(s.shape))
    mem = _mem_of(s, o)
    if mem == MEM_HOST:
        out = np.empty((n, 3, 8, 8), np.float32)
    else:
        import torch
        out = torch.empty((n, 3, 8, 8), dtype=torch.float32, device=s.device)
    with _lib.torch_order(ctx, s, o, out):
        check(ctx.lib.oth_tensor_input(ctx.handle, ptr(s), ptr(o), ptr(out), n, mem))
    return out


def perft(depth: int, self_b: int = START_SELF, opp_b: int = START_OPP, ctx: Context | None = None) -> int:
    ctx = ctx or Context.default()
    out = C.c_uint64(0)
    check(ctx.lib.oth_perft(ctx.handle, int(self_b), int(opp_b), int(depth), C.byref(out)))
    return int(out.value)


def random_playouts(n_games: int, seed: int = 0, want_finals: bool = False, ctx: Context | None = None):
    """benchmark.py's random games, all on the GPU.  Returns a dict like the oracle's."""
    ctx = ctx or Context.default()
    total = C.c_int64(0)
    hist = (C.c_int64 * 3)()
    fs = fo = pl = None
    if want_finals:
        fs = np.empty(n_games, np.uint64); fo = np.empty(n_games, np.uint64); pl = np.empty(n_games, np.int32)
    check(ctx.lib.oth_random_playouts(ctx.handle, int(n_games), int(seed) & (2**64 - 1), C.byref(total), hist,
                                      ptr(fs), ptr(fo), ptr(pl), MEM_HOST))
    return {"total_plies": int(total.value), "winner_hist": np.array(list(hist), np.int64),
            "final_self": fs, "final_opp": fo, "plies": pl}


class BoardBatch:
    """n games as SoA arrays (host numpy).  Thin convenience over the batched calls."""

    def __init__(self, n: int, ctx: Context | None = None):
        self.ctx = ctx or Context.default()
        self.self_board = np.full(n, START_SELF, np.uint64)
        self.opp_board = np.full(n, START_OPP, np.uint64)
        self.move_count = np.zeros(n, np.int32)

    def __len__(self):
        return self.self_board.size

    def reset(self):
        self.self_board[:] = START_SELF; self.opp_board[:] = START_OPP; self.move_count[:] = 0

    def legal_moves(self):
        return legal_moves(self.self_board, self.opp_board, self.ctx)

    def make_move(self, action):
        return make_move(self.self_board, self.opp_board, self.move_count, action, self.ctx)

    def terminal_winner(self):
        return terminal_winner(self.self_board, self.opp_board, self.ctx)

    def tensor_input(self):
        return tensor_input(self.self_board, self.opp_board, self.ctx)


# ---------------------------------------------------------------------------------------------
# the reference's single-board class
# ---------------------------------------------------------------------------------------------

class OthelloBitboard:
    """Drop-in for `src.cython.bitboard.OthelloBitboard` (bitboard.pxd:11-48).

    Every rule evaluation is one call of `oth_board_step` (one launch, result in a mapped page-locked mailbox: no
    allocation, no copies): `make_move` applies the move AND returns the legal mask, terminal flag, winner and disc
    counts of the position reached, so the calls that follow a move in the reference's loops (arena.py:106-119,
    self_play.py:71-120: is_terminal, get_legal_moves, get_winner, get_stone_counts, get_tensor_input) are answered from
    that result without touching the GPU again.  Writing `self_board` / `opp_board` drops the cached answers."""

    __slots__ = ("_s", "_o", "move_count", "passed", "_st", "_ctx")

    def __init__(self):
        self._ctx = None
        self.reset()

    # -- state (public, read/write like the cdef public fields) --------------------------------
    @property
    def self_board(self) -> int:
        return self._s

    @self_board.setter
    def self_board(self, v: int) -> None:
        self._s = int(v) & 0xFFFFFFFFFFFFFFFF
        self._st = None

    @property
    def opp_board(self) -> int:
        return self._o

    @opp_board.setter
    def opp_board(self, v: int) -> None:
        self._o = int(v) & 0xFFFFFFFFFFFFFFFF
        self._st = None

    def _context(self) -> Context:
        if self._ctx is None:
            self._ctx = Context.default()
        return self._ctx

    def reset(self) -> None:
        """bitboard.pyx:52-69"""
        self._s = START_SELF
        self._o = START_OPP
        self.move_count = 0
        self.passed = False
        self._st = None

    # -- rules: all evaluated by the CUDA library ------------------------------------------------
    def _step(self, action: int) -> "_lib.BoardState":
        ctx = self._context()
        st = _lib.BoardState()
        check(ctx.lib.oth_board_step(ctx.handle, self._s, self._o, int(self.move_count), action, C.byref(st)))
        return st

    def _state(self) -> "_lib.BoardState":
        """(legal, terminal, winner, counts) of the current position: cached from the last make_move, else one launch."""
        st = self._st
        if st is None:
            st = self._st = self._step(_lib.ACTION_NONE)
        return st

    def get_legal_moves_bits(self) -> int:
        """bitboard.pyx:187-193"""
        return int(self._state().legal)

    def get_legal_moves(self) -> list:
        """bitboard.pyx:166-185: ascending squares, or [64] when the side to move must pass."""
        m = int(self._state().legal)
        if m == 0:
            return [PASS]
        out = []
        while m:
            low = m & -m
            out.append(low.bit_length() - 1)
            m ^= low
        return out

    def make_move(self, pos: int) -> bool:
        """bitboard.pyx:195-247: True if applied; False (state untouched) otherwise."""
        try:
            pos = int(pos)
        except (TypeError, ValueError):
            return False
        if pos < -2**31 or pos >= 2**31 or pos == _lib.ACTION_NONE:
            return False
        st = self._step(pos)
        if not st.ok:
            return False
        self._s, self._o, self.move_count = int(st.self_b), int(st.opp_b), int(st.move_count)
        self.passed = pos == PASS
        self._st = st                    # the position reached comes with its legal mask / terminal / winner / counts
        return True

    def is_terminal(self) -> bool:
        """bitboard.pyx:249-264"""
        return bool(self._state().terminal)

    def get_winner(self) -> int:
        """bitboard.pyx:266-282: +1 side to move has more discs, -1 fewer, 0 equal."""
        return int(self._state().winner)

    def get_stone_counts(self) -> tuple:
        """bitboard.pyx:292-298"""
        st = self._state()
        return (int(st.self_count), int(st.opp_count))

    def get_tensor_input(self) -> np.ndarray:
        """bitboard.pyx:300-323: fresh float32 (3,8,8): self, opp, legal.  Pure format expansion of the three words
        (the legal mask is the kernel's)."""
        words = np.array([self._s, self._o, int(self._state().legal)], dtype="<u8")
        return np.unpackbits(words.view(np.uint8).reshape(3, 8), axis=1, bitorder="little").reshape(3, 8, 8).astype(np.float32)

    def copy(self) -> "OthelloBitboard":
        """bitboard.pyx:325-336"""
        b = OthelloBitboard.__new__(OthelloBitboard)
        b._ctx = self._ctx
        b._s, b._o, b.move_count, b.passed, b._st = self._s, self._o, self.move_count, self.passed, self._st
        return b

    def get_symmetries(self, pi) -> list:
        """bitboard.pyx:338-370: the 8 dihedral images of (planes, policy); pass prob carried."""
        planes = self.get_tensor_input()
        pi = np.asarray(pi)
        grid = pi[:64].reshape(8, 8)
        out = []
        for k in range(4):
            rp = np.rot90(planes, k, axes=(1, 2)); rg = np.rot90(grid, k)
            for flip in (False, True):
                p2 = np.flip(rp, axis=2) if flip else rp
                g2 = np.flip(rg, axis=1) if flip else rg
                v = np.zeros(65, np.float32)
                v[:64] = g2.reshape(-1)
                v[64] = pi[64]
                out.append((np.ascontiguousarray(p2), v))
        return out

    def __repr__(self) -> str:
        rows = ["  A B C D E F G H"]
        for r in range(8):
            cells = []
            for c in range(8):
                i = r * 8 + c
                cells.append("● " if (self._s >> i) & 1 else ("○ " if (self._o >> i) & 1 else ". "))
            rows.append(f"{r + 1} " + "".join(cells))
        return "\n".join(rows)

    def to_string(self) -> str:
        return self.__repr__()
