"""ctypes binding of libothello_b200.so (the C ABI declared in include/othello_b200.h).

There is no CPU fallback: if the shared object cannot be loaded (and cannot be
built because nvcc is absent) importing this module raises; if no sm_100 device
is present, creating a context raises with the library's own message.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OTH_LIB_PATH") or os.path.join(PKG_DIR, "libothello_b200.so")   # override: experiment builds

OTH_OK = 0
MEM_DEVICE, MEM_HOST = 0, 1
ENGINE_TCGEN05, ENGINE_SIMT, ENGINE_TCGEN05_PAIR = 0, 1, 2
OUT_LOGPROBS, OUT_PROBS, OUT_PRIORS = 0, 1, 2
FLAG_ROOT_N_SUM, FLAG_Q_CANONICAL, FLAG_WINNER_BLACK, FLAG_EVAL_HASHNET, FLAG_EVAL_CACHE = 1, 2, 4, 8, 16
FLAG_NO_SEARCH_SHARING = 32
SCHEDULE_AUTO, SCHEDULE_LOCKSTEP, SCHEDULE_ASYNC = 0, 1, 2
ACTIONS = 65


class OthelloB200Error(RuntimeError):
    pass


class SelfPlayConfig(C.Structure):
    _fields_ = [
        ("num_simulations", C.c_int32), ("temperature_threshold", C.c_int32),
        ("add_dirichlet_noise", C.c_int32), ("concurrent_games", C.c_int32),
        ("c_puct", C.c_double), ("dirichlet_alpha", C.c_double), ("dirichlet_epsilon", C.c_double),
        ("flags", C.c_uint32), ("schedule", C.c_uint32), ("seed", C.c_uint64),
    ]


class BoardState(C.Structure):
    """oth_board_state (include/othello_b200.h)"""
    _fields_ = [("self_b", C.c_uint64), ("opp_b", C.c_uint64), ("legal", C.c_uint64), ("move_count", C.c_int32),
                ("ok", C.c_int32), ("terminal", C.c_int32), ("winner", C.c_int32), ("self_count", C.c_int32),
                ("opp_count", C.c_int32), ("seq", C.c_uint64)]


ACTION_NONE = -1000

SAMPLE_DTYPE = np.dtype([
    ("self_b", np.uint64), ("opp_b", np.uint64), ("legal", np.uint64),
    ("game", np.int32), ("ply", np.int16), ("value", np.int8), ("n_children", np.uint8),
    ("visits", np.uint16, (ACTIONS,)), ("pad", np.uint16, (3,)),
])
assert SAMPLE_DTYPE.itemsize == 168

_p = C.c_void_p
_i64 = C.c_int64
_PROTOS = {
    # name: (restype, argtypes)
    "oth_last_error": (C.c_char_p, []),
    "oth_version": (C.c_char_p, []),
    "oth_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "oth_ctx_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "oth_ctx_destroy": (C.c_int, [_p]),
    "oth_ctx_sync": (C.c_int, [_p]),
    "oth_ctx_stream": (C.c_uint64, [_p]),
    "oth_ctx_launch_count": (C.c_uint64, [_p]),
    "oth_ctx_timing_enable": (C.c_int, [_p, C.c_int]),
    "oth_ctx_timing_read": (C.c_int, [_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "oth_legal_moves": (C.c_int, [_p, _p, _p, _p, _i64, C.c_int]),
    "oth_flips": (C.c_int, [_p, _p, _p, _p, _p, _i64, C.c_int]),
    "oth_make_move": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, C.c_int]),
    "oth_terminal_winner": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, C.c_int]),
    "oth_tensor_input": (C.c_int, [_p, _p, _p, _p, _i64, C.c_int]),
    "oth_board_step": (C.c_int, [_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(BoardState)]),
    "oth_perft": (C.c_int, [_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]),
    "oth_random_playouts": (C.c_int, [_p, _i64, C.c_uint64, C.POINTER(_i64), C.POINTER(_i64), _p, _p, _p, C.c_int]),
    "oth_choose_random": (C.c_int, [_p, _p, _p, _p, C.c_uint64, _p, _i64, C.c_int]),
    "oth_choose_greedy": (C.c_int, [_p, _p, _p, _p, _p, _i64, C.c_int]),
    "oth_replay_create": (C.c_int, [_p, _i64, C.POINTER(_p)]),
    "oth_replay_destroy": (C.c_int, [_p]),
    "oth_replay_size": (_i64, [_p]),
    "oth_replay_clear": (C.c_int, [_p]),
    "oth_replay_add": (C.c_int, [_p, _p, _i64, C.c_int]),
    "oth_replay_gather": (C.c_int, [_p, _p, _i64, _p, _p, _p, C.c_int]),
    "oth_replay_gather_sym": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, C.c_int]),
    "oth_replay_check": (C.c_int, [_p]),
    "oth_replay_value_stats": (C.c_int, [_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "oth_net_create": (C.c_int, [_p, C.c_int, C.c_int, C.POINTER(_p)]),
    "oth_net_destroy": (C.c_int, [_p]),
    "oth_net_param_count": (_i64, [_p]),
    "oth_net_load_weights": (C.c_int, [_p, _p, _i64]),
    "oth_net_set_engine": (C.c_int, [_p, C.c_int]),
    "oth_net_engine": (C.c_int, [_p]),
    "oth_net_forward": (C.c_int, [_p, _p, _p, _i64, _p, _p, C.c_int, C.c_int]),
    "oth_search_create": (C.c_int, [_p, _i64, C.c_int, C.POINTER(_p)]),
    "oth_search_destroy": (C.c_int, [_p]),
    "oth_search_configure": (C.c_int, [_p, C.c_double, C.c_double, C.c_double, C.c_uint32]),
    "oth_search_begin": (C.c_int, [_p, _p, _p, _i64, C.c_int]),
    "oth_search_collect": (C.c_int, [_p, _p, _p, _p, C.c_int]),
    "oth_search_apply": (C.c_int, [_p, _p, _p, C.c_int]),
    "oth_search_run": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_uint64]),
    "oth_search_run_waves": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "oth_search_results": (C.c_int, [_p, _p, _p, _p, C.c_int]),
    "oth_search_policy": (C.c_int, [_p, C.c_double, _p, C.c_int]),
    "oth_search_stats": (C.c_int, [_p, C.POINTER(C.c_uint64)]),
    "oth_search_invalidate_cache": (C.c_int, [_p]),
    "oth_selfplay_stats": (C.c_int, [_p, C.POINTER(C.c_uint64)]),
    "oth_selfplay_timing": (C.c_int, [_p, C.POINTER(C.c_double)]),
    "oth_selfplay_set_seed": (C.c_int, [_p, C.c_uint64]),
    "oth_selfplay_create": (C.c_int, [_p, C.POINTER(SelfPlayConfig), C.POINTER(_p)]),
    "oth_selfplay_destroy": (C.c_int, [_p]),
    "oth_selfplay_run": (C.c_int, [_p, _p, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "oth_selfplay_fetch": (C.c_int, [_p, _p, _i64, C.c_int]),
    "oth_selfplay_samples_device": (C.c_int, [_p, C.POINTER(_p), C.POINTER(_i64)]),
    "oth_debug_net_trace": (C.c_int, [_p, _p, _p, _i64, _p, C.c_int]),
    "oth_debug_umma_probe": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int] + [C.c_uint32] * 8 + [_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None
_lock = threading.Lock()


def load():
    """dlopen the in-tree library (building it first if it is missing and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build as _build
            try:
                _build.build()
            except Exception as e:  # no nvcc, compile error ...
                raise ImportError(
                    f"{LIB_PATH} is missing and could not be built ({e}); "
                    "this package has no CPU fallback") from e
        lib = C.CDLL(LIB_PATH)
        missing = [n for n in _PROTOS if not hasattr(lib, n)]
        if missing:
            # stale binary (sources moved on): rebuild once if a compiler is around, else fail loudly
            from . import build as _build
            try:
                _build.build()
            except Exception as e:
                raise ImportError(f"{LIB_PATH} lacks {missing} and could not be rebuilt ({e})") from e
            lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)      # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().oth_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OTH_OK:
        raise OthelloB200Error(f"[{rc}] {last_error()}")


def ptr(a) -> int | None:
    """Address of a numpy array / torch tensor / raw int pointer (None stays None)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)!r}")


class Context:
    """One CUDA device + stream (oth_ctx).  One host thread at a time."""

    _default: dict[int, "Context"] = {}

    def __init__(self, device: int = 0):
        lib = load()
        h = _p()
        check(lib.oth_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)
        self.lib = lib

    @classmethod
    def default(cls, device: int = 0) -> "Context":
        ctx = cls._default.get(device)
        if ctx is None:
            ctx = cls._default[device] = Context(device)
        return ctx

    def sync(self) -> None:
        check(self.lib.oth_ctx_sync(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.oth_ctx_stream(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.oth_ctx_launch_count(self.handle))

    def timing_enable(self, on: bool = True) -> None:
        check(self.lib.oth_ctx_timing_enable(self.handle, int(on)))

    def timing_read(self) -> dict:
        """{'net': (ms, launches), 'tree': (...), 'move': (...)} since the last read."""
        ms = (C.c_double * 3)(); cnt = (C.c_uint64 * 3)()
        check(self.lib.oth_ctx_timing_read(self.handle, ms, cnt))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(("net", "tree", "move"))}

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.oth_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class torch_order:
    """Orders a library call that takes torch CUDA tensors with torch's own stream.

    The library works on its context stream (created non-blocking), torch on its current stream: without this the
    kernel could read inputs torch has not produced yet, and torch could consume -- or its caching allocator could
    recycle -- outputs the kernel has not written yet.  On entry the context stream waits for an event recorded on
    torch's current stream; on exit the current stream waits for the context stream and every tensor is marked as
    used by the context stream (`record_stream`).  No host synchronisation.  A no-op for host arrays, and when
    torch's current stream IS the context stream (bench.py runs that way)."""

    def __init__(self, ctx: "Context", *tensors):
        self.ctx = ctx
        self.tensors = [t for t in tensors if hasattr(t, "is_cuda") and t.is_cuda]

    def __enter__(self):
        self.ext = None
        if not self.tensors:
            return self
        import torch
        dev = self.tensors[0].device
        cur = torch.cuda.current_stream(dev)
        if cur.cuda_stream == self.ctx.stream:
            return self
        self.cur, self.ext = cur, torch.cuda.ExternalStream(self.ctx.stream, device=dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.ext.wait_event(ev)
        return self

    def __exit__(self, *exc):
        if self.ext is None:
            return False
        import torch
        ev = torch.cuda.Event()
        ev.record(self.ext)
        self.cur.wait_event(ev)
        for t in self.tensors:
            t.record_stream(self.ext)
        return False


def device_index(device) -> int:
    """torch.device / str / int -> CUDA ordinal of the engine.

    `device` in the reference's signatures says where the torch module lives.  The
    engine itself only exists on CUDA: "cuda:N" selects GPU N, anything else (e.g. a
    model kept on the CPU) runs the engine on GPU 0.  Without a GPU, context creation
    fails with the library's "no CPU fallback" error.
    """
    if device is None:
        return 0
    if isinstance(device, (int, np.integer)):
        return int(device)
    s = str(device)
    if s.startswith("cuda") and ":" in s:
        return int(s.split(":")[1])
    return 0
