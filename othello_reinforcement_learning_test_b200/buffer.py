"""ReplayBuffer with the reference's interface, resident in HBM.

Mirrors `src/train/buffer.py:15-123`: `ReplayBuffer(max_size)`, `add(data)`, `add_single`,
`sample(batch_size) -> (states f32[B,3,8,8], policies f32[B,65], values f32[B,1])`, `__len__`, `clear`,
`is_ready`, `get_statistics`.  Samples are kept packed (168 B: three bit-planes, 65 visit counts, label) in
a device ring; `sample` draws indices with `random.sample` -- exactly what the reference does, so a seeded
run returns the same minibatch -- and one kernel gathers + expands them.  `sample_torch` hands the trainer
CUDA tensors without a host hop; `add_packed` / `add_from_worker` take the self-play engine's records as
they are (host array or straight from device memory).
"""
from __future__ import annotations

import ctypes as C
import random
from typing import List, Tuple

import numpy as np

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, Context, check, ptr


def pack_training_data(data) -> np.ndarray:
    """[(state f32[3,8,8], policy f32[65], value)] -> packed oth_sample records.

    The planes are 0/1 by construction (get_tensor_input).  Policies are stored as 16-bit counts: a visit
    distribution n_i / N is recovered exactly when N <= 65535 (counts = round(p * N) with N inferred as the
    smallest total that makes every entry integral); arbitrary float policies are quantised to 1/65535."""
    n = len(data)
    out = np.zeros(n, _lib.SAMPLE_DTYPE)
    if n == 0:
        return out
    w = (1 << np.arange(64, dtype=np.uint64))
    states = np.asarray([d[0] for d in data], dtype=np.float32).reshape(n, 3, 64)
    for k, name in enumerate(("self_b", "opp_b", "legal")):
        out[name] = ((states[:, k] > 0.5).astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
    pol = np.asarray([d[1] for d in data], dtype=np.float64)                      # [n,65]
    counts = np.rint(pol * 65535.0)                                                # fallback: 1/65535 quantisation
    todo = np.ones(n, bool)
    # any N with p*N integral reproduces the distribution exactly (counts / sum is the same correctly rounded quotient);
    # the usual simulation counts are tried first so that a campaign's samples resolve in one pass
    common = [50, 100, 25, 200, 400, 800, 10, 20, 30, 40, 60, 64, 128, 256, 512, 1000, 1600]
    for total in common + [t for t in range(1, 1025) if t not in common]:
        if not todo.any():
            break
        c = pol[todo] * total
        r = np.rint(c)
        good = (np.abs(c - r).max(axis=1) < 2e-3) & (r.sum(axis=1) == total)
        if good.any():
            idx = np.flatnonzero(todo)[good]
            counts[idx] = r[good]
            todo[idx] = False
    out["visits"] = counts.astype(np.uint16)
    out["n_children"] = (counts > 0).sum(axis=1).astype(np.uint8)
    out["value"] = np.rint([float(d[2]) for d in data]).astype(np.int8)
    out["game"] = -1
    return out


class ReplayBuffer:
    """Drop-in for `src.train.buffer.ReplayBuffer` (buffer.py:15-123), device-resident."""

    def __init__(self, max_size: int = 100000, ctx: Context | None = None):
        self.max_size = int(max_size)
        self.ctx = ctx or Context.default()
        h = C.c_void_p()
        check(self.ctx.lib.oth_replay_create(self.ctx.handle, self.max_size, C.byref(h)))
        self.handle = h

    # ---- filling ----------------------------------------------------------------------------------
    def add(self, data: List[Tuple[np.ndarray, np.ndarray, float]]) -> None:
        """buffer.py:33-45"""
        self.add_packed(pack_training_data(list(data)))

    def add_single(self, state: np.ndarray, policy: np.ndarray, value: float) -> None:
        """buffer.py:47-56"""
        self.add([(state, policy, value)])

    def add_packed(self, samples: np.ndarray) -> None:
        assert samples.dtype == _lib.SAMPLE_DTYPE
        samples = np.ascontiguousarray(samples)
        check(self.ctx.lib.oth_replay_add(self.handle, ptr(samples), samples.size, MEM_HOST))

    def add_from_worker(self, worker) -> int:
        """Append the last campaign of a ParallelSelfPlayWorker device-to-device (no host hop)."""
        dev_ptr, n = worker._engine.samples_device()
        if n:
            check(self.ctx.lib.oth_replay_add(self.handle, dev_ptr, n, MEM_DEVICE))
        return n

    def add_device(self, records, count: int) -> None:
        """Append `count` packed records held in a CUDA uint8 tensor (e.g. dist.all_gather_samples_device's result);
        the copy into the ring is ordered after whatever torch stream produced the tensor."""
        if count:
            with _lib.torch_order(self.ctx, records):
                check(self.ctx.lib.oth_replay_add(self.handle, ptr(records), int(count), MEM_DEVICE))

    # ---- sampling ---------------------------------------------------------------------------------
    def _draw(self, batch_size: int) -> np.ndarray:
        n = len(self)
        if n < batch_size:
            raise ValueError(f"Buffer size ({n}) is smaller than batch size ({batch_size})")    # buffer.py:71-74
        return np.asarray(random.sample(range(n), batch_size), np.int64)                        # buffer.py:78

    def sample(self, batch_size: int, augment: bool = False):
        """buffer.py:58-84 -> numpy (states [B,3,8,8], policies [B,65], values [B,1]).
        augment=True: every sample comes back as a random one of its 8 dihedral images (see `gather`)."""
        idx = self._draw(batch_size)
        sym = np.asarray([random.randrange(8) for _ in range(batch_size)], np.uint8) if augment else None
        return self.gather(idx, sym)

    def gather(self, idx, sym=None):
        """Expand the samples at logical indices `idx` (0 = oldest); `sym[i]` in 0..7 (optional) selects the dihedral image
        2*k + flip of sample i in the order of OthelloBitboard.get_symmetries (bitboard.pyx:338-370).  The images are bit
        permutations applied to the packed 168-byte records inside the gather kernel (csrc/replay.cu): this is the
        augmentation self_play.py:166-212 declares and never wires in."""
        idx = np.ascontiguousarray(idx, np.int64)
        n = int(idx.size)
        st = np.empty((n, 3, 8, 8), np.float32); po = np.empty((n, 65), np.float32); va = np.empty((n, 1), np.float32)
        if sym is None:
            check(self.ctx.lib.oth_replay_gather(self.handle, ptr(idx), n, ptr(st), ptr(po), ptr(va), MEM_HOST))
        else:
            sym = np.ascontiguousarray(sym, np.uint8)
            assert sym.size == n
            check(self.ctx.lib.oth_replay_gather_sym(self.handle, ptr(idx), ptr(sym), n, ptr(st), ptr(po), ptr(va), MEM_HOST))
        return st, po, va

    def sample_torch(self, batch_size: int, device=None, augment: bool = False):
        """Same minibatch as CUDA tensors, written by the gather kernel (for trainer.py:264-269)."""
        import torch
        dev = torch.device(device) if device is not None else torch.device("cuda", self.ctx.device)
        idx = torch.from_numpy(self._draw(batch_size)).to(dev)
        sym = torch.randint(0, 8, (batch_size,), dtype=torch.uint8, device=dev) if augment else None
        st = torch.empty((batch_size, 3, 8, 8), dtype=torch.float32, device=dev)
        po = torch.empty((batch_size, 65), dtype=torch.float32, device=dev)
        va = torch.empty((batch_size, 1), dtype=torch.float32, device=dev)
        with _lib.torch_order(self.ctx, idx, sym, st, po, va):     # ordered with torch's stream on the device, no host sync
            if sym is None:
                check(self.ctx.lib.oth_replay_gather(self.handle, ptr(idx), batch_size, ptr(st), ptr(po), ptr(va), MEM_DEVICE))
            else:
                check(self.ctx.lib.oth_replay_gather_sym(self.handle, ptr(idx), ptr(sym), batch_size, ptr(st), ptr(po), ptr(va), MEM_DEVICE))
        return st, po, va

    def check(self) -> None:
        """Raise if a device-side gather (`sample_torch`, device indices) used an index outside [0, len)."""
        check(self.ctx.lib.oth_replay_check(self.handle))

    # ---- bookkeeping --------------------------------------------------------------------------------
    def __len__(self) -> int:
        return int(self.ctx.lib.oth_replay_size(self.handle))

    def clear(self) -> None:
        check(self.ctx.lib.oth_replay_clear(self.handle))

    def is_ready(self, min_size: int) -> bool:
        return len(self) >= min_size

    def get_statistics(self) -> dict:
        """buffer.py:102-123"""
        n = len(self)
        if n == 0:
            return {"size": 0, "max_size": self.max_size, "fill_rate": 0.0, "value_mean": 0.0, "value_std": 0.0}
        m, s = C.c_double(0), C.c_double(0)
        check(self.ctx.lib.oth_replay_value_stats(self.handle, C.byref(m), C.byref(s)))
        return {"size": n, "max_size": self.max_size, "fill_rate": n / self.max_size, "value_mean": m.value, "value_std": s.value}

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.ctx.lib.oth_replay_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PrioritizedReplayBuffer(ReplayBuffer):
    """buffer.py:126-177: the reference's class is a stub that samples uniformly; so does this one."""

    def __init__(self, max_size: int = 100000, alpha: float = 0.6, ctx: Context | None = None):
        super().__init__(max_size, ctx)
        self.alpha = alpha
