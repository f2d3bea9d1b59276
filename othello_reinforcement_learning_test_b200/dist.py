"""Multi-GPU plumbing: one process per GPU, games sharded, no traffic inside the move loop.

Self-play games are independent, so a campaign of E episodes on W ranks is W independent
campaigns of ceil(E/W) episodes.  `torch.distributed` (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests) is used for exactly two things, both
outside the per-move loop:

  * `broadcast_weights(model, src)`   -- new network weights from the trainer rank;
  * `all_gather_samples(samples)`     -- packed trajectories (168 B records) to every rank,
                                         i.e. into the replay buffer's rank.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._lib import SAMPLE_DTYPE


def shard_episodes(num_episodes: int, rank: int, world_size: int) -> int:
    """Episodes this rank plays: contiguous, balanced, sums to num_episodes."""
    base, rem = divmod(int(num_episodes), int(world_size))
    return base + (1 if rank < rem else 0)


def _comm_device() -> torch.device:
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_weights(model: torch.nn.Module, src: int = 0) -> int:
    """Broadcast every parameter and buffer of `model` from rank `src` as ONE flat buffer
    (2.98 M floats for 10x128: a single latency-bound collective).  Returns the byte count."""
    tensors = [t for t in model.state_dict().values()]
    dev = _comm_device()
    flat = torch.cat([t.detach().reshape(-1).to(dev, torch.float32) for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].reshape(t.shape).to(t.dtype))
            off += n
    return flat.numel() * 4


def all_gather_samples(samples: np.ndarray) -> np.ndarray:
    """All-gather variable-length packed sample arrays; every rank gets the concatenation in
    rank order with GLOBALLY UNIQUE episode ids: rank r's `game` field is offset by the number of
    episodes of the ranks below it (every rank numbers its own episodes from 0), so grouping or
    sorting by (game, ply) downstream never interleaves two ranks' games."""
    assert samples.dtype == SAMPLE_DTYPE
    world = dist.get_world_size()
    dev = _comm_device()
    episodes = int(samples["game"].max()) + 1 if samples.size else 0
    count = torch.tensor([samples.size, episodes], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count)
    episodes_per_rank = [int(c[1].item()) for c in counts]
    counts = [int(c[0].item()) for c in counts]
    biggest = max(counts) if counts else 0
    rec = SAMPLE_DTYPE.itemsize
    buf = torch.zeros(max(biggest, 1) * rec, dtype=torch.uint8, device=dev)
    if samples.size:
        buf[: samples.size * rec] = torch.from_numpy(samples.view(np.uint8).reshape(-1).copy()).to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    parts, base = [], 0
    for g, c, e in zip(gathered, counts, episodes_per_rank):
        part = g[: c * rec].cpu().numpy().view(SAMPLE_DTYPE).copy()
        part["game"] += base
        base += e
        parts.append(part)
    return np.concatenate(parts) if parts else np.empty(0, SAMPLE_DTYPE)


class _DevView:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


_GAME_WORD = SAMPLE_DTYPE.fields["game"][1] // 4          # int32 index of the `game` field inside a record
_REC_WORDS = SAMPLE_DTYPE.itemsize // 4


class DeviceSampleGather:
    """All-gather of packed samples that already live in device memory, without a host hop and without staging copies.

    NCCL's all-gather wants equally sized contributions: every rank contributes the first `biggest` records of its engine's
    sample buffer (the buffer is at least that long; what lies behind a rank's own count is ignored), straight from the
    engine's memory, into a receive buffer that is kept from call to call (grow-only).  Returns one (uint8 CUDA tensor,
    count) segment per rank, in rank order, with globally unique episode ids when `episodes` is given; feed them to
    `ReplayBuffer.add_device`."""

    def __init__(self, device: torch.device):
        self.device = device
        self.recv = None

    def reserve(self, records_per_rank: int) -> None:
        """Size the receive buffer up front (setup time, like cudaMalloc) for `records_per_rank` records from every rank."""
        need = dist.get_world_size() * int(records_per_rank) * SAMPLE_DTYPE.itemsize
        if self.recv is None or self.recv.numel() < need:
            self.recv = None
            self.recv = torch.empty(need + 4096, dtype=torch.uint8, device=self.device)

    def gather(self, dev_ptr: int, count: int, capacity: int, episodes: int | None = None):
        """`capacity`: records the memory at dev_ptr is known to hold (>= count)."""
        world = dist.get_world_size()
        rec = SAMPLE_DTYPE.itemsize
        cnt = torch.tensor([count, -1 if episodes is None else episodes], dtype=torch.int64, device=self.device)
        counts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(counts, cnt)
        episodes_per_rank = [int(c[1].item()) for c in counts]
        counts = [int(c[0].item()) for c in counts]
        biggest = max(max(counts), 1)
        need = world * biggest * rec
        if self.recv is None or self.recv.numel() < need:
            self.recv = None                                  # release before growing
            self.recv = torch.empty(int(need * 1.05) + 4096, dtype=torch.uint8, device=self.device)
        out = self.recv[:need]
        if biggest <= capacity:
            mine = torch.as_tensor(_DevView(dev_ptr, biggest * rec), device=self.device)  # zero-copy view of the engine's buffer
        else:                                                 # this rank's buffer is shorter than the largest contribution: pad a copy
            mine = torch.zeros(biggest * rec, dtype=torch.uint8, device=self.device)
            if count:
                mine[: count * rec] = torch.as_tensor(_DevView(dev_ptr, count * rec), device=self.device)
        dist.all_gather_into_tensor(out, mine)
        if all(e >= 0 for e in episodes_per_rank):
            words = out.view(torch.int32).view(world, biggest, _REC_WORDS)
            base = 0
            for r in range(1, world):
                base += episodes_per_rank[r - 1]
                if counts[r]:
                    words[r, : counts[r], _GAME_WORD] += base
        return [(out[r * biggest * rec: r * biggest * rec + c * rec], c) for r, c in enumerate(counts)]


def all_gather_samples_device(dev_ptr: int, count: int, device: torch.device, episodes: int | None = None):
    """One-shot form of DeviceSampleGather: (uint8 CUDA tensor holding every rank's records back to back in rank order,
    total count).  Prefer a kept DeviceSampleGather in a loop (no re-allocation, no concatenation)."""
    segs = DeviceSampleGather(device).gather(dev_ptr, count, max(count, 1), episodes)
    return torch.cat([t for t, _ in segs]), sum(c for _, c in segs)


def renumber_games(samples_per_rank: list[np.ndarray]) -> np.ndarray:
    """Concatenate per-rank sample arrays giving every episode a globally unique id."""
    out, base = [], 0
    for s in samples_per_rank:
        s = s.copy()
        if s.size:
            s["game"] += base
            base = int(s["game"].max()) + 1
        out.append(s)
    return np.concatenate(out) if out else np.empty(0, SAMPLE_DTYPE)
