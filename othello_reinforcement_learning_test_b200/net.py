"""OthelloResNet: the torch module (weights, training) and its B200 inference engine.

* `OthelloResNet` has the parameter/buffer names and shapes of the reference module
  (src/model/net.py:139-205; state_dict keys listed in SURVEY.md 8(a) R-NN), so checkpoints
  written by the reference Trainer load unchanged and training stays plain PyTorch.
* `InferenceNet` owns the device-side copy used by self-play: eval-mode BatchNorm folded
  into bf16 conv weights, trunk on tcgen05 tensor cores, fused heads (csrc/net_tc.cu).
  `sync_from(module)` re-uploads only when the module's tensors changed (the Trainer
  mutates the model in place between iterations, trainer.py:258,300-326).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import MEM_DEVICE, MEM_HOST, Context, check, ptr


class _ConvBN(nn.Module):
    def __init__(self, cin: int, cout: int, k: int):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class _ResBlock(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(c)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + x)


class _PolicyHead(nn.Module):
    def __init__(self, c: int, board: int):
        super().__init__()
        self.board_size = board
        self.conv = nn.Conv2d(c, 2, 1, bias=False)
        self.bn = nn.BatchNorm2d(2)
        self.fc = nn.Linear(2 * board * board, board * board + 1)

    def forward(self, x):
        x = F.relu(self.bn(self.conv(x))).flatten(1)       # channel-major flatten (net.py:90)
        return F.log_softmax(self.fc(x), dim=1)


class _ValueHead(nn.Module):
    def __init__(self, c: int, board: int, hidden: int = 256):
        super().__init__()
        self.board_size = board
        self.conv = nn.Conv2d(c, 1, 1, bias=False)
        self.bn = nn.BatchNorm2d(1)
        self.fc1 = nn.Linear(board * board, hidden)
        self.fc2 = nn.Linear(hidden, 1)

    def forward(self, x):
        x = F.relu(self.bn(self.conv(x))).flatten(1)
        return torch.tanh(self.fc2(F.relu(self.fc1(x))))


class OthelloResNet(nn.Module):
    """Dual-head ResNet; forward(x[B,3,8,8]) -> (log_probs[B,65], value[B,1])."""

    def __init__(self, num_blocks: int = 10, num_filters: int = 128, board_size: int = 8):
        super().__init__()
        if board_size != 8:
            raise ValueError("the board is 8x8 (the reference hard-wires it too, bitboard.pyx:153,309)")
        self.num_blocks, self.num_filters, self.board_size = num_blocks, num_filters, board_size
        # construction order == the reference's, so torch.manual_seed(s) gives the same init
        self.conv_block = _ConvBN(3, num_filters, 3)
        self.res_blocks = nn.ModuleList([_ResBlock(num_filters) for _ in range(num_blocks)])
        self.policy_head = _PolicyHead(num_filters, board_size)
        self.value_head = _ValueHead(num_filters, board_size)

    def forward(self, x):
        x = self.conv_block(x)
        for blk in self.res_blocks:
            x = blk(x)
        return self.policy_head(x), self.value_head(x)

    def predict(self, board_tensor):
        """net.py:207-236: probabilities + value, eval mode, no grad."""
        single = board_tensor.dim() == 3
        if single:
            board_tensor = board_tensor.unsqueeze(0)
        self.eval()
        with torch.no_grad():
            logp, v = self.forward(board_tensor)
            p = torch.exp(logp)
        return (p.squeeze(0), v.squeeze(0)) if single else (p, v)

    def get_param_count(self):
        total = sum(p.numel() for p in self.parameters())
        return {"total": total, "trainable": sum(p.numel() for p in self.parameters() if p.requires_grad)}


def create_model(config: dict) -> OthelloResNet:
    """net.py:245-265"""
    return OthelloResNet(num_blocks=config.get("num_blocks", 10), num_filters=config.get("num_filters", 128),
                         board_size=config.get("board_size", 8))


# ---------------------------------------------------------------------------------------------

def infer_shape(state_dict) -> tuple[int, int]:
    """(num_blocks, num_filters) from the keys, like MCTSPlayer.from_checkpoint (players.py:159-223)."""
    nf = int(state_dict["conv_block.conv.weight"].shape[0])
    nb = len({k.split(".")[1] for k in state_dict if k.startswith("res_blocks.")})
    return nb, nf


def is_resnet_state_dict(sd) -> bool:
    need = ("conv_block.conv.weight", "policy_head.fc.weight", "value_head.fc2.bias")
    return all(k in sd for k in need)


def flatten_state_dict(sd) -> np.ndarray:
    """Concatenate the floating-point tensors in the reference key order -> float32 host vector."""
    nb, _ = infer_shape(sd)

    def bn(p):
        return [f"{p}.weight", f"{p}.bias", f"{p}.running_mean", f"{p}.running_var"]
    keys = ["conv_block.conv.weight"] + bn("conv_block.bn")
    for i in range(nb):
        keys += [f"res_blocks.{i}.conv1.weight"] + bn(f"res_blocks.{i}.bn1")
        keys += [f"res_blocks.{i}.conv2.weight"] + bn(f"res_blocks.{i}.bn2")
    keys += ["policy_head.conv.weight"] + bn("policy_head.bn") + ["policy_head.fc.weight", "policy_head.fc.bias"]
    keys += ["value_head.conv.weight"] + bn("value_head.bn")
    keys += ["value_head.fc1.weight", "value_head.fc1.bias", "value_head.fc2.weight", "value_head.fc2.bias"]
    parts = [sd[k].detach().to("cpu", torch.float32).reshape(-1) for k in keys]
    return torch.cat(parts).contiguous().numpy()


class InferenceNet:
    """Device-side network (oth_net).  forward() takes bitboards, not planes."""

    def __init__(self, num_blocks: int, num_filters: int, ctx: Context | None = None, engine: str | None = None):
        self.ctx = ctx or Context.default()
        self.num_blocks, self.num_filters = int(num_blocks), int(num_filters)
        h = C.c_void_p()
        check(self.ctx.lib.oth_net_create(self.ctx.handle, self.num_blocks, self.num_filters, C.byref(h)))
        self.handle = h
        self._version_key = None
        if engine is not None:
            self.set_engine(engine)

    @classmethod
    def from_module(cls, module, ctx: Context | None = None, engine: str | None = None) -> "InferenceNet":
        sd = module.state_dict()
        nb, nf = infer_shape(sd)
        net = cls(nb, nf, ctx, engine)
        net.sync_from(module, force=True)
        return net

    def set_engine(self, engine: str) -> None:
        code = {"tcgen05": _lib.ENGINE_TCGEN05, "simt": _lib.ENGINE_SIMT, "tcgen05_pair": _lib.ENGINE_TCGEN05_PAIR}[engine]
        check(self.ctx.lib.oth_net_set_engine(self.handle, code))

    @property
    def engine(self) -> str:
        """The engine that really runs: "tcgen05" (64 / 128 filters), "tcgen05_pair" or the validation engine "simt"."""
        code = int(self.ctx.lib.oth_net_engine(self.handle))
        return {_lib.ENGINE_TCGEN05: "tcgen05", _lib.ENGINE_SIMT: "simt", _lib.ENGINE_TCGEN05_PAIR: "tcgen05_pair"}[code]

    def load_state_dict(self, sd) -> None:
        flat = flatten_state_dict(sd)
        check(self.ctx.lib.oth_net_load_weights(self.handle, ptr(flat), flat.size))

    def sync_from(self, module, force: bool = False) -> bool:
        """Re-fold + upload if any parameter/buffer of `module` changed since the last sync."""
        sd = module.state_dict()
        key = tuple((k, v._version, v.data_ptr()) for k, v in sd.items())
        if not force and key == self._version_key:
            return False
        self.load_state_dict(sd)
        self._version_key = key
        return True

    def forward(self, self_b, opp_b, out: str = "logprobs"):
        """-> (policy [n,65], value [n]); numpy in -> numpy out, CUDA tensors in -> CUDA tensors out."""
        kind = {"logprobs": _lib.OUT_LOGPROBS, "probs": _lib.OUT_PROBS, "priors": _lib.OUT_PRIORS}[out]
        on_dev = hasattr(self_b, "is_cuda") and self_b.is_cuda
        if on_dev:
            s, o = self_b.contiguous(), opp_b.contiguous()
            n = s.numel()
            pol = torch.empty((n, 65), dtype=torch.float32, device=s.device)
            val = torch.empty((n,), dtype=torch.float32, device=s.device)
            mem = MEM_DEVICE
        else:
            s = np.ascontiguousarray(self_b, np.uint64); o = np.ascontiguousarray(opp_b, np.uint64)
            n = s.size
            pol = np.empty((n, 65), np.float32); val = np.empty((n,), np.float32)
            mem = MEM_HOST
        with _lib.torch_order(self.ctx, s, o, pol, val):      # CUDA tensors: ordered with torch's current stream
            check(self.ctx.lib.oth_net_forward(self.handle, ptr(s), ptr(o), n, ptr(pol), ptr(val), kind, mem))
        return pol, val

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.ctx.lib.oth_net_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
