"""oracle/net_oracle.py -- TEST INFRASTRUCTURE ONLY.

fp32 eval-mode restatement of the reference network's forward pass
(src/model/net.py:182-205 and the four sub-modules at :15-136) as a pure
function of a ``state_dict``, plus a platform-stable synthetic ``state_dict``
generator used by the golden fixtures.

The arithmetic itself is PyTorch's (third-party; reference pins torch 2.9.1 in
uv.lock, this image has 2.11): parity at this boundary is tolerance-based, the
tolerance is stated in the tests.  Pinned against the reference's own
``OthelloResNet`` module by oracle/gen_golden.py (tests/golden/net_ref.npz).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, net.py:25

# conv gain of the synthetic fixture weights per network shape: the deep net is toned down so that its policy is not
# saturated (gain 1.0 gives p_max 0.99, where bf16 rounding noise alone moves single probabilities by > 1e-2)
SYNTH_GAIN = {(2, 32): 1.0, (5, 64): 1.0, (10, 128): 0.7}


def state_dict_keys(num_blocks: int):
    """Key order of the reference module's state_dict (SURVEY 8(a) R-NN)."""
    def bn(p):
        return [f"{p}.weight", f"{p}.bias", f"{p}.running_mean", f"{p}.running_var", f"{p}.num_batches_tracked"]
    keys = ["conv_block.conv.weight"] + bn("conv_block.bn")
    for i in range(num_blocks):
        keys += [f"res_blocks.{i}.conv1.weight"] + bn(f"res_blocks.{i}.bn1")
        keys += [f"res_blocks.{i}.conv2.weight"] + bn(f"res_blocks.{i}.bn2")
    keys += ["policy_head.conv.weight"] + bn("policy_head.bn") + ["policy_head.fc.weight", "policy_head.fc.bias"]
    keys += ["value_head.conv.weight"] + bn("value_head.bn")
    keys += ["value_head.fc1.weight", "value_head.fc1.bias", "value_head.fc2.weight", "value_head.fc2.bias"]
    return keys


def make_state_dict(num_blocks: int, num_filters: int, seed: int, gain: float = 1.0):
    """Synthetic weights from numpy's PCG64 (stable across platforms/torch versions).

    Non-trivial BN statistics so that BN folding mistakes show up.  Conv weights
    are scaled ~He so activations neither vanish nor explode over many blocks.
    """
    rng = np.random.default_rng(seed)
    Fn = num_filters
    sd = {}

    def conv(name, cout, cin, k):
        std = gain * np.sqrt(2.0 / (cin * k * k))
        sd[name] = torch.from_numpy((rng.standard_normal((cout, cin, k, k)) * std).astype(np.float32))

    def bn(p, c):
        sd[f"{p}.weight"] = torch.from_numpy(rng.uniform(0.6, 1.4, c).astype(np.float32))
        sd[f"{p}.bias"] = torch.from_numpy(rng.uniform(-0.3, 0.3, c).astype(np.float32))
        sd[f"{p}.running_mean"] = torch.from_numpy(rng.uniform(-0.2, 0.2, c).astype(np.float32))
        sd[f"{p}.running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32))
        sd[f"{p}.num_batches_tracked"] = torch.tensor(7, dtype=torch.long)

    def lin(p, cout, cin):
        sd[f"{p}.weight"] = torch.from_numpy((rng.standard_normal((cout, cin)) * np.sqrt(1.0 / cin)).astype(np.float32))
        sd[f"{p}.bias"] = torch.from_numpy(rng.uniform(-0.2, 0.2, cout).astype(np.float32))

    conv("conv_block.conv.weight", Fn, 3, 3); bn("conv_block.bn", Fn)
    for i in range(num_blocks):
        conv(f"res_blocks.{i}.conv1.weight", Fn, Fn, 3); bn(f"res_blocks.{i}.bn1", Fn)
        # second conv of a block is scaled down so the residual stream stays O(1)
        std = 0.5 * np.sqrt(2.0 / (Fn * 9))
        sd[f"res_blocks.{i}.conv2.weight"] = torch.from_numpy(
            (rng.standard_normal((Fn, Fn, 3, 3)) * std).astype(np.float32))
        bn(f"res_blocks.{i}.bn2", Fn)
    conv("policy_head.conv.weight", 2, Fn, 1); bn("policy_head.bn", 2); lin("policy_head.fc", 65, 128)
    conv("value_head.conv.weight", 1, Fn, 1); bn("value_head.bn", 1)
    lin("value_head.fc1", 256, 64); lin("value_head.fc2", 1, 256)
    assert list(sd.keys()) == state_dict_keys(num_blocks)
    return sd


def infer_shape(sd):
    """(num_blocks, num_filters) from the keys, as players.py:159-223 does."""
    nf = int(sd["conv_block.conv.weight"].shape[0])
    nb = len({k.split(".")[1] for k in sd if k.startswith("res_blocks.")})
    return nb, nf


def _bn(x, sd, p):
    return F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"], sd[f"{p}.weight"], sd[f"{p}.bias"],
                        training=False, eps=BN_EPS)


@torch.no_grad()
def forward_fp32(sd, x: torch.Tensor):
    """x: float32 [N,3,8,8] -> (log_probs [N,65], value [N,1]); net.py:182-205."""
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    nb, _ = infer_shape(sd)
    h = F.relu(_bn(F.conv2d(x, sd["conv_block.conv.weight"], padding=1), sd, "conv_block.bn"))   # net.py:27-31
    for i in range(nb):                                                                            # net.py:47-61
        t = F.relu(_bn(F.conv2d(h, sd[f"res_blocks.{i}.conv1.weight"], padding=1), sd, f"res_blocks.{i}.bn1"))
        t = _bn(F.conv2d(t, sd[f"res_blocks.{i}.conv2.weight"], padding=1), sd, f"res_blocks.{i}.bn2")
        h = F.relu(t + h)
    p = F.relu(_bn(F.conv2d(h, sd["policy_head.conv.weight"]), sd, "policy_head.bn"))             # net.py:83-96
    p = F.linear(p.flatten(1), sd["policy_head.fc.weight"], sd["policy_head.fc.bias"])
    logp = F.log_softmax(p, dim=1)
    v = F.relu(_bn(F.conv2d(h, sd["value_head.conv.weight"]), sd, "value_head.bn"))               # net.py:119-136
    v = F.relu(F.linear(v.flatten(1), sd["value_head.fc1.weight"], sd["value_head.fc1.bias"]))
    v = torch.tanh(F.linear(v, sd["value_head.fc2.weight"], sd["value_head.fc2.bias"]))
    return logp, v


def _round_bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


@torch.no_grad()
def forward_bf16_emulated(sd, x: torch.Tensor, fp32_residual: bool = False):
    """Model of the CUDA kernels' numerics on the CPU (for tolerance budgeting):
    BN folded into bf16-rounded conv weights + fp32 bias, bf16-rounded
    activations between layers, fp32 accumulation, fp32 heads."""
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    nb, _ = infer_shape(sd)

    def fold(cw, p):
        s = sd[f"{p}.weight"] / torch.sqrt(sd[f"{p}.running_var"] + BN_EPS)
        return sd[cw] * s.view(-1, 1, 1, 1), sd[f"{p}.bias"] - sd[f"{p}.running_mean"] * s

    w, b = fold("conv_block.conv.weight", "conv_block.bn")
    h = F.relu(F.conv2d(x, _round_bf16(w), padding=1) + b.view(1, -1, 1, 1))
    res = h
    h = _round_bf16(h)
    for i in range(nb):
        w, b = fold(f"res_blocks.{i}.conv1.weight", f"res_blocks.{i}.bn1")
        t = _round_bf16(F.relu(F.conv2d(h, _round_bf16(w), padding=1) + b.view(1, -1, 1, 1)))
        w, b = fold(f"res_blocks.{i}.conv2.weight", f"res_blocks.{i}.bn2")
        t = F.conv2d(t, _round_bf16(w), padding=1) + b.view(1, -1, 1, 1)
        res = F.relu(t + (res if fp32_residual else h))
        h = _round_bf16(res)
    w, b = fold("policy_head.conv.weight", "policy_head.bn")
    p = F.relu(F.conv2d(h, w) + b.view(1, -1, 1, 1))
    p = F.linear(p.flatten(1), sd["policy_head.fc.weight"], sd["policy_head.fc.bias"])
    logp = F.log_softmax(p, dim=1)
    w, b = fold("value_head.conv.weight", "value_head.bn")
    v = F.relu(F.conv2d(h, w) + b.view(1, -1, 1, 1))
    v = F.relu(F.linear(v.flatten(1), sd["value_head.fc1.weight"], sd["value_head.fc1.bias"]))
    v = torch.tanh(F.linear(v, sd["value_head.fc2.weight"], sd["value_head.fc2.bias"]))
    return logp, v


def boards_to_tensor(self_b, opp_b) -> torch.Tensor:
    from . import cref
    return torch.from_numpy(cref.tensor_input_batch(np.asarray(self_b, np.uint64), np.asarray(opp_b, np.uint64)))
