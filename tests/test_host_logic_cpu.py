"""CPU suite: host-side logic of the package (no device work) + the device header compiled for the
host (tests/hostcheck) against the oracle."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from oracle import cref, net_oracle
from oth_testutil import random_boards

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostcheck():
    so = os.path.join(tempfile.gettempdir(), f"oth_hostcheck_{os.getpid()}.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so,
                    os.path.join(ROOT, "tests", "hostcheck", "bitboard_hostcheck.cpp")], check=True)
    L = C.CDLL(so)
    u64 = np.ctypeslib.ndpointer(np.uint64); i32 = np.ctypeslib.ndpointer(np.int32)
    u8 = np.ctypeslib.ndpointer(np.uint8); i8 = np.ctypeslib.ndpointer(np.int8)
    L.hc_legal.argtypes = [u64, u64, u64, C.c_int64]
    L.hc_flips.argtypes = [u64, u64, i32, u64, C.c_int64]
    L.hc_make_move.argtypes = [u64, u64, i32, i32, u8, C.c_int64]
    L.hc_terminal_winner.argtypes = [u64, u64, u8, i8, C.c_int64]
    L.hc_nth_set_bit.argtypes = [C.c_uint64, C.c_int]
    L.hc_playout.argtypes = [C.c_uint64, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    yield L
    os.unlink(so)


def test_branch_free_rules_equal_the_oracle(hostcheck, golden_games, golden_moves65):
    L = hostcheck
    S, O = golden_games["self_b"], golden_games["opp_b"]
    out = np.empty_like(S); L.hc_legal(S, O, out, S.size)
    assert np.array_equal(out, golden_games["legal"])
    s, o = random_boards(300000, 1)
    out = np.empty_like(s); L.hc_legal(s, o, out, s.size)
    assert np.array_equal(out, cref.legal_batch(s, o))
    pos = np.random.default_rng(2).integers(0, 64, s.size).astype(np.int32)
    f = np.empty_like(s); L.hc_flips(s, o, pos, f, s.size)
    assert np.array_equal(f, cref.flips_batch(s, o, pos))
    t = np.empty(s.size, np.uint8); w = np.empty(s.size, np.int8)
    L.hc_terminal_winner(s, o, t, w, s.size)
    t2, w2 = cref.terminal_winner_batch(s, o)
    assert np.array_equal(t, t2) and np.array_equal(w, w2)
    m = golden_moves65
    a = m["self_b"].copy(); b = m["opp_b"].copy(); mc = m["move_count"].copy(); ok = np.empty(a.size, np.uint8)
    L.hc_make_move(a, b, mc, m["action"], ok, a.size)
    assert np.array_equal(ok, m["ok"]) and np.array_equal(a, m["self_after"]) and np.array_equal(b, m["opp_after"])
    assert np.array_equal(mc, m["move_count_after"])


def test_nth_set_bit_and_playout_rng(hostcheck):
    L = hostcheck
    rng = np.random.default_rng(5)
    for _ in range(2000):
        m = int(rng.integers(1, 2**63)) | (int(rng.integers(0, 2)) << 63)
        bits = [i for i in range(64) if (m >> i) & 1]
        k = int(rng.integers(0, len(bits)))
        assert L.hc_nth_set_bit(m, k) == bits[k]
    r = cref.random_playouts(500, 99, threads=1, want_finals=True)
    for g in range(500):
        fs, fo = C.c_uint64(), C.c_uint64()
        p = L.hc_playout(99, g, C.byref(fs), C.byref(fo))
        assert (p, fs.value, fo.value) == (int(r["plies"][g]), int(r["final_self"][g]), int(r["final_opp"][g]))


def test_module_has_the_reference_state_dict_and_init(golden_net):
    from othello_reinforcement_learning_test_b200.net import OthelloResNet, flatten_state_dict, infer_shape
    for nb, nf in [(5, 64), (10, 128)]:
        torch.manual_seed(42)
        m = OthelloResNet(nb, nf)
        sd = m.state_dict()
        assert list(sd.keys()) == net_oracle.state_dict_keys(nb)
        assert infer_shape(sd) == (nb, nf)
        chk = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in sd.values()])
        assert np.allclose(chk, golden_net[f"init42_{nb}x{nf}_checksum"], rtol=1e-12, atol=1e-12), \
            "torch.manual_seed(42) must give the reference's initial weights"
        flat = flatten_state_dict(sd)
        n_tracked = sum(1 for k in sd if k.endswith("num_batches_tracked"))
        assert flat.size == sum(v.numel() for v in sd.values()) - n_tracked
    assert sum(p.numel() for p in OthelloResNet(10, 128).parameters()) == 2_983_624
    # forward agrees with the fp32 oracle and with the reference module's recorded outputs
    torch.manual_seed(42)
    m = OthelloResNet(5, 64).eval()
    x = net_oracle.boards_to_tensor(golden_net["self_b"], golden_net["opp_b"])
    with torch.no_grad():
        lp, v = m(x)
    assert lp.shape == (x.shape[0], 65) and v.shape == (x.shape[0], 1)
    assert np.allclose(lp.numpy(), golden_net["init42_5x64_logp"], atol=2e-5)
    assert np.allclose(v.numpy().reshape(-1), golden_net["init42_5x64_value"], atol=2e-5)


def test_sample_expansion_matches_the_reference_format(golden_games):
    from othello_reinforcement_learning_test_b200 import _lib
    from othello_reinforcement_learning_test_b200.self_play import planes_from_bits, samples_to_training_data
    S, O, L = golden_games["self_b"][:500], golden_games["opp_b"][:500], golden_games["legal"][:500]
    assert np.array_equal(planes_from_bits(S, O, L), cref.tensor_input_batch(S, O))
    smp = np.zeros(6, _lib.SAMPLE_DTYPE)
    smp["game"] = [1, 0, 1, 0, 0, 1]; smp["ply"] = [1, 2, 0, 0, 1, 2]
    smp["self_b"] = S[:6]; smp["opp_b"] = O[:6]; smp["legal"] = L[:6]
    smp["visits"][:, 19] = 30; smp["visits"][:, 26] = 20
    smp["value"] = [1, -1, 0, 1, -1, 1]
    data = samples_to_training_data(smp)
    assert len(data) == 6
    order = [3, 4, 1, 2, 0, 5]
    for (st, pol, val), i in zip(data, order):
        assert st.shape == (3, 8, 8) and st.dtype == np.float32 and pol.dtype == np.float32 and isinstance(val, float)
        assert np.array_equal(st, cref.tensor_input_batch(S[i:i + 1], O[i:i + 1])[0])
        assert pol[19] == np.float32(0.6) and pol[26] == np.float32(0.4) and abs(pol.sum() - 1) < 1e-6
        assert val == float(smp["value"][i])


def test_dropin_routes_the_reference_import_paths():
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200 import dropin
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    try:
        for k in saved:
            del sys.modules[k]
        dropin.install()
        from src.cython.bitboard import OthelloBitboard
        from src.mcts.mcts import MCTS
        from src.train.parallel_self_play import BatchMCTS, create_parallel_self_play_worker
        from src.train.self_play import SelfPlayWorker
        from src.model.net import OthelloResNet
        assert OthelloBitboard is pkg.OthelloBitboard and MCTS is pkg.MCTS and BatchMCTS is pkg.BatchMCTS
        assert SelfPlayWorker is pkg.SelfPlayWorker and OthelloResNet is pkg.OthelloResNet
        w = create_parallel_self_play_worker({"mcts": {"num_simulations": 50, "c_puct": 1.5},
                                              "self_play": {"temperature_threshold": 20, "num_parallel_games": 16}},
                                             OthelloResNet(1, 16), "cuda")
        assert (w.num_simulations, w.temperature_threshold, w.num_parallel_games, w.batch_mcts.c_puct) == (50, 20, 16, 1.5)
        m = MCTS(None, "cpu", c_puct=1.5)
        assert (m.model, m.device, m.c_puct, m.dirichlet_alpha, m.dirichlet_epsilon) == (None, "cpu", 1.5, 0.3, 0.25)
        # the rows SURVEY 8(f) marks "next" are opt-in: not routed by default ...
        assert "src.train.buffer" not in sys.modules and "src.eval.arena" not in sys.modules
        dropin.uninstall()
        dropin.install(replay_buffer=True, arena=True)
        from src.train.buffer import ReplayBuffer
        from src.eval.arena import Arena, BatchArena, MatchResult, evaluate_player
        assert Arena is BatchArena and hasattr(Arena, "play_game") and hasattr(Arena, "play_matches")
        from src.eval.players import RandomPlayer, GreedyPlayer, MCTSPlayer
        assert ReplayBuffer is pkg.ReplayBuffer and BatchArena is pkg.BatchArena and MatchResult is pkg.MatchResult
        assert RandomPlayer is pkg.RandomPlayer and GreedyPlayer is pkg.GreedyPlayer and MCTSPlayer is pkg.MCTSPlayer
        assert evaluate_player is pkg.evaluate_player
    finally:
        dropin.uninstall()
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_board_object_host_side_behaviour():
    import othello_reinforcement_learning_test_b200 as pkg
    b = pkg.OthelloBitboard()
    assert (b.self_board, b.opp_board, b.move_count, b.passed) == (cref.START_SELF, cref.START_OPP, 0, False)
    c = b.copy()
    c.self_board = 123; c.move_count = 7
    assert b.self_board == cref.START_SELF and b.move_count == 0 and c.self_board == 123
    assert "● " in repr(b) and b.to_string() == repr(b) and repr(b).splitlines()[0] == "  A B C D E F G H"


_DIST_SCRIPT = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from othello_reinforcement_learning_test_b200 import _lib, dist as odist
from othello_reinforcement_learning_test_b200.net import OthelloResNet
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
assert [odist.shard_episodes(101, r, 2) for r in range(2)] == [51, 50]
torch.manual_seed(100 + rank)
m = OthelloResNet(1, 16)
with torch.no_grad():
    m.conv_block.bn.running_mean.add_(rank + 1.0)
nbytes = odist.broadcast_weights(m, src=0)
torch.manual_seed(100)
ref = OthelloResNet(1, 16)
with torch.no_grad():
    ref.conv_block.bn.running_mean.add_(1.0)
for (k, a), (_, b) in zip(m.state_dict().items(), ref.state_dict().items()):
    assert torch.equal(a, b), k
smp = np.zeros(3 + 2 * rank, _lib.SAMPLE_DTYPE)
smp["game"] = np.arange(smp.size) % 2
smp["ply"] = rank * 10 + np.arange(smp.size)
smp["visits"][:, rank] = 7
allv = odist.all_gather_samples(smp)
assert allv.size == 3 + 5
assert (allv["ply"][:3] == [0, 1, 2]).all() and (allv["ply"][3:] == [10, 11, 12, 13, 14]).all()
assert (allv["visits"][:3, 0] == 7).all() and (allv["visits"][3:, 1] == 7).all()
# episode ids are globally unique after the gather: rank 1's games {0,1} follow rank 0's {0,1}
assert allv["game"][:3].tolist() == [0, 1, 0] and allv["game"][3:].tolist() == [2, 3, 2, 3, 2]
glob = odist.renumber_games([smp, smp])
assert sorted(set(glob["game"].tolist())) == [0, 1, 2, 3]
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", nbytes)
"""


def test_multi_process_plumbing_gloo_world_size_2():
    """The N>1 path (weight broadcast + trajectory all-gather) on CPU with gloo."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [subprocess.Popen([sys.executable, "-c", _DIST_SCRIPT, ROOT, str(port), str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_pack_training_data_round_trips_exactly(golden_games):
    from oracle.gen_golden import replay_fixture_data
    from othello_reinforcement_learning_test_b200.buffer import pack_training_data
    from othello_reinforcement_learning_test_b200.self_play import planes_from_bits
    data = replay_fixture_data(golden_games)[:200]
    packed = pack_training_data(data)
    st = planes_from_bits(packed["self_b"], packed["opp_b"], packed["legal"])
    counts = packed["visits"].astype(np.float32)
    pol = counts / counts.sum(axis=1, keepdims=True, dtype=np.float32)
    for i, (s, p, v) in enumerate(data):
        assert np.array_equal(st[i], s) and np.array_equal(pol[i], p) and float(packed["value"][i]) == v


def test_mcts_player_from_checkpoint_reads_the_network_shape(tmp_path, capsys):
    """players.py:159-222: blocks / filters come from the checkpoint's state_dict; the weights are loaded into our module."""
    import torch
    import othello_reinforcement_learning_test_b200 as pkg
    torch.manual_seed(3)
    src = pkg.OthelloResNet(3, 64)
    path = tmp_path / "ckpt.pt"
    torch.save({"model_state_dict": src.state_dict(), "config": {"num_blocks": 99, "num_filters": 7}}, path)
    player = pkg.MCTSPlayer.from_checkpoint(str(path), "cuda", num_simulations=25)
    assert "blocks=3, filters=64" in capsys.readouterr().out
    assert player.name == "MCTS-AI-25sim" and player.num_simulations == 25 and not player.model.training
    got = player.model.state_dict()
    assert all(torch.equal(got[k], v) for k, v in src.state_dict().items())
