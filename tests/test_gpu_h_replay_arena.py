"""GPU parity for the rows SURVEY.md 8(f) marks "next": device-resident replay buffer and batched arena,
against fixtures produced by the reference's own ReplayBuffer / Arena (oracle/gen_golden.py)."""
import os
import random

import numpy as np
import pytest

from oracle import cref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_replay_buffer_minibatches_equal_the_reference(golden_games):
    import torch
    import othello_reinforcement_learning_test_b200 as pkg
    from oracle.gen_golden import replay_fixture_data
    g = dict(np.load(os.path.join(GOLDEN, "replay_ref.npz")))
    data = replay_fixture_data(golden_games)
    buf = pkg.ReplayBuffer(max_size=500)
    assert len(buf) == 0 and not buf.is_ready(1) and buf.get_statistics()["size"] == 0
    with pytest.raises(ValueError):
        buf.sample(1)                                            # buffer.py:71-74
    buf.add(data[:300]); buf.add(data[300:])                     # 700 samples into 500 slots: the oldest 200 fall out
    assert len(buf) == 500 and buf.is_ready(500)
    random.seed(5)
    st, po, va = buf.sample(48)                                  # random.sample, like buffer.py:78
    assert st.shape == (48, 3, 8, 8) and po.shape == (48, 65) and va.shape == (48, 1)
    assert np.array_equal(st, g["states"]) and np.array_equal(po, g["policies"]) and np.array_equal(va, g["values"])
    random.seed(6)
    st2, po2, va2 = buf.sample(500)
    assert np.array_equal(st2.sum(axis=0), g["all_states_sum"]) and np.array_equal(po2, g["all_policies"])
    assert np.array_equal(va2, g["all_values"])
    s = buf.get_statistics()
    assert [s["size"], s["max_size"], s["fill_rate"]] == g["stats"][:3].tolist()
    assert abs(s["value_mean"] - g["stats"][3]) < 1e-12 and abs(s["value_std"] - g["stats"][4]) < 1e-12
    random.seed(5)
    ts, tp, tv = buf.sample_torch(48)                            # same minibatch, straight into CUDA tensors
    assert ts.is_cuda and np.array_equal(ts.cpu().numpy(), g["states"]) and np.array_equal(tp.cpu().numpy(), g["policies"])
    assert np.array_equal(tv.cpu().numpy(), g["values"])
    buf.clear()
    assert len(buf) == 0
    buf.add_single(*data[0])
    assert len(buf) == 1 and np.array_equal(buf.sample(1)[1][0], data[0][1])


def test_replay_buffer_takes_a_campaign_device_to_device():
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.self_play import samples_to_training_data
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", num_simulations=10, temperature_threshold=15,
                                   num_parallel_games=8, seed=4, verbose=False)
    packed = w.execute_episodes_packed(16)
    buf = pkg.ReplayBuffer(max_size=10000)
    n = buf.add_from_worker(w)
    assert n == packed.size == len(buf)
    random.seed(1)
    idx = random.sample(range(n), 64)
    random.seed(1)
    st, po, va = buf.sample(64)
    want = samples_to_training_data(packed)                     # sorted by (game, ply); the device order is flush order
    by_key = {(int(p["game"]), int(p["ply"])): i for i, p in enumerate(packed[np.lexsort((packed["ply"], packed["game"]))])}
    for k, i in enumerate(idx):
        j = by_key[(int(packed[i]["game"]), int(packed[i]["ply"]))]
        assert np.array_equal(st[k], want[j][0]) and np.array_equal(po[k], want[j][1]) and va[k, 0] == want[j][2]


def test_batched_arena_equals_the_reference_arena():
    import othello_reinforcement_learning_test_b200 as pkg
    g = dict(np.load(os.path.join(GOLDEN, "arena_ref.npz")))

    def pack(results):
        return np.array([[r.winner, r.player1_score, r.player2_score, r.num_moves] for r in results], np.int32)
    arena = pkg.BatchArena()
    assert np.array_equal(pack(arena.play_matches(pkg.GreedyPlayer("A"), pkg.GreedyPlayer("B"), num_games=4)), g["greedy_vs_greedy"])
    for sims in (8, 25):
        mp = pkg.MCTSPlayer(None, "cuda", num_simulations=sims)              # model None = the integer test evaluator
        assert np.array_equal(pack(arena.play_matches(mp, pkg.GreedyPlayer("G"), num_games=4)), g[f"mcts{sims}_vs_greedy"])
        assert np.array_equal(pack(arena.play_matches(pkg.GreedyPlayer("G"), mp, num_games=2, alternate_colors=False)),
                              g[f"greedy_vs_mcts{sims}"])
    ev = pkg.evaluate_player(pkg.GreedyPlayer("A"), pkg.GreedyPlayer("B"), num_games=6)
    assert [ev["win_rate"], ev["avg_score"], ev["avg_moves"]] == g["evaluate_greedy"].tolist()
    assert str(ev["results"][0]).startswith(("A wins", "B wins", "Draw"))


def test_greedy_and_random_choices_against_the_oracle(golden_games):
    import othello_reinforcement_learning_test_b200 as pkg
    live = np.flatnonzero(golden_games["terminal"] == 0)
    S, O, MC = golden_games["self_b"][live], golden_games["opp_b"][live], golden_games["move_count"][live]
    act = pkg.GreedyPlayer().get_actions(S, O, MC, np.arange(S.size))
    rnd = pkg.RandomPlayer(seed=3).get_actions(S, O, MC, np.arange(S.size))
    hist = {}
    for i in range(0, S.size, 7):
        s, o, mc = int(S[i]), int(O[i]), int(MC[i])
        legal = cref.legal_list(s, o)
        best, best_score = legal[0], -1                            # players.py:87-113 restated
        if legal != [64]:
            for a in legal:
                ok, s2, o2, _ = cref.make_move(s, o, mc, a)
                score = bin(o2).count("1") if mc % 2 == 0 else bin(s2).count("1")
                if score > best_score:
                    best, best_score = a, score
        else:
            best = 64
        assert act[i] == best, i
        assert int(rnd[i]) in legal
        hist[len(legal)] = hist.get(len(legal), 0) + 1
    # a whole match against the random player finishes, with plausible scores
    res = pkg.BatchArena().play_matches(pkg.MCTSPlayer(None, "cuda", num_simulations=5), pkg.RandomPlayer(seed=9), num_games=32)
    assert len(res) == 32 and all(r.player1_score + r.player2_score <= 64 and 9 <= r.num_moves <= 130 for r in res)
    assert len({(r.player1_score, r.num_moves) for r in res}) > 8     # the random opponent really varies


def test_arena_play_game_is_one_game_of_a_match(ctx, capsys):
    """Arena.play_game (arena.py:68-162): the single-game entry point agrees with play_matches for both colour assignments."""
    import othello_reinforcement_learning_test_b200 as pkg
    a, b = pkg.GreedyPlayer(name="A"), pkg.GreedyPlayer(name="B")          # deterministic players
    arena = pkg.Arena(verbose=False)
    pair = arena.play_matches(a, b, num_games=2, alternate_colors=True)     # game 0: A starts, game 1: B starts
    for starting, want in ((1, pair[0]), (-1, pair[1])):
        got = arena.play_game(a, b, starting_player=starting)
        assert (got.winner, got.player1_score, got.player2_score, got.num_moves) == \
               (want.winner, want.player1_score, want.player2_score, want.num_moves)
        assert (got.player1_name, got.player2_name) == ("A", "B")
    pkg.Arena(verbose=True).play_matches(a, b, num_games=2)
    out = capsys.readouterr().out
    assert "Match Summary" in out and "Total Games: 2" in out and "Average Moves:" in out


def test_device_side_symmetry_augmentation_equals_the_reference_get_symmetries(golden_symmetry):
    """the 8 dihedral images as bit permutations of the packed records inside the gather kernel, against
    what the compiled reference's get_symmetries (bitboard.pyx:338-370) returned (tests/golden/symmetry_ref.npz)."""
    import torch
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200 import _lib
    g = golden_symmetry
    n = g["self_b"].size
    rec = np.zeros(n, _lib.SAMPLE_DTYPE)
    rec["self_b"], rec["opp_b"] = g["self_b"], g["opp_b"]
    rec["legal"] = cref.legal_batch(g["self_b"], g["opp_b"])
    rec["visits"][:, :64] = np.arange(1, 65, dtype=np.uint16)[None, :]        # a distinct count on every square
    rec["visits"][:, 64] = 100
    rec["value"] = np.where(np.arange(n) % 2 == 0, 1, -1)
    total = np.float32(rec["visits"][0].astype(np.int64).sum())
    buf = pkg.ReplayBuffer(max_size=n)
    buf.add_packed(rec)
    w8 = (1 << np.arange(64, dtype=np.uint64))
    for t in range(8):
        st, po, va = buf.gather(np.arange(n), np.full(n, t, np.uint8))
        bits = ((st.reshape(n, 3, 64) > 0.5).astype(np.uint64) * w8).sum(axis=2, dtype=np.uint64)
        assert np.array_equal(bits, g["planes"][:, t]), t                      # planes: bit-exact with the reference's images
        for i in range(n):
            # where did the reference move square s's probability?  pi has distinct values, so the permutation is readable
            src = np.array([int(np.flatnonzero(g["pi"][i, :64] == v)[0]) for v in g["policy"][i, t, :64]])
            assert np.array_equal(po[i, :64], (src + 1).astype(np.float32) / total), (t, i)
            assert po[i, 64] == np.float32(100) / total and g["policy"][i, t, 64] == g["pi"][i, 64]
        assert np.array_equal(va[:, 0], rec["value"].astype(np.float32))
    # identity image == plain gather; device path (CUDA tensors) == host path
    a = buf.gather(np.arange(n), np.zeros(n, np.uint8)); b = buf.gather(np.arange(n))
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    random.seed(3)
    ts, tp, tv = buf.sample_torch(32, augment=True)
    assert ts.shape == (32, 3, 8, 8) and torch.allclose(tp.sum(dim=1), torch.ones(32, device=tp.device), atol=1e-6)
    assert ((ts[:, 0] + ts[:, 1]) <= 1).all()                                  # still a board: no square owned twice
    buf.check()
    # a bad device-side index is flagged instead of reading a stale ring slot
    from othello_reinforcement_learning_test_b200._lib import MEM_DEVICE, check, ptr
    idx = torch.tensor([0, n + 5], dtype=torch.int64, device="cuda")
    o1 = torch.empty((2, 3, 8, 8), device="cuda"); o2 = torch.empty((2, 65), device="cuda"); o3 = torch.empty((2, 1), device="cuda")
    torch.cuda.synchronize()
    check(buf.ctx.lib.oth_replay_gather(buf.handle, ptr(idx), 2, ptr(o1), ptr(o2), ptr(o3), MEM_DEVICE))
    with pytest.raises(pkg.OthelloB200Error):
        buf.check()
    buf.check()                                                                 # the flag is cleared by the report


def test_single_board_step_is_one_call_and_matches_the_oracle(golden_games):
    """oth_board_step: make_move + legal mask + terminal + winner + counts of the position reached in ONE launch, against
    the oracle on reference game positions (accepts, rejects, passes, terminal positions); latency is printed."""
    import ctypes as C
    import time
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200 import _lib
    ctx = pkg.Context.default(0)
    g = golden_games
    rng = np.random.default_rng(12)
    pick = rng.choice(g["self_b"].size, 400, replace=False)
    st = _lib.BoardState()
    for i in pick:
        s, o, mc = int(g["self_b"][i]), int(g["opp_b"][i]), int(g["move_count"][i])
        for a in (_lib.ACTION_NONE, int(g["action"][i]), int(rng.integers(0, 65)), -1, 99):
            launches = ctx.launch_count
            _lib.check(ctx.lib.oth_board_step(ctx.handle, s, o, mc, a, C.byref(st)))
            assert ctx.launch_count == launches + 1
            if a == _lib.ACTION_NONE:
                ok, s2, o2, mc2 = True, s, o, mc
            else:
                ok, s2, o2, mc2 = cref.make_move(s, o, mc, a)
                if not ok:
                    s2, o2, mc2 = s, o, mc
            assert (bool(st.ok), int(st.self_b), int(st.opp_b), int(st.move_count)) == (bool(ok), s2, o2, mc2), (i, a)
            assert int(st.legal) == cref.legal(s2, o2) and bool(st.terminal) == cref.is_terminal(s2, o2)
            assert int(st.winner) == cref.winner(s2, o2)
            assert (int(st.self_count), int(st.opp_count)) == (bin(s2).count("1"), bin(o2).count("1"))
    # the board class: one launch per move, everything else answered from that result
    b = pkg.OthelloBitboard()
    b.get_legal_moves()
    n_moves, t0, l0 = 0, time.perf_counter(), ctx.launch_count
    while not b.is_terminal():
        lm = b.get_legal_moves()
        assert b.make_move(lm[len(lm) // 2])
        b.get_winner(); b.get_stone_counts(); b.get_tensor_input()
        n_moves += 1
    dt = time.perf_counter() - t0
    assert ctx.launch_count - l0 == n_moves
    print(f"\\nsingle-board path: {1e6 * dt / n_moves:.1f} us per move incl. is_terminal/get_legal_moves/get_winner/get_stone_counts/"
          f"get_tensor_input ({n_moves} moves, 1 launch each)")


def test_device_sample_gather_into_the_replay_ring_single_rank():
    """The N > 1 data path of bench.py / INTEGRATION section 4 on one GPU (NCCL group of one rank): trajectories go from the
    engine's device buffer through DeviceSampleGather straight into the device replay ring -- no host hop -- and come out
    of the ring exactly as the host copy of the same campaign says."""
    import socket
    import torch
    import torch.distributed as dist
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200 import dist as odist
    from othello_reinforcement_learning_test_b200.self_play import samples_to_arrays
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", num_simulations=12, temperature_threshold=15,
                                       num_parallel_games=8, concurrent_games=48, seed=8, verbose=False)
        host = w.execute_episodes_packed(48)
        dptr, cnt = w._engine.samples_device()
        assert cnt == host.size
        gat = odist.DeviceSampleGather(torch.device("cuda", 0))
        gat.reserve(48 * 70)
        segs = gat.gather(dptr, cnt, 48 * 128, episodes=48)
        assert len(segs) == 1 and segs[0][1] == cnt
        buf = pkg.ReplayBuffer(max_size=cnt + 10)
        for seg, c in segs:
            buf.add_device(seg, c)
        torch.cuda.synchronize()
        assert len(buf) == cnt
        st, po, va = buf.gather(np.arange(cnt))
        want_st, want_po, want_va = samples_to_arrays(host, sort=False)          # ring order == the engine's flush order
        assert np.array_equal(st, want_st) and np.array_equal(po, want_po) and np.array_equal(va[:, 0], want_va.astype(np.float32))
        # a too-short source buffer is padded, not over-read
        segs2 = gat.gather(dptr, cnt, cnt, episodes=48)
        assert segs2[0][1] == cnt and bytes(segs2[0][0].cpu().numpy().tobytes()) == host.tobytes()
    finally:
        dist.destroy_process_group()
