"""GPU parity: bitboard kernels (through the C ABI) vs the oracle and the reference's golden vectors.
Bar: bit-exact."""
import numpy as np
import pytest

from oracle import cref
from oth_testutil import random_boards

pytestmark = pytest.mark.gpu


def test_golden_facts_through_the_board_class(golden_bitboard):
    import othello_reinforcement_learning_test_b200 as pkg
    b = pkg.OthelloBitboard()
    assert b.get_legal_moves() == [19, 26, 37, 44]                       # tests/test_bitboard.py:36
    assert b.get_legal_moves_bits() == golden_bitboard["initial_legal_bits"]
    assert b.get_stone_counts() == (2, 2) and not b.is_terminal()
    assert b.make_move(19) is True and b.move_count == 1
    assert list(b.get_stone_counts()) == golden_bitboard["after_move_19_counts"]   # tests/test_bitboard.py:69-71
    before = (b.self_board, b.opp_board, b.move_count)
    assert b.make_move(19) is False and b.make_move(0) is False and b.make_move(64) is False   # occupied / no flip / pass with moves
    assert b.make_move(-1) is False and b.make_move(65) is False
    assert (b.self_board, b.opp_board, b.move_count) == before
    for ev in golden_bitboard["edge_vectors"]:                            # REF rules on the A/H files
        e = pkg.OthelloBitboard(); e.self_board = ev["self"]; e.opp_board = ev["opp"]
        assert e.get_legal_moves() == ev["legal"]
    x = pkg.OthelloBitboard().get_tensor_input()
    assert x.shape == (3, 8, 8) and x.dtype == np.float32
    assert np.array_equal(x, cref.tensor_input_batch(np.array([cref.START_SELF], np.uint64),
                                                     np.array([cref.START_OPP], np.uint64))[0])
    sym = pkg.OthelloBitboard().get_symmetries(np.arange(65, dtype=np.float32))
    assert len(sym) == 8 and all(s.shape == (3, 8, 8) and p.shape == (65,) and p[64] == 64 for s, p in sym)


def test_random_game_through_the_board_class_reaches_terminal():
    import othello_reinforcement_learning_test_b200 as pkg
    rng = np.random.default_rng(0)
    b = pkg.OthelloBitboard()
    s, o, mc = cref.START_SELF, cref.START_OPP, 0
    for _ in range(130):
        if b.is_terminal():
            break
        moves = b.get_legal_moves()
        assert moves == cref.legal_list(s, o)
        a = int(rng.choice(moves))
        assert b.make_move(a)
        ok, s, o, mc = cref.make_move(s, o, mc, a)
        assert (b.self_board, b.opp_board, b.move_count, b.passed) == (s, o, mc, a == 64)
    assert b.is_terminal() and cref.is_terminal(s, o) and b.get_winner() == cref.winner(s, o)


def test_kernels_on_every_ply_of_the_reference_games(ctx, golden_games, golden_moves65):
    from othello_reinforcement_learning_test_b200 import bitboard as bb
    g = golden_games
    S, O = g["self_b"], g["opp_b"]
    assert np.array_equal(bb.legal_moves(S, O, ctx), g["legal"])
    t, w, c = bb.terminal_winner(S, O, ctx)
    assert np.array_equal(t, g["terminal"]) and np.array_equal(w, g["winner"])
    assert np.array_equal(c[:, 0], [bin(int(x)).count("1") for x in S])
    assert np.array_equal(bb.tensor_input(S[:2000], O[:2000], ctx), cref.tensor_input_batch(S[:2000], O[:2000]))
    m = golden_moves65
    a = m["self_b"].copy(); b = m["opp_b"].copy(); mc = m["move_count"].copy()
    ok = bb.make_move(a, b, mc, m["action"], ctx)
    assert np.array_equal(ok, m["ok"]) and np.array_equal(a, m["self_after"]) and np.array_equal(b, m["opp_after"])
    assert np.array_equal(mc, m["move_count_after"])


@pytest.mark.parametrize("n", [0, 1, 31, 33, 1_000_000])
def test_kernels_on_random_boards_and_ragged_sizes(ctx, n):
    from othello_reinforcement_learning_test_b200 import bitboard as bb
    s, o = random_boards(n, 11 + n)
    assert np.array_equal(bb.legal_moves(s, o, ctx), cref.legal_batch(s, o))
    pos = np.random.default_rng(n).integers(-2, 67, n).astype(np.int32)
    want = cref.flips_batch(s, o, np.clip(pos, 0, 63))
    want[(pos < 0) | (pos > 63)] = 0
    assert np.array_equal(bb.flips(s, o, pos, ctx), want)
    if n:
        t, w, _ = bb.terminal_winner(s, o, ctx)
        t2, w2 = cref.terminal_winner_batch(s, o)
        assert np.array_equal(t, t2) and np.array_equal(w, w2)
        act = np.random.default_rng(n + 1).integers(-1, 66, n).astype(np.int32)
        a, b, mc = s.copy(), o.copy(), np.zeros(n, np.int32)
        ok = bb.make_move(a, b, mc, act, ctx)
        a2, b2, mc2, ok2 = cref.make_move_batch(s, o, np.zeros(n, np.int32), act)
        assert np.array_equal(ok, ok2) and np.array_equal(a, a2) and np.array_equal(b, b2) and np.array_equal(mc, mc2)


def test_device_pointer_path(ctx):
    import torch
    from othello_reinforcement_learning_test_b200 import bitboard as bb
    s, o = random_boards(4096, 3)
    ds = torch.from_numpy(s.view(np.int64)).cuda(); do = torch.from_numpy(o.view(np.int64)).cuda()
    out = bb.legal_moves(ds, do, ctx)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), cref.legal_batch(s, o))


def test_perft_bit_exact(ctx, golden_bitboard):
    from othello_reinforcement_learning_test_b200 import bitboard as bb
    want = golden_bitboard["perft_ref_rules_oracle_0_11"]      # 0..9 from the reference, 10..11 oracle == SURVEY.md 8(c)
    for d in range(12):
        assert bb.perft(d, ctx=ctx) == want[d], f"perft({d})"
    assert bb.perft(7, ctx=ctx) == 55130 != golden_bitboard["perft_standard_for_contrast"][7]
    # from a late position with passes and terminal leaves inside the horizon
    s, o = 0x00003C3C3C3C0000 & ~0x0000001818000000, 0x0000001818000000
    for d in range(1, 7):
        assert bb.perft(d, s, o, ctx) == cref.perft(s, o, d)


def test_random_playouts_bit_exact_with_the_oracle(ctx):
    from othello_reinforcement_learning_test_b200 import bitboard as bb
    n = 20000
    got = bb.random_playouts(n, seed=1234, want_finals=True, ctx=ctx)
    want = cref.random_playouts(n, 1234, threads=0, want_finals=True)
    assert got["total_plies"] == want["total_plies"]
    assert np.array_equal(got["winner_hist"], want["winner_hist"])
    assert np.array_equal(got["plies"], want["plies"])
    assert np.array_equal(got["final_self"], want["final_self"]) and np.array_equal(got["final_opp"], want["final_opp"])
    # size-independent properties at scale: every game ends, mean length ~60.2 plies, winner split ~31/5/64 %
    big = bb.random_playouts(2_000_000, seed=7, ctx=ctx)
    assert big["winner_hist"].sum() == 2_000_000
    assert 59.5 < big["total_plies"] / 2_000_000 < 61.0
    frac = big["winner_hist"] / 2_000_000
    assert 0.60 < frac[0] < 0.68 and 0.28 < frac[2] < 0.35
