"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY (build container only).

Runs the REFERENCE ITSELF (compiled Cython bitboard from oracle/_ref + the
reference's unmodified Python MCTS / network imported from /root/reference),
checks our C restatement (oracle/ref_rules.c) and fp32 network restatement
(oracle/net_oracle.py) against it, and writes the committed fixtures under
tests/golden/:

  bitboard.json      perft 0..9 from the reference, edge vectors, API facts
  ref_games.npz      300 random games played by the reference, every ply
  ref_moves65.npz    all 65 actions tried on sampled positions (reject paths)
  mcts_ref.npz       visit vectors of the unmodified reference MCTS (stub model)
                     + the exact network outputs it saw; hash-net visit vectors
  selfplay_ref.npz   deterministic (threshold 0) self-play traces, batched + serial
  net_ref.npz        reference OthelloResNet fp32 outputs for synthetic weights
                     + checksum of the torch.manual_seed(42) initialisation
  replay_ref.npz     reference ReplayBuffer: seeded minibatches and statistics
  arena_ref.npz      reference Arena results with deterministic players

Usage:  python -m oracle.gen_golden
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

from . import cref, net_oracle, refload

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _bits_from_tensor(t: torch.Tensor):
    a = t.detach().cpu().numpy().reshape(-1, 3, 64)
    w = (1 << np.arange(64, dtype=np.uint64))
    s = (a[:, 0].astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
    o = (a[:, 1].astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
    return s, o


class StubModel:
    """Drop-in for `model` in the reference MCTS (only .eval() and __call__ are used,
    mcts.py:186-188): log of the hash-net pseudo-probabilities and its value."""

    def __init__(self):
        self.seen = []          # (self, opp, probs-as-the-reference-sees-them, value)

    def eval(self):
        return self

    def __call__(self, x):
        s, o = _bits_from_tensor(x)
        P, V = cref.hashnet_batch(s, o)
        logp = torch.log(torch.from_numpy(P))
        seen_p = torch.exp(logp).numpy()        # what mcts.py:191 will compute
        for i in range(len(s)):
            self.seen.append((int(s[i]), int(o[i]), seen_p[i].copy(), float(V[i])))
        return logp, torch.from_numpy(V.reshape(-1, 1).copy())


def ref_perft(Board, board, depth):
    if depth == 0 or board.is_terminal():
        return 1
    n = 0
    for a in board.get_legal_moves():
        c = board.copy()
        c.make_move(a)
        n += ref_perft(Board, c, depth - 1)
    return n


def gen_bitboard(ref, out):
    Board = ref.OthelloBitboard
    t0 = time.time()
    perft = [ref_perft(Board, Board(), d) for d in range(0, 10)]
    print("reference perft 0..9:", perft, f"({time.time() - t0:.0f}s)")
    mine = [cref.perft(cref.START_SELF, cref.START_OPP, d) for d in range(0, 12)]
    assert mine[:10] == perft, (mine, perft)
    survey = [1, 4, 12, 56, 244, 1396, 8200, 55130, 391014, 3020878, 24794040, 215593800]
    assert mine == survey, mine
    edge = []
    for s, o in [(1 << 0, 1 << 1), (1 << 7, 1 << 6), (1 << 6, 1 << 7), (1 << 9, 1 << 8), (1 << 0, 1 << 9)]:
        b = Board(); b.self_board = s; b.opp_board = o
        lm = b.get_legal_moves()
        assert lm == cref.legal_list(s, o)
        edge.append({"self": s, "opp": o, "legal": lm})
    b = Board()
    init_legal = b.get_legal_moves(); init_bits = int(b.get_legal_moves_bits())
    b.make_move(19)
    after19 = list(b.get_stone_counts())
    doc = {
        "source": "compiled reference (oracle/_ref) driven by oracle/gen_golden.py",
        "perft_ref_rules": perft,
        "perft_ref_rules_oracle_0_11": mine,
        "perft_standard_for_contrast": [1, 4, 12, 56, 244, 1396, 8200, 55092, 390216, 3005288, 24571284],
        "edge_vectors": edge,
        "initial_legal": init_legal, "initial_legal_bits": init_bits,
        "after_move_19_counts": after19,
        "start_self": cref.START_SELF, "start_opp": cref.START_OPP,
    }
    with open(os.path.join(out, "bitboard.json"), "w") as f:
        json.dump(doc, f, indent=1)


def gen_games(ref, out, n_games=300, seed=20260101):
    Board = ref.OthelloBitboard
    rng = np.random.default_rng(seed)
    rows = []
    n_pass = 0
    for g in range(n_games):
        b = Board()
        while True:
            s, o, mc = int(b.self_board), int(b.opp_board), int(b.move_count)
            legal = int(b.get_legal_moves_bits())
            term = bool(b.is_terminal())
            win = int(b.get_winner())
            # the oracle must agree at every ply
            assert cref.legal(s, o) == legal and cref.is_terminal(s, o) == term and cref.winner(s, o) == win
            if term:
                rows.append((g, s, o, mc, legal, -1, 1, win))
                break
            moves = b.get_legal_moves()
            a = int(rng.choice(moves))
            n_pass += a == 64
            rows.append((g, s, o, mc, legal, a, 0, win))
            ok = b.make_move(a)
            assert ok
            ok2, s2, o2, mc2 = cref.make_move(s, o, mc, a)
            assert ok2 and s2 == int(b.self_board) and o2 == int(b.opp_board) and mc2 == int(b.move_count)
            if a != 64:
                fl = cref.flips(a, s, o)
                assert (s | (1 << a) | fl) == int(b.opp_board) and (o & ~fl) == int(b.self_board)
    arr = np.array(rows, dtype=object)
    np.savez_compressed(
        os.path.join(out, "ref_games.npz"),
        game=np.array([r[0] for r in rows], np.int32),
        self_b=np.array([r[1] for r in rows], np.uint64),
        opp_b=np.array([r[2] for r in rows], np.uint64),
        move_count=np.array([r[3] for r in rows], np.int32),
        legal=np.array([r[4] for r in rows], np.uint64),
        action=np.array([r[5] for r in rows], np.int32),
        terminal=np.array([r[6] for r in rows], np.uint8),
        winner=np.array([r[7] for r in rows], np.int8),
    )
    print(f"ref_games: {len(rows)} plies, {n_pass} passes")
    # reject paths: all 65 actions on a sample of positions
    idx = rng.choice(len(rows), 400, replace=False)
    S, O, MC, A, OK, S2, O2, MC2 = [], [], [], [], [], [], [], []
    for i in idx:
        _, s, o, mc = rows[i][:4]
        for a in list(range(65)) + [-1, 65, 100]:
            b = Board(); b.self_board = s; b.opp_board = o; b.move_count = mc
            ok = bool(b.make_move(a))
            S.append(s); O.append(o); MC.append(mc); A.append(a); OK.append(ok)
            S2.append(int(b.self_board)); O2.append(int(b.opp_board)); MC2.append(int(b.move_count))
            r = cref.make_move(s, o, mc, a)
            assert r == (ok, S2[-1], O2[-1], MC2[-1]), (r, ok, a)
    np.savez_compressed(os.path.join(out, "ref_moves65.npz"),
                        self_b=np.array(S, np.uint64), opp_b=np.array(O, np.uint64), move_count=np.array(MC, np.int32),
                        action=np.array(A, np.int32), ok=np.array(OK, np.uint8),
                        self_after=np.array(S2, np.uint64), opp_after=np.array(O2, np.uint64),
                        move_count_after=np.array(MC2, np.int32))
    return rows


def _board(ref, s, o, mc=0):
    b = ref.OthelloBitboard(); b.self_board = s; b.opp_board = o; b.move_count = mc
    return b


def gen_mcts(ref, out, rows):
    rng = np.random.default_rng(7)
    nonterm = [r for r in rows if not r[6]]
    # ---- tier 1: unmodified reference MCTS with a stub model --------------------
    pick = [nonterm[i] for i in rng.choice(len(nonterm), 11, replace=False)]
    pick.insert(0, (0, cref.START_SELF, cref.START_OPP, 0, 0, 0, 0, 0))
    t1 = {k: [] for k in ("root_self", "root_opp", "c_puct", "sims", "visits", "policy_t1", "policy_t0",
                          "tab_off", "tab_self", "tab_opp", "tab_probs", "tab_value")}
    off = 0
    for (cp, sims) in [(1.0, 50), (1.5, 100)]:
        for r in pick:
            s, o = r[1], r[2]
            stub = StubModel()
            m = ref.MCTS(stub, torch.device("cpu"), c_puct=cp)
            pol1, rv = m.search(_board(ref, s, o), num_simulations=sims, temperature=1.0, add_dirichlet_noise=False)
            assert rv == 0.0
            seen = list(stub.seen)
            table = {(a, b): (p, v) for a, b, p, v in seen}
            # determinism facts from the survey: noise is inert, T=0 is the argmax one-hot
            np.random.seed(1)
            pol1n, _ = ref.MCTS(StubModel(), torch.device("cpu"), c_puct=cp).search(
                _board(ref, s, o), num_simulations=sims, temperature=1.0, add_dirichlet_noise=True)
            assert np.array_equal(pol1, pol1n), "Dirichlet noise changed the result?!"
            pol0, _ = ref.MCTS(StubModel(), torch.device("cpu"), c_puct=cp).search(
                _board(ref, s, o), num_simulations=sims, temperature=0.0)
            # our oracle with the very same network outputs
            res = cref.mcts_search(s, o, sims, cp, evaluator=lambda a, b: table[(a, b)])
            pol1_o = cref.policy_from_visits(res["visits"], res["is_child"], 1.0)
            pol0_o = cref.policy_from_visits(res["visits"], res["is_child"], 0.0)
            assert np.array_equal(pol1_o, pol1), (pol1_o, pol1)
            assert np.array_equal(pol0_o, pol0)
            assert res["n_evals"] == len(seen)
            t1["root_self"].append(s); t1["root_opp"].append(o); t1["c_puct"].append(cp); t1["sims"].append(sims)
            t1["visits"].append(res["visits"]); t1["policy_t1"].append(pol1); t1["policy_t0"].append(pol0)
            t1["tab_off"].append(off)
            for a, b, p, v in seen:
                t1["tab_self"].append(a); t1["tab_opp"].append(b); t1["tab_probs"].append(p); t1["tab_value"].append(v)
            off += len(seen)
    t1["tab_off"].append(off)
    print(f"mcts tier1: {len(t1['root_self'])} searches, {off} recorded evaluations -- oracle == reference")

    # ---- tier 2: exact hash-net through a patched NN boundary (_predict) ---------
    pick2 = [nonterm[i] for i in rng.choice(len(nonterm), 300, replace=False)]
    t2 = {k: [] for k in ("root_self", "root_opp", "c_puct", "sims", "visits")}
    for (cp, sims) in [(1.0, 50), (1.5, 100), (1.1, 25)]:
        S = np.array([r[1] for r in pick2], np.uint64); O = np.array([r[2] for r in pick2], np.uint64)
        vis_o, _ = cref.mcts_search_hashnet_batch(S, O, sims, cp)
        for j, r in enumerate(pick2):
            m = ref.MCTS(None, torch.device("cpu"), c_puct=cp)

            def pred(board_tensor):
                a, b = _bits_from_tensor(board_tensor)
                p, v = cref.hashnet(int(a[0]), int(b[0]))
                return p, torch.tensor([[v]], dtype=torch.float32)
            m._predict = pred
            pol, _ = m.search(_board(ref, r[1], r[2]), num_simulations=sims, temperature=1.0)
            vis_ref = np.rint(pol * sims).astype(np.int32)
            assert np.array_equal(vis_ref, vis_o[j]), (j, vis_ref, vis_o[j])
            assert np.array_equal(pol, cref.policy_from_visits(vis_o[j], (vis_o[j] >= 0) & (pol >= 0) &
                                                                np.isin(np.arange(65), cref.legal_list(r[1], r[2])), 1.0))
            t2["root_self"].append(r[1]); t2["root_opp"].append(r[2]); t2["c_puct"].append(cp); t2["sims"].append(sims)
            t2["visits"].append(vis_o[j])
    print(f"mcts tier2: {len(t2['root_self'])} hash-net searches -- oracle == reference")

    # ---- BatchMCTS == serial (survey 0.4) ----------------------------------------
    bm = ref.BatchMCTS(None, torch.device("cpu"), c_puct=1.5)

    def bpred(boards):
        S = np.array([int(b.self_board) for b in boards], np.uint64); O = np.array([int(b.opp_board) for b in boards], np.uint64)
        P, V = cref.hashnet_batch(S, O)
        return P, V.reshape(-1, 1)
    bm.batch_predict = bpred
    sub = pick2[:16]
    resb = bm.search_batch([_board(ref, r[1], r[2]) for r in sub], num_simulations=50, temperature=1.0)
    vis_o, _ = cref.mcts_search_hashnet_batch(np.array([r[1] for r in sub], np.uint64),
                                              np.array([r[2] for r in sub], np.uint64), 50, 1.5)
    for j, (pol, _) in enumerate(resb):
        assert np.array_equal(np.rint(pol * 50).astype(np.int32), vis_o[j])
    print("BatchMCTS.search_batch == oracle on 16 positions")

    np.savez_compressed(
        os.path.join(out, "mcts_ref.npz"),
        t1_root_self=np.array(t1["root_self"], np.uint64), t1_root_opp=np.array(t1["root_opp"], np.uint64),
        t1_c_puct=np.array(t1["c_puct"], np.float64), t1_sims=np.array(t1["sims"], np.int32),
        t1_visits=np.array(t1["visits"], np.int32), t1_policy_t1=np.array(t1["policy_t1"], np.float32),
        t1_policy_t0=np.array(t1["policy_t0"], np.float32), t1_tab_off=np.array(t1["tab_off"], np.int64),
        t1_tab_self=np.array(t1["tab_self"], np.uint64), t1_tab_opp=np.array(t1["tab_opp"], np.uint64),
        t1_tab_probs=np.array(t1["tab_probs"], np.float32), t1_tab_value=np.array(t1["tab_value"], np.float32),
        t2_root_self=np.array(t2["root_self"], np.uint64), t2_root_opp=np.array(t2["root_opp"], np.uint64),
        t2_c_puct=np.array(t2["c_puct"], np.float64), t2_sims=np.array(t2["sims"], np.int32),
        t2_visits=np.array(t2["visits"], np.int8),
    )


def gen_selfplay(ref, out):
    """Deterministic traces: temperature_threshold=0 makes every move an argmax."""
    res = {}
    for tag, (cp, sims) in {"a": (1.0, 50), "b": (1.5, 100)}.items():
        w = ref.ParallelSelfPlayWorker(ref.OthelloBitboard, None, torch.device("cpu"), num_simulations=sims,
                                       temperature_threshold=0, num_parallel_games=1, c_puct=cp)

        def bpred(boards):
            S = np.array([int(b.self_board) for b in boards], np.uint64)
            O = np.array([int(b.opp_board) for b in boards], np.uint64)
            P, V = cref.hashnet_batch(S, O)
            return P, V.reshape(-1, 1)
        w.batch_mcts.batch_predict = bpred
        data = w._execute_batch(batch_size=1, add_dirichlet_noise=True)
        S, O = _bits_from_tensor(torch.from_numpy(np.stack([d[0] for d in data])))
        res[f"{tag}_batched_self"] = S; res[f"{tag}_batched_opp"] = O
        res[f"{tag}_batched_legal"] = np.array(
            [int((d[0][2].reshape(64).astype(np.uint64) * (1 << np.arange(64, dtype=np.uint64))).sum(dtype=np.uint64))
             for d in data], np.uint64)
        res[f"{tag}_batched_policy"] = np.stack([d[1] for d in data]).astype(np.float32)
        res[f"{tag}_batched_value"] = np.array([d[2] for d in data], np.float32)
        # serial worker (stores the one-hot policy after the threshold, self_play.py:87-98)
        m = ref.MCTS(None, torch.device("cpu"), c_puct=cp)

        def pred(board_tensor):
            a, b = _bits_from_tensor(board_tensor)
            p, v = cref.hashnet(int(a[0]), int(b[0]))
            return p, torch.tensor([[v]], dtype=torch.float32)
        m._predict = pred
        sw = ref.SelfPlayWorker(ref.OthelloBitboard, m, num_simulations=sims, temperature_threshold=0)
        data2 = sw.execute_episode(add_dirichlet_noise=False)
        S2, O2 = _bits_from_tensor(torch.from_numpy(np.stack([d[0] for d in data2])))
        assert np.array_equal(S, S2) and np.array_equal(O, O2)
        res[f"{tag}_serial_policy"] = np.stack([d[1] for d in data2]).astype(np.float32)
        res[f"{tag}_serial_value"] = np.array([d[2] for d in data2], np.float32)
        res[f"{tag}_cfg"] = np.array([cp, sims], np.float64)
        print(f"selfplay trace {tag}: {len(data)} plies, label of first sample {data[0][2]}")
    np.savez_compressed(os.path.join(out, "selfplay_ref.npz"), **res)


def gen_net(ref, out, rows):
    rng = np.random.default_rng(11)
    nonterm = [r for r in rows if not r[6]]
    pick = [nonterm[i] for i in rng.choice(len(nonterm), 96, replace=False)]
    S = np.array([cref.START_SELF] + [r[1] for r in pick], np.uint64)
    O = np.array([cref.START_OPP] + [r[2] for r in pick], np.uint64)
    x = net_oracle.boards_to_tensor(S, O)
    # the reference's own tensor builder agrees with ours
    for i in range(len(S)):
        assert np.array_equal(_board(ref, int(S[i]), int(O[i])).get_tensor_input(), x[i].numpy())
    res = {"self_b": S, "opp_b": O}
    for nb, nf, seed in [(2, 32, 5), (5, 64, 6), (10, 128, 7)]:
        sd = net_oracle.make_state_dict(nb, nf, seed, gain=net_oracle.SYNTH_GAIN[(nb, nf)])
        m = ref.OthelloResNet(nb, nf); m.load_state_dict(sd); m.eval()
        with torch.no_grad():
            lp, v = m(x)
        lp2, v2 = net_oracle.forward_fp32(sd, x)
        assert torch.equal(lp, lp2) and torch.equal(v, v2)
        res[f"logp_{nb}x{nf}_s{seed}"] = lp.numpy(); res[f"value_{nb}x{nf}_s{seed}"] = v.numpy().reshape(-1)
    # seed-42 initialisation (BASELINE configs): checksum + outputs
    for nb, nf in [(5, 64), (10, 128)]:
        torch.manual_seed(42)
        m = ref.OthelloResNet(nb, nf); m.eval()
        sd = m.state_dict()
        chk = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()], np.float64)
        with torch.no_grad():
            lp, v = m(x)
        res[f"init42_{nb}x{nf}_checksum"] = chk
        res[f"init42_{nb}x{nf}_logp"] = lp.numpy(); res[f"init42_{nb}x{nf}_value"] = v.numpy().reshape(-1)
    np.savez_compressed(os.path.join(out, "net_ref.npz"), **res)
    print("net_ref written")


def bulk_positions(n=10240, seed=123):
    """`n` distinct self-play positions: games played by the CPU port of the batched worker (C tree, integer test
    evaluator, 16 simulations, every move sampled), i.e. the kind of position the search hands to the network."""
    from . import selfplay_port
    sp = selfplay_port.CpuSelfPlay(None, num_simulations=16, temperature_threshold=200, num_parallel_games=16, seed=seed)
    pos = set()
    while len(pos) < n + n // 10:
        for h in sp.play(16)["samples"]:
            pos.update((s, o) for (s, o, _, _) in h)
    pos = sorted(pos)
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(len(pos), n, replace=False))
    S = np.array([pos[i][0] for i in pick], np.uint64); O = np.array([pos[i][1] for i in pick], np.uint64)
    return S, O


def gen_net_bulk(ref, out, n=10240, n_hot=2048):
    """fp32 outputs of the REFERENCE's own OthelloResNet (10x128) on 10,240 self-play positions: the seed-42
    initialisation every BASELINE config uses, and -- on the first 2,048 -- the synthetic weights at gain 1.0
    (saturated policies, p_max ~0.99: the hard case for bf16).  VERDICT r1 item 4(a,b)."""
    S, O = bulk_positions(n)
    x = net_oracle.boards_to_tensor(S, O)
    res = {"self_b": S, "opp_b": O}
    torch.manual_seed(42)
    m = ref.OthelloResNet(10, 128); m.eval()
    with torch.no_grad():
        lp, v = m(x)
    lp2, v2 = net_oracle.forward_fp32(m.state_dict(), x)
    assert torch.equal(lp, lp2) and torch.equal(v, v2)
    res["init42_10x128_logp"] = lp.numpy(); res["init42_10x128_value"] = v.numpy().reshape(-1)
    sd = net_oracle.make_state_dict(10, 128, 7, gain=1.0)
    m = ref.OthelloResNet(10, 128); m.load_state_dict(sd); m.eval()
    with torch.no_grad():
        lp, v = m(x[:n_hot])
    lp2, v2 = net_oracle.forward_fp32(sd, x[:n_hot])
    assert torch.equal(lp, lp2) and torch.equal(v, v2)
    res["gain1_10x128_s7_logp"] = lp.numpy(); res["gain1_10x128_s7_value"] = v.numpy().reshape(-1)
    np.savez_compressed(os.path.join(out, "net_bulk_ref.npz"), **res)
    print("net_bulk_ref written:", n, "positions; p_max init42", float(np.exp(res["init42_10x128_logp"]).max()),
          "gain1", float(np.exp(res["gain1_10x128_s7_logp"]).max()))


def gen_symmetries(ref, out):
    """get_symmetries of the compiled reference board (bitboard.pyx:338-370) on 64 positions with random policies:
    the 8 dihedral images of (planes, pi) in the reference's own order.  Stored as bit-planes + float32 policies."""
    g = np.load(os.path.join(out, "ref_games.npz"))
    live = np.flatnonzero(g["terminal"] == 0)
    idx = np.random.default_rng(8).choice(live, 64, replace=False)
    S, O = g["self_b"][idx], g["opp_b"][idx]
    rng = np.random.default_rng(9)
    pis = rng.random((64, 65)).astype(np.float32)
    pis /= pis.sum(axis=1, keepdims=True)
    w = (1 << np.arange(64, dtype=np.uint64))
    planes = np.zeros((64, 8, 3), np.uint64); pol = np.zeros((64, 8, 65), np.float32)
    for i in range(64):
        b = _board(ref, int(S[i]), int(O[i]))
        sym = b.get_symmetries(pis[i])
        assert len(sym) == 8
        for k, (pl, p) in enumerate(sym):
            assert pl.shape == (3, 8, 8) and p.shape == (65,)
            planes[i, k] = ((pl.reshape(3, 64) > 0.5).astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
            pol[i, k] = p
    np.savez_compressed(os.path.join(out, "symmetry_ref.npz"), self_b=S, opp_b=O, pi=pis, planes=planes, policy=pol)
    print("symmetry_ref written")


def replay_fixture_data(games):
    """Deterministic training tuples built from the committed reference games (shared with the tests)."""
    rng = np.random.default_rng(99)
    live = np.flatnonzero(games["terminal"] == 0)[:700]
    S, O = games["self_b"][live], games["opp_b"][live]
    states = cref.tensor_input_batch(S, O)
    data = []
    for i in range(len(live)):
        legal = cref.legal_list(int(S[i]), int(O[i]))
        sims = int(rng.choice([25, 50, 100]))
        counts = rng.multinomial(sims, np.ones(len(legal)) / len(legal))
        pol = np.zeros(65, np.float32)
        pol[legal] = counts.astype(np.float32) / np.float32(sims)
        data.append((states[i], pol, float(rng.integers(-1, 2))))
    return data


def gen_replay(ref, out):
    import random
    games = dict(np.load(os.path.join(out, "ref_games.npz")))
    data = replay_fixture_data(games)
    buf = ref.ReplayBuffer(max_size=500)                     # smaller than the data: the oldest 200 fall out
    buf.add(data[:300]); buf.add(data[300:])
    random.seed(5)
    st, po, va = buf.sample(48)
    random.seed(6)
    st2, po2, va2 = buf.sample(500)
    stats = buf.get_statistics()
    np.savez_compressed(os.path.join(out, "replay_ref.npz"), states=st, policies=po, values=va,
                        all_states_sum=st2.sum(axis=0), all_policies=po2, all_values=va2,
                        stats=np.array([stats["size"], stats["max_size"], stats["fill_rate"], stats["value_mean"], stats["value_std"]]))
    print("replay_ref written", st.shape, po.shape, va.shape, stats)


def gen_arena(ref, out):
    """Reference Arena with deterministic players: Greedy vs Greedy, hash-net MCTS vs Greedy."""
    res = {}
    arena = ref.Arena(verbose=False)

    def pack(results):
        return np.array([[r.winner, r.player1_score, r.player2_score, r.num_moves] for r in results], np.int32)
    res["greedy_vs_greedy"] = pack(arena.play_matches(ref.GreedyPlayer("A"), ref.GreedyPlayer("B"), num_games=4))
    for sims in (8, 25):
        mp = ref.MCTSPlayer(None, torch.device("cpu"), num_simulations=sims)

        def pred(board_tensor):
            a, b = _bits_from_tensor(board_tensor)
            p, v = cref.hashnet(int(a[0]), int(b[0]))
            return p, torch.tensor([[v]], dtype=torch.float32)
        mp.mcts._predict = pred
        res[f"mcts{sims}_vs_greedy"] = pack(arena.play_matches(mp, ref.GreedyPlayer("G"), num_games=4))
        res[f"greedy_vs_mcts{sims}"] = pack(arena.play_matches(ref.GreedyPlayer("G"), mp, num_games=2, alternate_colors=False))
    ev = ref.evaluate_player(ref.GreedyPlayer("A"), ref.GreedyPlayer("B"), num_games=6, verbose=False)
    res["evaluate_greedy"] = np.array([ev["win_rate"], ev["avg_score"], ev["avg_moves"]])
    np.savez_compressed(os.path.join(out, "arena_ref.npz"), **res)
    print("arena_ref written", {k: v.tolist() for k, v in res.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = refload.reference_python()
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    if only == "net":          # the network fixture alone (the positions come from the committed ref_games.npz)
        g = np.load(os.path.join(OUT, "ref_games.npz"))
        rows = [(int(a), int(b), int(c), int(d), int(e), int(f), int(t), int(w)) for a, b, c, d, e, f, t, w in
                zip(g["game"], g["self_b"], g["opp_b"], g["move_count"], g["legal"], g["action"], g["terminal"], g["winner"])]
        gen_net(ref, OUT, rows)
        return
    if only == "net_bulk":
        gen_net_bulk(ref, OUT)
        return
    if only == "sym":
        gen_symmetries(ref, OUT)
        return
    if only == "replay":
        gen_replay(ref, OUT)
        return
    if only == "arena":
        gen_arena(ref, OUT)
        return
    gen_bitboard(ref, OUT)
    rows = gen_games(ref, OUT)
    gen_mcts(ref, OUT, rows)
    gen_selfplay(ref, OUT)
    gen_net(ref, OUT, rows)
    gen_net_bulk(ref, OUT)
    gen_symmetries(ref, OUT)
    gen_replay(ref, OUT)
    gen_arena(ref, OUT)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    sys.exit(main())
