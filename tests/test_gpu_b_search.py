"""GPU parity: the SoA tree (select / expand / backup) vs the reference MCTS.
Bar: identical visit-count vectors under identical network outputs."""
import numpy as np
import pytest

from oracle import cref

pytestmark = pytest.mark.gpu


def _positions(golden_games, n, seed):
    live = np.flatnonzero(golden_games["terminal"] == 0)
    idx = np.random.default_rng(seed).choice(live, n, replace=False)
    return golden_games["self_b"][idx], golden_games["opp_b"][idx]


def test_tier1_unmodified_reference_visits_via_external_evaluator(golden_mcts):
    """Feed the GPU tree exactly the network outputs the unmodified reference MCTS saw."""
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.mcts import TreeSearch
    g = golden_mcts
    off = g["t1_tab_off"]
    ctx = pkg.Context.default(0)
    for i in range(len(g["t1_root_self"])):
        tab = {(int(a), int(b)): (p, v) for a, b, p, v in
               zip(g["t1_tab_self"][off[i]:off[i + 1]], g["t1_tab_opp"][off[i]:off[i + 1]],
                   g["t1_tab_probs"][off[i]:off[i + 1]], g["t1_tab_value"][off[i]:off[i + 1]])}
        sims = int(g["t1_sims"][i])
        t = TreeSearch(ctx, 1, sims)
        t.configure(float(g["t1_c_puct"][i]), 0.3, 0.25, 0)
        t.begin(g["t1_root_self"][i:i + 1].copy(), g["t1_root_opp"][i:i + 1].copy())
        n_eval = 0
        for _ in range(sims + 1):
            ls, lo, need = t.collect()
            probs = np.zeros((1, 65), np.float32); val = np.zeros(1, np.float32)
            if need[0]:
                probs[0], val[0] = tab[(int(ls[0]), int(lo[0]))]
                n_eval += 1
            t.apply(probs, val)
        vis, q, nev = t.results()
        assert np.array_equal(vis[0], g["t1_visits"][i]), f"search {i}"
        assert nev[0] == n_eval == off[i + 1] - off[i]
        assert np.array_equal(t.policy(1.0)[0], g["t1_policy_t1"][i])
        assert np.array_equal(t.policy(0.0)[0], g["t1_policy_t0"][i])
        t.close()


def test_tier2_hashnet_on_device_equals_reference(golden_mcts):
    import othello_reinforcement_learning_test_b200 as pkg
    g = golden_mcts
    for cp, sims in sorted({(float(c), int(s)) for c, s in zip(g["t2_c_puct"], g["t2_sims"])}):
        sel = (g["t2_c_puct"] == cp) & (g["t2_sims"] == sims)
        m = pkg.MCTS(None, "cuda", c_puct=cp)            # model None -> built-in integer evaluator
        vis, q, nev = m.search_arrays(g["t2_root_self"][sel], g["t2_root_opp"][sel], sims)
        assert np.array_equal(vis, g["t2_visits"][sel].astype(np.int32)), (cp, sims)


@pytest.mark.parametrize("cp,sims,n", [(1.0, 50, 10000), (1.5, 100, 3000), (2.7, 17, 3000)])
def test_bulk_visit_parity_with_the_oracle(golden_games, cp, sims, n):
    import othello_reinforcement_learning_test_b200 as pkg
    s, o = _positions(golden_games, n, 5)
    want, want_ev = cref.mcts_search_hashnet_batch(s, o, sims, cp)
    m = pkg.MCTS(None, "cuda", c_puct=cp)
    vis, q, nev = m.search_arrays(s, o, sims)
    bad = np.flatnonzero((vis != want).any(axis=1))
    assert bad.size == 0, f"{bad.size} of {n} searches differ, first: {bad[:5]}"
    assert np.array_equal(nev, want_ev)
    assert (vis.sum(axis=1) == sims).all()


def test_q_values_and_hint_evaluations(golden_games):
    import othello_reinforcement_learning_test_b200 as pkg
    s, o = _positions(golden_games, 64, 9)
    m = pkg.MCTS(None, "cuda", c_puct=1.0)
    vis, q, _ = m.search_arrays(s, o, 50)
    for i in range(64):
        res = cref.mcts_search(int(s[i]), int(o[i]), 50, 1.0)
        assert np.array_equal(q[i], res["q"])                       # float64, bit for bit
    b = pkg.OthelloBitboard(); b.self_board = int(s[0]); b.opp_board = int(o[0])
    ev = m.get_action_evaluations(b, 50)
    res = cref.mcts_search(int(s[0]), int(o[0]), 50, 1.0)
    want = np.zeros(65, np.int32)
    for a in cref.legal_list(int(s[0]), int(o[0])):
        want[a] = max(0, min(100, int((res["q"][a] + 1.0) * 50.0)))   # mcts.py:356-360
    assert ev.dtype == np.int32 and np.array_equal(ev, want)


def test_mcts_class_api_like_the_reference_tests(golden_games):
    """Shape/contract checks in the spirit of the reference's tests/test_mcts.py."""
    import othello_reinforcement_learning_test_b200 as pkg
    m = pkg.MCTS(None, "cuda", c_puct=1.0, dirichlet_alpha=0.3, dirichlet_epsilon=0.25)
    b = pkg.OthelloBitboard()
    pol, rv = m.search(b, num_simulations=25, temperature=1.0, add_dirichlet_noise=False)
    assert pol.shape == (65,) and pol.dtype == np.float32 and rv == 0.0
    assert abs(pol.sum() - 1.0) < 1e-6
    legal = b.get_legal_moves()
    assert all(pol[a] == 0 for a in range(65) if a not in legal)      # tests/test_mcts.py:158-176
    pol0, _ = m.search(b, num_simulations=25, temperature=0.0)
    assert sorted(pol0.tolist())[-1] == 1.0 and pol0.sum() == 1.0     # one-hot (tests/test_mcts.py:236-256)
    np.random.seed(1); p1 = m.get_action_probs(b, 25, add_dirichlet_noise=True)
    np.random.seed(2); p2 = m.get_action_probs(b, 25, add_dirichlet_noise=True)
    assert np.array_equal(p1, p2) and np.array_equal(p1, pol)          # noise is inert in REF mode
    a = m.get_best_action(b, 25)
    assert a in legal and a == int(np.argmax(pol0))
    assert m.get_best_action(b, 0) == legal[0]
    # a whole game is playable
    for _ in range(130):
        if b.is_terminal():
            break
        assert b.make_move(m.get_best_action(b, 5))
    assert b.is_terminal()
    # batched == one at a time (BatchMCTS.search_batch is result-identical to MCTS.search)
    s, o = _positions(golden_games, 8, 21)
    boards = []
    for i in range(8):
        x = pkg.OthelloBitboard(); x.self_board = int(s[i]); x.opp_board = int(o[i]); boards.append(x)
    bm = pkg.BatchMCTS(None, "cuda", c_puct=1.5)
    res = bm.search_batch(boards, 50, temperature=1.0)
    for x, (p, v) in zip(boards, res):
        assert np.array_equal(p, bm.search(x, 50)[0]) and v == 0.0
    assert bm.search_batch([], 10) == []


def test_external_evaluator_with_an_arbitrary_callable(golden_games):
    """`model` can be any callable returning (log_probs, value): the tree stays on the GPU."""
    import torch
    import othello_reinforcement_learning_test_b200 as pkg

    class Stub:
        calls = 0

        def eval(self):
            return self

        def __call__(self, x):
            Stub.calls += 1
            a = x.reshape(x.shape[0], 3, 64).cpu().numpy()
            w = (1 << np.arange(64, dtype=np.uint64))
            s = (a[:, 0].astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
            o = (a[:, 1].astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
            P, V = cref.hashnet_batch(s, o)
            return torch.log(torch.from_numpy(P)), torch.from_numpy(V.reshape(-1, 1).copy())

    s, o = _positions(golden_games, 4, 33)
    m = pkg.MCTS(Stub(), torch.device("cpu"), c_puct=1.0)
    assert m.evaluator == "external"
    vis, _, nev = m.search_arrays(s, o, 30)
    for i in range(4):
        def ev(a, b):
            p, v = cref.hashnet(a, b)
            return torch.exp(torch.log(torch.from_numpy(p))).numpy(), v
        assert np.array_equal(vis[i], cref.mcts_search(int(s[i]), int(o[i]), 30, 1.0, evaluator=ev)["visits"])
    assert Stub.calls >= 1


def test_canonical_flags_match_the_oracles_opt_in_mode(golden_games):
    import othello_reinforcement_learning_test_b200 as pkg
    s, o = _positions(golden_games, 200, 41)
    m = pkg.MCTS(None, "cuda", c_puct=1.25, root_n_sum=True, q_canonical=True)
    vis, _, _ = m.search_arrays(s, o, 40)
    for i in range(200):
        res = cref.mcts_search(int(s[i]), int(o[i]), 40, 1.25, root_n_sum=True, q_canonical=True)
        assert np.array_equal(vis[i], res["visits"]), i
    # with the root's N in play the search spreads its visits (unlike REF mode)
    ref_mode, _, _ = pkg.MCTS(None, "cuda", c_puct=1.25).search_arrays(s, o, 40)
    assert (vis > 0).sum() > (ref_mode > 0).sum()


def test_inflight_waves_k1_is_exact_and_larger_k_is_close(golden_games):
    """Opt-in throughput mode: K simulations per game per launch with virtual loss.
    K = 1 must reproduce the reference exactly; K > 1 deviates by construction -- report how much."""
    import othello_reinforcement_learning_test_b200 as pkg
    s, o = _positions(golden_games, 300, 77)
    want, want_ev = cref.mcts_search_hashnet_batch(s, o, 50, 1.0)
    v1, _, e1 = pkg.MCTS(None, "cuda", c_puct=1.0, inflight=1).search_arrays(s, o, 50)
    assert np.array_equal(v1, want)
    # the wave machinery itself with K = 1 (goes through oth_search_run_waves)
    from othello_reinforcement_learning_test_b200.mcts import TreeSearch
    ctx = pkg.Context.default(0)
    t = TreeSearch(ctx, 300, 50); t.configure(1.0, 0.3, 0.25, pkg._lib.FLAG_EVAL_HASHNET); t.begin(s.copy(), o.copy())
    pkg._lib.check(ctx.lib.oth_search_run_waves(t.handle, None, 50, 1, 0, 0))
    vw, _, ew = t.results()
    assert np.array_equal(vw, want) and np.array_equal(ew, want_ev)
    tvs = {}
    for K in (4, 8, 16):
        vk, _, _ = pkg.MCTS(None, "cuda", c_puct=1.0, inflight=K).search_arrays(s, o, 50)
        assert (vk.sum(axis=1) == 50).all() and ((vk > 0) <= (want >= 0)).all()
        legal = np.array([[a in cref.legal_list(int(x), int(y)) for a in range(65)] for x, y in zip(s, o)])
        assert not (vk[~legal] > 0).any()
        tvs[K] = float(0.5 * np.abs(vk / 50.0 - want / 50.0).sum(axis=1).mean())
    print("mean total-variation distance to the K=1 visit distribution:", tvs)
    assert tvs[4] <= tvs[16] + 0.05 and tvs[16] < 0.6
    # deterministic
    a, _, _ = pkg.MCTS(None, "cuda", inflight=8).search_arrays(s, o, 50)
    b, _, _ = pkg.MCTS(None, "cuda", inflight=8).search_arrays(s, o, 50)
    assert np.array_equal(a, b)
