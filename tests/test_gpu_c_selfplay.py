"""GPU parity: the device-resident self-play campaign vs the reference's ParallelSelfPlayWorker."""
import numpy as np
import pytest

from oracle import cref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("schedule", ["lockstep", "async"])
@pytest.mark.parametrize("tag", ["a", "b"])
def test_deterministic_trace_equals_the_reference(golden_selfplay, tag, schedule):
    """temperature_threshold = 0 makes every move an arg-max: the whole game, the stored visit
    distributions and the value labels must equal what the reference's worker produced."""
    import othello_reinforcement_learning_test_b200 as pkg
    g = golden_selfplay
    cp, sims = float(g[f"{tag}_cfg"][0]), int(g[f"{tag}_cfg"][1])
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", num_simulations=sims, temperature_threshold=0,
                                   num_parallel_games=1, c_puct=cp, seed=1, verbose=False, schedule=schedule)
    data = w.execute_episodes(3, add_dirichlet_noise=True)        # three identical deterministic games
    assert w.last_stats["schedule"] == schedule
    n = len(g[f"{tag}_batched_value"])
    assert len(data) == 3 * n
    want_states = cref.tensor_input_batch(g[f"{tag}_batched_self"], g[f"{tag}_batched_opp"])
    for rep in range(3):
        for i in range(n):
            st, pol, val = data[rep * n + i]
            assert np.array_equal(st, want_states[i]), (rep, i)
            assert np.array_equal(pol, g[f"{tag}_batched_policy"][i]), (rep, i)
            assert val == float(g[f"{tag}_batched_value"][i])
    assert w.last_stats["nn_evals"] > 3 * n * sims * 0.8


def test_serial_worker_equals_the_reference(golden_selfplay):
    import othello_reinforcement_learning_test_b200 as pkg
    g = golden_selfplay
    cp, sims = float(g["a_cfg"][0]), int(g["a_cfg"][1])
    m = pkg.MCTS(None, "cuda", c_puct=cp)
    data = pkg.SelfPlayWorker(pkg.OthelloBitboard, m, num_simulations=sims, temperature_threshold=0).execute_episode(False)
    assert len(data) == len(g["a_serial_value"])
    for i, (st, pol, val) in enumerate(data):
        assert st.shape == (3, 8, 8) and pol.shape == (65,)
        assert np.array_equal(pol, g["a_serial_policy"][i]) and val == float(g["a_serial_value"][i])


@pytest.mark.parametrize("schedule", ["lockstep", "async"])
def test_sampled_campaign_is_consistent_under_replay(schedule):
    """With sampling (own RNG, not comparable to numpy's stream) every recorded game must still be a
    legal REF-rules game whose records agree with the oracle ply by ply."""
    import othello_reinforcement_learning_test_b200 as pkg
    sims = 20
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", num_simulations=sims, temperature_threshold=15,
                                   num_parallel_games=16, c_puct=1.0, seed=5, concurrent_games=64, verbose=False, schedule=schedule)
    smp = w.execute_episodes_packed(200, add_dirichlet_noise=True)     # 64 slots, 200 episodes: exercises refill
    assert w.last_stats["schedule"] == schedule
    assert sorted(set(smp["game"].tolist())) == list(range(200))
    order = np.lexsort((smp["ply"], smp["game"]))
    smp = smp[order]
    n_distinct_openings = set()
    for gid in range(200):
        rec = smp[smp["game"] == gid]
        assert (rec["ply"] == np.arange(rec.size)).all()
        s, o, mc = cref.START_SELF, cref.START_OPP, 0
        for i, r in enumerate(rec):
            assert (int(r["self_b"]), int(r["opp_b"])) == (s, o), (gid, i)
            assert int(r["legal"]) == cref.legal(s, o)
            vis = r["visits"].astype(np.int64)
            legal = cref.legal_list(s, o)
            assert vis.sum() == sims and all(vis[a] == 0 for a in range(65) if a not in legal)
            assert r["n_children"] == len(legal)
            want = cref.mcts_search(s, o, sims, 1.0)["visits"]
            assert np.array_equal(vis, want)
            if i + 1 < rec.size:
                nxt = (int(rec[i + 1]["self_b"]), int(rec[i + 1]["opp_b"]))
                cands = []
                for a in legal:
                    ok, s2, o2, _ = cref.make_move(s, o, mc, a)
                    if ok and (s2, o2) == nxt and vis[a] > 0:
                        cands.append(a)
                assert cands, (gid, i)
                if i >= 15:
                    assert cands[0] == int(np.argmax(vis))             # arg-max after the threshold
                ok, s, o, mc = cref.make_move(s, o, mc, cands[0])
        # the last record's move ends the game; label = winner at the terminal position x player
        last = rec[-1]
        vis = last["visits"].astype(np.int64)
        ends = []
        for a in cref.legal_list(s, o):
            ok, s2, o2, _ = cref.make_move(s, o, mc, a)
            if ok and vis[a] > 0 and cref.is_terminal(s2, o2):
                ends.append(cref.winner(s2, o2))
        assert ends
        labels = rec["value"].astype(np.int64)
        assert any((labels == np.array([wv * (1 if i % 2 == 0 else -1) for i in range(rec.size)])).all() for wv in ends)
        n_distinct_openings.add(tuple(rec["self_b"][:6].tolist()))
    assert len(n_distinct_openings) > 20          # sampling really happens before the threshold


def test_training_data_format_like_the_reference_tests():
    import othello_reinforcement_learning_test_b200 as pkg
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", num_simulations=5, temperature_threshold=15,
                                   num_parallel_games=4, seed=3, verbose=False)
    data = w.execute_episodes(4, add_dirichlet_noise=True)
    assert len(data) >= 4 * 8
    for st, pol, val in data:                                          # tests/test_train.py:116-130
        assert st.shape == (3, 8, 8) and st.dtype == np.float32
        assert pol.shape == (65,) and abs(pol.sum() - 1.0) < 1e-5
        assert val in (-1.0, 0.0, 1.0)
    assert w.execute_episodes(0) == []


def test_winner_black_flag_only_changes_labels():
    import othello_reinforcement_learning_test_b200 as pkg
    kw = dict(num_simulations=10, temperature_threshold=0, num_parallel_games=1, seed=1, verbose=False)
    a = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", **kw).execute_episodes_packed(1)
    b = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", winner_black=True, **kw).execute_episodes_packed(1)
    assert np.array_equal(a["self_b"], b["self_b"]) and np.array_equal(a["visits"], b["visits"])
    plies = a.size
    # REF label: winner is seen from the side to move at the end; black-relative flips it on odd game lengths
    assert np.array_equal(b["value"], a["value"] * (-1 if plies % 2 else 1))


def test_async_schedule_produces_the_lockstep_records_hashnet():
    """Run-until-miss vs lock-step with the integer evaluator: same episodes, same records, byte for byte
    (refills, sampling before the threshold, c_puct 1.5, 100 simulations: BASELINE config 4's search settings)."""
    import othello_reinforcement_learning_test_b200 as pkg
    kw = dict(num_simulations=100, temperature_threshold=20, num_parallel_games=16, c_puct=1.5, seed=9, concurrent_games=40, verbose=False)
    order = lambda a: a[np.lexsort((a["ply"], a["game"]))]
    a = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", schedule="lockstep", **kw)
    b = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", schedule="async", **kw)
    for n in (100, 7):                                                  # second campaign: derived seed, partly filled slots
        sa, sb = order(a.execute_episodes_packed(n)), order(b.execute_episodes_packed(n))
        assert sa.size == sb.size and sa.tobytes() == sb.tobytes()
        assert a.last_stats["nn_evals"] == b.last_stats["nn_evals"]
        assert (a.last_stats["schedule"], b.last_stats["schedule"]) == ("lockstep", "async")


def test_campaigns_on_one_worker_draw_from_different_streams():
    """game ids restart at 0 every run, so the run index must enter the sampling hash."""
    import othello_reinforcement_learning_test_b200 as pkg
    kw = dict(num_simulations=10, temperature_threshold=15, num_parallel_games=16, seed=5, concurrent_games=32, verbose=False)
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", **kw)
    order = lambda a: a[np.lexsort((a["ply"], a["game"]))]
    first, second = order(w.execute_episodes_packed(32)).copy(), order(w.execute_episodes_packed(32)).copy()
    assert first.tobytes() != second.tobytes()
    again = order(pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", **kw).execute_episodes_packed(32))
    assert again.tobytes() == first.tobytes()                          # same seed, same first campaign


def test_async_schedule_honours_the_variant_flags():
    """q_canonical / winner_black are read inside the run-until-miss kernel too: same records as lock-step."""
    import othello_reinforcement_learning_test_b200 as pkg
    kw = dict(num_simulations=30, temperature_threshold=10, num_parallel_games=16, c_puct=1.25, seed=3, concurrent_games=24, verbose=False,
              q_canonical=True, winner_black=True)
    order = lambda a: a[np.lexsort((a["ply"], a["game"]))]
    a = order(pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", schedule="lockstep", **kw).execute_episodes_packed(60))
    b = order(pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", schedule="async", **kw).execute_episodes_packed(60))
    assert a.size == b.size and a.tobytes() == b.tobytes()
    # root_n_sum with noise: per-game noise enters the search, the engine must fall back to lock-step by itself
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", schedule="async", root_n_sum=True, num_simulations=10,
                                   num_parallel_games=4, concurrent_games=8, seed=1, verbose=False)
    w.execute_episodes_packed(8, add_dirichlet_noise=True)
    assert w.last_stats["schedule"] == "lockstep"
