"""GPU: the whole path with the real network -- search and self-play through the public classes."""
import numpy as np
import pytest
import torch

from oracle import cref

pytestmark = pytest.mark.gpu


def _positions(golden_games, n, seed):
    live = np.flatnonzero(golden_games["terminal"] == 0)
    idx = np.random.default_rng(seed).choice(live, n, replace=False)
    return golden_games["self_b"][idx], golden_games["opp_b"][idx]


@pytest.mark.parametrize("nb,nf,engine,n_pos", [(2, 32, "simt", 48), (5, 64, "tcgen05", 48), (10, 128, "tcgen05", 32)])
def test_native_search_equals_reference_search_under_identical_network_outputs(ctx, golden_games, nb, nf, engine, n_pos):
    """Visit counts of the all-on-device search == visit counts of the reference algorithm (oracle)
    when the oracle is given exactly the network outputs the device network produces."""
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.net import OthelloResNet
    torch.manual_seed(7)
    model = OthelloResNet(nb, nf).eval()
    m = pkg.MCTS(model, "cuda", c_puct=1.0)
    assert m.evaluator == "native"
    S, O = _positions(golden_games, n_pos, 3)
    vis, q, nev = m.search_arrays(S, O, 50)
    net = m._native_net()
    assert net.engine == engine          # 32 filters: the CUDA-core validation engine; 64 / 128: the tcgen05 product engine

    def ev(a, b):
        p, v = net.forward(np.array([a], np.uint64), np.array([b], np.uint64), out="probs")
        return p[0], float(v[0])
    for i in range(n_pos):
        res = cref.mcts_search(int(S[i]), int(O[i]), 50, 1.0, evaluator=ev)
        assert np.array_equal(vis[i], res["visits"]), i
        assert nev[i] == res["n_evals"]


def test_selfplay_with_the_network_through_the_reference_api(ctx):
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.net import OthelloResNet
    torch.manual_seed(42)
    model = OthelloResNet(5, 64).eval()
    cfg = {"mcts": {"num_simulations": 50, "c_puct": 1.0, "dirichlet_alpha": 0.3, "dirichlet_epsilon": 0.25},
           "self_play": {"temperature_threshold": 15, "num_parallel_games": 16}}       # "debug" config of BASELINE.json
    w = pkg.create_parallel_self_play_worker(cfg, model, torch.device("cuda"), seed=11, verbose=False)
    data = w.execute_episodes(num_episodes=32, add_dirichlet_noise=True)
    assert 32 * 50 < len(data) < 32 * 125
    for st, pol, val in data[:200]:
        assert st.shape == (3, 8, 8) and pol.shape == (65,) and abs(pol.sum() - 1) < 1e-5 and val in (-1.0, 0.0, 1.0)
        assert (pol[:64][st[2].reshape(-1) == 0] == 0).all()            # visits only on legal squares
    evals_per_game = w.last_stats["nn_evals"] / 32
    assert 2000 < evals_per_game < 51 * 125                              # reference: ~2,882 at 50 sims
    # weights change in place between iterations (trainer.py:258): the engine must pick that up
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(0.5)
    w2 = w.execute_episodes(num_episodes=2, add_dirichlet_noise=True)
    assert len(w2) > 50
    # the same model through MCTSPlayer-style calls
    m = pkg.MCTS(model, torch.device("cuda"))
    b = pkg.OthelloBitboard()
    assert m.get_best_action(b, 50) in b.get_legal_moves()
    ev = m.get_action_evaluations(b, 20)
    assert ev.dtype == np.int32 and ev.shape == (65,) and ((ev >= 0) & (ev <= 100)).all()


def test_eval_cache_and_dedup_are_result_transparent(ctx, golden_games):
    """Position-keyed cache + same-step dedup must not change a single visit count or sample."""
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.net import OthelloResNet
    torch.manual_seed(42)
    model = OthelloResNet(5, 64).eval()
    kw = dict(num_simulations=30, temperature_threshold=15, num_parallel_games=16, seed=21, concurrent_games=96, verbose=False)
    a = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", eval_cache=False, **kw)
    b = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", eval_cache=True, **kw)
    sa = a.execute_episodes_packed(160); sb = b.execute_episodes_packed(160)          # 96 slots, 160 episodes: refill
    sa = sa[np.lexsort((sa["ply"], sa["game"]))]; sb = sb[np.lexsort((sb["ply"], sb["game"]))]
    assert sa.size == sb.size and sa.tobytes() == sb.tobytes()
    st_a, st_b = a.last_stats, b.last_stats
    assert st_a["nn_evals"] == st_b["nn_evals"]                                       # expansions: same searches
    # a = no cache but searches of identical roots are shared; c = nothing shared at all: one evaluation per expansion
    c = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", eval_cache=False, share_searches=False, **kw)
    sc = c.execute_episodes_packed(160)
    sc = sc[np.lexsort((sc["ply"], sc["game"]))]
    assert sc.tobytes() == sa.tobytes()
    st_c = c.last_stats
    assert st_c["nn_positions"] == st_c["nn_evals"] == st_a["nn_evals"] and st_c["cache_hits"] == 0
    assert st_c["searches_run"] == sc.size and st_a["searches_run"] < sc.size          # one search per recorded ply vs shared
    assert st_a["nn_positions"] < st_a["nn_evals"] and st_a["cache_hits"] == 0
    assert st_b["nn_positions"] < st_a["nn_positions"]
    assert st_b["cache_hits"] > 0
    # second campaign on the same worker starts from an emptied cache and still agrees with the uncached engine
    sa2 = a.execute_episodes_packed(40); sb2 = b.execute_episodes_packed(40)
    sa2 = sa2[np.lexsort((sa2["ply"], sa2["game"]))]; sb2 = sb2[np.lexsort((sb2["ply"], sb2["game"]))]
    assert sa2.tobytes() == sb2.tobytes()
    # plain searches: transpositions inside one search hit the table too
    S, O = _positions(golden_games, 64, 17)
    m0 = pkg.MCTS(model, "cuda", c_puct=1.5); m1 = pkg.MCTS(model, "cuda", c_puct=1.5, eval_cache=True)
    v0, q0, e0 = m0.search_arrays(S, O, 100); v1, q1, e1 = m1.search_arrays(S, O, 100)
    assert np.array_equal(v0, v1) and np.array_equal(q0, q1) and np.array_equal(e0, e1)
    st = m1._tree.stats()
    assert st["nn_positions"] + st["cache_hits"] + st["same_step_duplicates"] == int(e1.sum())
    assert st["cache_hits"] + st["same_step_duplicates"] > 0


@pytest.mark.parametrize("nb,nf,games,slots", [(5, 64, 150, 96), (10, 128, 100, 100)])
def test_async_schedule_equals_lockstep_with_the_network(ctx, nb, nf, games, slots):
    """Run-until-miss (cache hits answered inside the advance kernel, one-tile network launches for small batches) must
    reproduce the lock-step campaign byte for byte -- with the cache on, and against the engine with everything off."""
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.net import OthelloResNet
    torch.manual_seed(42)
    model = OthelloResNet(nb, nf).eval()
    kw = dict(num_simulations=50, temperature_threshold=15, num_parallel_games=16, seed=33, concurrent_games=slots, verbose=False)
    order = lambda a: a[np.lexsort((a["ply"], a["game"]))]
    plain = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", eval_cache=False, share_searches=False, schedule="lockstep", **kw)
    lock = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", eval_cache=True, schedule="lockstep", **kw)
    asyn = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", eval_cache=True, schedule="async", **kw)
    asyn_nc = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", eval_cache=False, schedule="async", **kw)
    want = order(plain.execute_episodes_packed(games))
    for w in (lock, asyn, asyn_nc):
        got = order(w.execute_episodes_packed(games))
        assert got.size == want.size and got.tobytes() == want.tobytes()
        assert w.last_stats["nn_evals"] == plain.last_stats["nn_evals"]
    assert asyn.last_stats["schedule"] == "async" and lock.last_stats["schedule"] == "lockstep"
    # fewer network launches than 1 + sims per ply, and no more network positions than expansions
    assert asyn.last_stats["network_launches"] < lock.last_stats["network_launches"]
    assert asyn.last_stats["nn_positions"] + asyn.last_stats["cache_hits"] + asyn.last_stats["same_step_duplicates"] == plain.last_stats["nn_evals"]
    assert asyn_nc.last_stats["nn_positions"] == plain.last_stats["nn_evals"]
    auto = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", **kw)       # default: cache on, auto schedule
    got = order(auto.execute_episodes_packed(games))
    assert got.tobytes() == want.tobytes() and auto.last_stats["schedule"] == "async"


def test_searches_of_every_batch_size_agree_whatever_kernel_shape_evaluates_the_leaves(ctx, golden_games):
    """oth_search_run hands the network a batch whose size is only known on the device.  With 300 boards the bound is above
    the latency shape's 256 positions, so both network kernels are launched per step and the device count picks one; with
    40 boards only the latency shape runs.  The visit counts of a board must not depend on which batch it was searched in."""
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.net import OthelloResNet
    torch.manual_seed(5)
    model = OthelloResNet(5, 64).eval()
    S, O = _positions(golden_games, 300, 23)
    m = pkg.MCTS(model, "cuda", c_puct=1.0, eval_cache=True)
    v300, q300, e300 = m.search_arrays(S, O, 40)
    v40, q40, e40 = pkg.MCTS(model, "cuda", c_puct=1.0).search_arrays(S[:40], O[:40], 40)
    assert np.array_equal(v300[:40], v40) and np.array_equal(q300[:40], q40) and np.array_equal(e300[:40], e40)
    v1, q1, e1 = pkg.MCTS(model, "cuda", c_puct=1.0).search_arrays(S[7:8], O[7:8], 40)
    assert np.array_equal(v300[7:8], v1) and np.array_equal(q300[7:8], q1)


def test_packed_campaign_into_the_reusable_pinned_buffer(ctx):
    """execute_episodes_packed(reuse_buffer=True) is the same records, delivered as a view into one page-locked buffer."""
    import othello_reinforcement_learning_test_b200 as pkg
    kw = dict(num_simulations=8, temperature_threshold=15, num_parallel_games=16, concurrent_games=96, seed=11, verbose=False)
    fresh = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", **kw).execute_episodes_packed(96)
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", **kw)
    view = w.execute_episodes_packed(96, reuse_buffer=True)
    order = lambda a: a[np.lexsort((a["ply"], a["game"]))]                  # records are appended in completion order
    assert view.dtype == fresh.dtype and order(view).tobytes() == order(fresh).tobytes()
    addr = view.__array_interface__["data"][0]
    again = w.execute_episodes_packed(96, reuse_buffer=True)
    assert again.__array_interface__["data"][0] == addr and again.size > 0          # same buffer, overwritten in place
