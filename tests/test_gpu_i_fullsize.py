"""GPU: size-independent properties of a BASELINE-sized campaign (10x128 network, 50 simulations per move,
thousands of concurrent games, cache + shared searches on), checked against the oracle with vectorised C calls:
every recorded transition is a legal REF-rules move, visit vectors are consistent with the position, labels agree
with the terminal position, and the engine's switches (cache / sharing) do not change a byte."""
import numpy as np
import pytest
import torch

from oracle import cref

pytestmark = pytest.mark.gpu


def _check_campaign(smp, sims, threshold, n_games):
    smp = smp[np.lexsort((smp["ply"], smp["game"]))]
    S, O, L = smp["self_b"], smp["opp_b"], smp["legal"]
    vis = smp["visits"].astype(np.int64)
    assert sorted(set(smp["game"].tolist())) == list(range(n_games))
    first = np.r_[True, smp["game"][1:] != smp["game"][:-1]]
    last = np.r_[first[1:], True]
    assert (smp["ply"][first] == 0).all() and (S[first] == cref.START_SELF).all() and (O[first] == cref.START_OPP).all()
    assert (np.diff(smp["ply"])[~first[1:]] == 1).all()
    # the planes the trainer will see: legal mask is the oracle's, visits only on legal actions and they add up
    assert np.array_equal(L, cref.legal_batch(S, O))
    bits = ((L[:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    is_pass = L == 0
    assert (vis.sum(axis=1) == sims).all()
    assert not (vis[:, :64][~bits] > 0).any()
    assert (vis[~is_pass, 64] == 0).all() and (vis[is_pass, 64] == sims).all()
    assert np.array_equal(smp["n_children"], np.where(is_pass, 1, bits.sum(axis=1)))
    # transitions: the move played is the square that got occupied (or a pass); replaying it with the oracle gives the next record
    nxt = np.flatnonzero(~last)
    occ0 = S[nxt] | O[nxt]; occ1 = S[nxt + 1] | O[nxt + 1]
    placed = occ1 & ~occ0
    action = np.where(placed == 0, 64, np.log2(np.maximum(placed, 1).astype(np.float64)).astype(np.int64))
    assert ((placed & (placed - np.uint64(1))) == 0).all()                  # at most one new disc
    s2, o2, _, ok = cref.make_move_batch(S[nxt], O[nxt], np.zeros(nxt.size, np.int32), action.astype(np.int32))
    assert ok.all() and np.array_equal(s2, S[nxt + 1]) and np.array_equal(o2, O[nxt + 1])
    assert (vis[nxt, action] > 0).all()                                     # a sampled / arg-max move has visits
    late = smp["ply"][nxt] >= threshold
    assert (action[late] == vis[nxt][late].argmax(axis=1)).all()            # arg-max after the temperature threshold
    # labels: winner at the terminal position (side to move there) x player parity (parallel_self_play.py:397-404)
    li = np.flatnonzero(last)
    fin_s, fin_o, fin_w = [], [], []
    for i in li:                                                            # the last move is not recorded: try the visited children
        s, o = int(S[i]), int(O[i])
        ends = set()
        for a in np.flatnonzero(vis[i] > 0):
            ok1, s3, o3, _ = cref.make_move(s, o, 0, int(a))
            if ok1 and cref.is_terminal(s3, o3):
                ends.add(cref.winner(s3, o3))
        assert ends
        fin_w.append(ends)
    game_of = smp["game"]
    for k, i in enumerate(li):
        g = game_of[i]
        rec = smp[game_of == g] if n_games <= 512 else None
        if rec is None:
            continue
        labels = rec["value"].astype(np.int64)
        assert any((labels == np.array([w * (1 if j % 2 == 0 else -1) for j in range(rec.size)])).all() for w in fin_w[k])
    assert set(np.unique(smp["value"]).tolist()) <= {-1, 0, 1}
    return smp


def test_baseline_sized_campaign_properties(ctx):
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.net import OthelloResNet
    torch.manual_seed(42)
    model = OthelloResNet(10, 128).eval()                                   # BASELINE config 3/4 network
    G = 4736
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", num_simulations=50, temperature_threshold=15,
                                   num_parallel_games=16, concurrent_games=G, seed=77, verbose=False)
    smp = w.execute_episodes_packed(G)
    st = w.last_stats
    assert 55 * G < smp.size < 70 * G
    assert 2700 < st["nn_evals"] / G < 3100                                 # the reference measured 2,882 per game
    # 4,736 slots: the auto schedule is run-until-miss (every slot runs its own search, the cache removes the repeats)
    assert st["schedule"] == "async" and st["searches_run"] == smp.size
    assert st["nn_positions"] < 0.5 * st["nn_evals"]
    assert st["network_launches"] < 51 * 60                                 # lock-step needs (1 + sims) launches per ply of the longest game
    _check_campaign(smp, 50, 15, G)
    # the same campaign with every sharing switch off: identical bytes (checked on a smaller slice to keep it short)
    kw = dict(num_simulations=50, temperature_threshold=15, num_parallel_games=16, concurrent_games=384, seed=78, verbose=False)
    wa = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", schedule="lockstep", **kw)
    a = wa.execute_episodes_packed(384)
    assert wa.last_stats["schedule"] == "lockstep" and wa.last_stats["searches_run"] < a.size     # identical roots share a search
    b = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", eval_cache=False, share_searches=False,
                                   **kw).execute_episodes_packed(384)
    a = _check_campaign(a, 50, 15, 384)
    b = b[np.lexsort((b["ply"], b["game"]))]
    assert a.tobytes() == b.tobytes()


def test_strong_config_campaign_properties(ctx):
    """BASELINE config 4 (strong_8x8.yaml:29-37): 100 simulations per move, c_puct 1.5, temperature threshold 20,
    4096 concurrent games per GPU -- the same oracle-checked invariants as the default configuration."""
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.net import OthelloResNet
    torch.manual_seed(42)
    model = OthelloResNet(10, 128).eval()
    G = 4096
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, model, "cuda", num_simulations=100, c_puct=1.5, temperature_threshold=20,
                                   num_parallel_games=16, concurrent_games=G, seed=5, verbose=False)
    smp = w.execute_episodes_packed(G)
    st = w.last_stats
    assert 55 * G < smp.size < 70 * G
    assert 5300 < st["nn_evals"] / G < 6000                                 # the reference measured 5,617 per game at 100 sims
    _check_campaign(smp, 100, 20, G)


def test_symmetry_augmentation_matches_board_symmetries(ctx):
    import othello_reinforcement_learning_test_b200 as pkg
    w = pkg.ParallelSelfPlayWorker(pkg.OthelloBitboard, None, "cuda", num_simulations=8, num_parallel_games=2, seed=2, verbose=False)
    data = w.execute_episodes(2)[:12]
    assert pkg.augment_data_with_symmetries(data, pkg.OthelloBitboard) == data        # the reference's stub: unchanged
    full = pkg.augment_data_with_symmetries(data, pkg.OthelloBitboard, full=True)
    assert len(full) == 8 * len(data)
    for i, (st, pol, val) in enumerate(data):
        b = pkg.OthelloBitboard()
        w8 = (1 << np.arange(64, dtype=np.uint64))
        b.self_board = int((st[0].reshape(64).astype(np.uint64) * w8).sum(dtype=np.uint64))
        b.opp_board = int((st[1].reshape(64).astype(np.uint64) * w8).sum(dtype=np.uint64))
        sym = b.get_symmetries(pol)                                                    # bitboard.pyx:338-370
        for k in range(8):
            assert np.array_equal(full[8 * i + k][0], sym[k][0]) and np.array_equal(full[8 * i + k][1], sym[k][1])
            assert full[8 * i + k][2] == val


def _planes_to_bits(planes):
    w8 = (1 << np.arange(64, dtype=np.uint64))
    return ((np.asarray(planes).reshape(-1, 3, 64) > 0.5).astype(np.uint64) * w8).sum(axis=2, dtype=np.uint64)


def test_get_symmetries_equals_the_compiled_reference(ctx, golden_symmetry):
    """R-BB8: OthelloBitboard.get_symmetries (bitboard.pyx:338-370) -- the 8 dihedral images in the reference's own order,
    compared with what the compiled reference board returned for 64 positions and random policies
    (tests/golden/symmetry_ref.npz, oracle/gen_golden.py gen_symmetries); and the full augmentation built from it."""
    import othello_reinforcement_learning_test_b200 as pkg
    g = golden_symmetry
    data = []
    for i in range(64):
        b = pkg.OthelloBitboard(); b.self_board = int(g["self_b"][i]); b.opp_board = int(g["opp_b"][i])
        sym = b.get_symmetries(g["pi"][i])
        assert len(sym) == 8
        for k, (pl, p) in enumerate(sym):
            assert pl.shape == (3, 8, 8) and pl.dtype == np.float32 and p.shape == (65,)
            assert np.array_equal(_planes_to_bits(pl)[0], g["planes"][i, k]), (i, k)
            assert np.array_equal(p.astype(np.float32), g["policy"][i, k]), (i, k)
        data.append((b.get_tensor_input(), g["pi"][i], 1.0 if i % 2 else -1.0))
    full = pkg.augment_data_with_symmetries(data, pkg.OthelloBitboard, full=True)
    assert len(full) == 8 * 64
    for i in range(64):
        for k in range(8):
            st, p, v = full[8 * i + k]
            assert np.array_equal(_planes_to_bits(st)[0], g["planes"][i, k]) and np.array_equal(p, g["policy"][i, k]) and v == data[i][2]
