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


@pytest.mark.parametrize("nb,nf", [(2, 32), (5, 64)])
def test_native_search_equals_reference_search_under_identical_network_outputs(ctx, golden_games, nb, nf):
    """Visit counts of the all-on-device search == visit counts of the reference algorithm (oracle)
    when the oracle is given exactly the network outputs the device network produces."""
    import othello_reinforcement_learning_test_b200 as pkg
    from othello_reinforcement_learning_test_b200.net import OthelloResNet
    torch.manual_seed(7)
    model = OthelloResNet(nb, nf).eval()
    m = pkg.MCTS(model, "cuda", c_puct=1.0)
    assert m.evaluator == "native"
    S, O = _positions(golden_games, 48, 3)
    vis, q, nev = m.search_arrays(S, O, 50)
    net = m._native_net()

    def ev(a, b):
        p, v = net.forward(np.array([a], np.uint64), np.array([b], np.uint64), out="probs")
        return p[0], float(v[0])
    for i in range(48):
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
