"""GPU: "drops in unchanged" -- the REFERENCE'S OWN caller code (byte-compiled from the reference sources into
oracle/_ref by oracle/build.py, nothing modified) runs on top of this package's classes after `dropin.install()`:

  * src/eval/arena.py Arena.play_game / play_matches (:68-202) with src/eval/players.py GreedyPlayer / MCTSPlayer
    (:79-157) -> match results equal the goldens the unmodified reference produced (tests/golden/arena_ref.npz);
  * src/train/self_play.py SelfPlayWorker.execute_episode (:52-135) over our MCTS + OthelloBitboard -> the trace equals
    the reference's own (tests/golden/selfplay_ref.npz);
  * src/train/trainer.py AlphaZeroTrainer.train (:137-250), one iteration, wired as main.py:73-140 does, with our
    ParallelSelfPlayWorker as `self_play_worker` and the reference's ReplayBuffer.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def reference_callers():
    """sys.path / sys.modules set up the way a reference checkout with the drop-in installed looks; undone afterwards."""
    from oracle import refload
    import othello_reinforcement_learning_test_b200.dropin as dropin
    try:
        root = refload.reference_root(bytecode=True)
    except RuntimeError as e:                       # oracle/_ref did not travel: nothing to drive the drop-in with
        pytest.skip(str(e))
    refload.purge_reference_modules(keep_bitboard=False)
    sys.path.insert(0, root)

    def setup(only=None, **kw):
        dropin.install(only=only, **kw)
        import importlib
        return importlib
    yield setup
    dropin.uninstall()
    refload.purge_reference_modules(keep_bitboard=False)
    sys.path.remove(root)


def _pack(results):
    return np.array([[r.winner, r.player1_score, r.player2_score, r.num_moves] for r in results], np.int32)


def test_reference_arena_and_players_run_unchanged_on_the_dropin(reference_callers):
    import othello_reinforcement_learning_test_b200 as pkg
    il = reference_callers()
    arena_mod = il.import_module("src.eval.arena")                      # the reference's module (byte-code)
    players = il.import_module("src.eval.players")
    assert arena_mod.__file__.endswith(".pyc") and arena_mod.OthelloBitboard is pkg.OthelloBitboard
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "arena_ref.npz")))
    arena = arena_mod.Arena(verbose=False)
    got = _pack(arena.play_matches(players.GreedyPlayer("A"), players.GreedyPlayer("B"), num_games=4))
    assert np.array_equal(got, g["greedy_vs_greedy"])
    for sims in (8, 25):
        mp = players.MCTSPlayer(None, torch.device("cuda"), num_simulations=sims)      # players.py:138: `from src.mcts.mcts import MCTS`
        assert isinstance(mp.mcts, pkg.MCTS) and mp.mcts.evaluator == "hashnet"        # model None -> the integer test evaluator
        got = _pack(arena.play_matches(mp, players.GreedyPlayer("G"), num_games=4))
        assert np.array_equal(got, g[f"mcts{sims}_vs_greedy"]), sims
        got = _pack(arena.play_matches(players.GreedyPlayer("G"), mp, num_games=2, alternate_colors=False))
        assert np.array_equal(got, g[f"greedy_vs_mcts{sims}"]), sims
    ev = arena_mod.evaluate_player(players.GreedyPlayer("A"), players.GreedyPlayer("B"), num_games=6, verbose=False)
    assert np.allclose([ev["win_rate"], ev["avg_score"], ev["avg_moves"]], g["evaluate_greedy"])
    # MCTSPlayer.get_action (players.py:151-157) with a real network on the tcgen05 engine
    model = pkg.OthelloResNet(5, 64).eval()
    mp = players.MCTSPlayer(model, torch.device("cuda"), num_simulations=30)
    b = pkg.OthelloBitboard()
    r = arena.play_game(mp, players.RandomPlayer("R"))                  # arena.py:68-162, one whole game
    assert r.winner in (-1, 0, 1) and 8 <= r.num_moves <= 70 and r.player1_score + r.player2_score <= 64
    assert mp.get_action(b) in b.get_legal_moves()


def test_reference_serial_selfplay_worker_runs_unchanged_on_the_dropin(reference_callers, golden_selfplay):
    import othello_reinforcement_learning_test_b200 as pkg
    il = reference_callers(only=["src.cython.bitboard", "src.mcts.mcts", "src.model.net"])
    sp = il.import_module("src.train.self_play")                        # the reference's SelfPlayWorker, not ours
    assert sp.__file__.endswith(".pyc") and sp.SelfPlayWorker is not pkg.SelfPlayWorker
    g = golden_selfplay
    for tag in ("a", "b"):
        cp, sims = float(g[f"{tag}_cfg"][0]), int(g[f"{tag}_cfg"][1])
        m = il.import_module("src.mcts.mcts").MCTS(None, torch.device("cuda"), c_puct=cp)
        assert isinstance(m, pkg.MCTS)
        w = sp.SelfPlayWorker(pkg.OthelloBitboard, m, num_simulations=sims, temperature_threshold=0)
        data = w.execute_episode(add_dirichlet_noise=False)             # self_play.py:52-135
        assert len(data) == len(g[f"{tag}_serial_value"])
        for i, (st, pol, val) in enumerate(data):
            assert np.array_equal(pol, g[f"{tag}_serial_policy"][i]) and val == float(g[f"{tag}_serial_value"][i]), (tag, i)


def test_reference_trainer_iteration_runs_unchanged_on_the_dropin(reference_callers, tmp_path, capsys):
    """main.py:73-140 wiring: reference AlphaZeroTrainer + reference ReplayBuffer, our model class, our worker factory."""
    import othello_reinforcement_learning_test_b200 as pkg
    il = reference_callers()
    trainer_mod = il.import_module("src.train.trainer")
    buffer_mod = il.import_module("src.train.buffer")
    psp = il.import_module("src.train.parallel_self_play")              # routed to this package
    net_mod = il.import_module("src.model.net")
    assert trainer_mod.__file__.endswith(".pyc") and buffer_mod.__file__.endswith(".pyc")
    assert psp.create_parallel_self_play_worker is pkg.create_parallel_self_play_worker
    config = {"model": {"num_blocks": 5, "num_filters": 64, "board_size": 8},
              "mcts": {"num_simulations": 20, "c_puct": 1.0, "dirichlet_alpha": 0.3, "dirichlet_epsilon": 0.25},
              "self_play": {"temperature_threshold": 15, "num_parallel_games": 16},
              "training": {"replay_buffer_size": 5000}}
    torch.manual_seed(42)
    np.random.seed(42)
    device = torch.device("cuda")
    model = net_mod.OthelloResNet(num_blocks=5, num_filters=64, board_size=8)
    model.to(device)
    worker = psp.create_parallel_self_play_worker(config, model, device)
    buf = buffer_mod.ReplayBuffer(max_size=config["training"]["replay_buffer_size"])
    tr = trainer_mod.AlphaZeroTrainer(model=model, device=device, replay_buffer=buf, self_play_worker=worker, config={"lr": 0.01},
                                      checkpoint_dir=str(tmp_path / "models"), log_dir=str(tmp_path / "logs"))
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tr.train(num_iterations=1, self_play_episodes_per_iter=24, train_epochs_per_iter=3, batch_size=128, checkpoint_interval=1)
    out = capsys.readouterr().out
    assert "Training Completed" in out
    assert 24 * 50 < len(buf) <= 5000 and tr.global_step == 3 and len(tr.loss_history) == 1 and np.isfinite(tr.loss_history[0])
    assert (tmp_path / "models" / "final_model.pt").exists()
    changed = [k for k, v in model.state_dict().items() if v.is_floating_point() and not torch.equal(v, before[k])]
    assert len(changed) > 50                                            # SGD really updated the module in place (trainer.py:300-326)
    # the next self-play call must see the new weights (re-fold + upload, no stale device copy)
    net = worker.batch_mcts._native_net()
    key_before = net._version_key
    data = worker.execute_episodes(num_episodes=4, add_dirichlet_noise=True)
    assert len(data) > 100 and worker.batch_mcts._native_net()._version_key == key_before   # synced inside the call, stable afterwards
