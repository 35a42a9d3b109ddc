"""End-to-end plumbing test (BASELINE.json config 1 analogue): config/mini.yml's network (16 filters x 1
block, generic CUDA kernel) and play settings through the SelfPlayWorker mirror on the GPU: play_data files
in the reference's format (loadable by the trainer's convert_to_training_data logic), GGF lines, game-index
file, file pruning, and replay parity of every recorded position through the oracle env."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import bitboard as ob
from reversi_zero_b200.config import Config
from reversi_zero_b200.worker.self_play import SelfPlayWorker
from reversi_zero_b200.agent.player import ReversiPlayer
from reversi_zero_b200.env.reversi_env import ReversiEnv, Player

pytestmark = pytest.mark.gpu


def mini_config(tmp_path):
    cfg = Config(project_dir=str(tmp_path), data_dir=str(tmp_path / "data"))
    # config/mini.yml:3-33 values
    cfg.model.update(dict(cnn_filter_num=16, cnn_filter_size=3, res_layer_num=1, l2_reg=0.0001, value_fc_size=16))
    cfg.play.update(dict(simulation_num_per_move=10, share_mtcs_info_in_self_play=True, reset_mtcs_info_per_game=3, thinking_loop=2,
                         required_visit_to_decide_action=40, start_rethinking_turn=10, c_puct=5, change_tau_turn=10,
                         parallel_search_num=4, allowed_resign_turn=10, use_solver_turn=0, use_solver_turn_in_simulation=0,
                         schedule_of_simulation_num_per_move=[[0, 50]]))   # sim_per_move = 50
    cfg.play_data.update(dict(multi_process_num=1, nb_game_in_file=2, max_file_num=3, save_policy_of_tau_1=True, enable_ggf_data=True,
                              nb_game_in_ggf_file=2))
    cfg.b200.games_per_gpu = 8
    cfg.opts.new = True
    return cfg


def test_mini_selfplay_files(tmp_path):
    cfg = mini_config(tmp_path)
    cfg.b200.write_play_rows = True   # compact twin of every play_*.json for the device-side trainer ingest
    w = SelfPlayWorker(cfg)
    n = w.start(max_games=10)
    assert n >= 10
    files = sorted(glob.glob(os.path.join(cfg.resource.play_data_dir, "play_*.json")))
    assert 1 <= len(files) <= cfg.play_data.max_file_num        # pruned to max_file_num (self_play.py:209-217)
    assert int(open(cfg.resource.self_play_game_idx_file).read()) == n
    assert os.path.exists(cfg.resource.model_best_blob_path)     # `--new` saved the random-init weights
    n_rec = 0
    for f in files:
        data = json.load(open(f))
        assert len(data) % 8 == 0
        # the trainer's loader (worker/optimize.py:215-231)
        for state, policy, z in data:
            own = ob.bit_to_array(state[0], 64); enemy = ob.bit_to_array(state[1], 64)
            assert not np.any(own & enemy) and len(policy) == 64 and z in (1, 0, -1)
            assert abs(sum(policy) - 1) < 1e-9
            legal = ob.find_correct_moves(state[0], state[1])
            assert all((legal >> i) & 1 for i, p in enumerate(policy) if p > 0)   # visits only on legal moves
            n_rec += 1
        # records 0..7 of a ply are its 8 symmetries in the reference's order (player.py:166-179)
        o0, e0 = data[0][0]
        assert data[1][0] == [ob.rotate90(o0), ob.rotate90(e0)] and data[4][0] == [ob.flip_vertical(o0), ob.flip_vertical(e0)]
    assert n_rec > 0
    # trainer-side ingest: the row file next to each JSON file expands, on the device, to exactly the arrays
    # the reference trainer builds from the JSON (worker/optimize.py:215-231); pruning removed the twins of pruned files
    from reversi_zero_b200.worker import ingest as zi
    assert sorted(glob.glob(os.path.join(cfg.resource.play_data_dir, "play_*.rzrows"))) == [zi.rows_path_of(f) for f in files]
    for f in files:
        rows, tau1, ctt = zi.read_play_rows(zi.rows_path_of(f))
        states, policy, z = zi.to_training_arrays(rows, tau1, ctt)
        data = json.load(open(f))
        assert len(data) == len(states) and (tau1, ctt) == (True, 10)
        ref_states = np.array([[ob.bit_to_array(st[0], 64).reshape(8, 8), ob.bit_to_array(st[1], 64).reshape(8, 8)] for st, _, _ in data])
        assert np.array_equal(states, ref_states)
        assert np.array_equal(policy, np.array([p for _, p, _ in data]).astype(np.float32))
        assert np.array_equal(z, np.array([zz for _, _, zz in data]).astype(np.float32))
    ds = zi.load_play_data_dir(cfg.resource.play_data_dir, 0)
    assert ds[0].shape[0] == n_rec and ds[0].is_cuda and ds[1].shape == (n_rec, 64)
    ggf = glob.glob(os.path.join(cfg.resource.self_play_ggf_data_dir, "*.ggf"))
    assert ggf and all(line.startswith("(;GM[Othello]") for line in open(ggf[0]))
    st = w.engine.stats()
    assert st["games_finished"] >= 10 and st["expansions"] > 0
    # weight hot reload (agent/api.py:117-125): a new blob in the hand-off file is picked up by digest
    from reversi_zero_b200.agent import model as M
    old = w.net.digest
    assert not w.try_reload_model(force_check=True)
    planes = np.zeros((1, 2, 8, 8), np.uint8); planes[0, 0, 3, 4] = planes[0, 1, 3, 3] = 1
    p_old, _ = w.net.predict_planes(planes)
    np.save(cfg.resource.model_best_blob_path, M.weights_to_blob(cfg.model, M.build_random_weights(cfg.model, 99)))
    assert w.try_reload_model(force_check=True) and w.net.digest != old
    p_new, _ = w.net.predict_planes(planes)
    assert np.abs(p_new - p_old).max() > 1e-6
    assert w.start(max_games=2) >= 2      # keeps playing with the new weights


def test_reversi_player_mirror_plays_a_game(tmp_path):
    """ReversiPlayer.action / moves / finish_game contract (agent/player.py:71-134,357-364) driven like
    worker/evaluate.py:66-96 does."""
    cfg = mini_config(tmp_path)
    cfg.play.update(dict(simulation_num_per_move=30, thinking_loop=1, noise_eps=0, change_tau_turn=0, resign_threshold=None))
    black, white = ReversiPlayer(cfg, None, seed=1), ReversiPlayer(cfg, None, seed=2)   # None -> deterministic evaluator
    env = ReversiEnv().reset()
    while not env.done:
        own, enemy = env.get_own_and_enemy()
        pl = black if env.next_player == Player.black else white
        a = pl.action(own, enemy)
        assert (ob.find_correct_moves(own, enemy) >> a) & 1
        env.step(a)
    # CallbackInMCTS (agent/player.py:21,212-214): (q, n) reported every per_sim simulations
    from reversi_zero_b200.agent.player import CallbackInMCTS
    seen = []
    p3 = ReversiPlayer(cfg, None, seed=3)
    p3.action(0x0000001000000000, 0x0000000818080000, callback_in_mtcs=CallbackInMCTS(10, lambda q, n: seen.append(sum(n))))
    assert len(seen) == 3 and seen == sorted(seen) and seen[-1] >= 29
    black.finish_game(1); white.finish_game(-1)
    assert len(black.moves) % 8 == 0 and black.moves[0][2] == 1 and white.moves[-1][2] == -1
    assert env.turn <= 60 and env.winner is not None


def test_force_sim_beyond_the_arenas_drains_and_recreates_the_engine(tmp_path):
    """ADVICE r1 (high) on the real engine: the arenas are sized for the largest count of the schedule; a `.force-sim` value
    beyond them makes the worker let the resident games finish (rz_engine_set_max_games(1)), harvest them, and create a new
    engine sized for the new count whose game ids go on where the old one stopped -- nothing is swallowed, no game is lost."""
    cfg = mini_config(tmp_path)
    cfg.play.update(dict(thinking_loop=1, schedule_of_simulation_num_per_move=[[0, 8], [6, 12]]))
    cfg.play_data.update(dict(nb_game_in_file=1, max_file_num=1000, enable_ggf_data=False))
    w = SelfPlayWorker(cfg)
    n1 = w.start(max_games=10)
    first = w.engine
    assert first.cfg.arena_simulation_num == 12 and cfg.play.simulation_num_per_move == 12     # the schedule took effect (8 -> 12 at game 6)
    with open(cfg.resource.force_simulation_num_file, "wt") as f:
        f.write("40")
    n2 = w.start(max_games=12)
    if w.engine is first:
        # fast engine: the 12 games were over before the writer thread's request reached a control point; it stays queued and
        # the next start() applies it at its first control point
        assert w._cmds
        n2 += w.start(max_games=1)
    assert w.engine is not first and w.engine.cfg.simulation_num_per_move == 40 and w.engine.cfg.arena_simulation_num >= 40
    assert w.engine.cfg.first_game_id >= n1                       # ids continue after everything the old engine played
    assert n2 >= 12 and w.local_idx == n1 + n2
    files = glob.glob(os.path.join(cfg.resource.play_data_dir, "play_*.json"))
    assert len(files) == n1 + n2                                   # one file per game (no draws dropped: rate 0), none lost in the hand-over
    assert int(open(cfg.resource.self_play_game_idx_file).read()) == n1 + n2
