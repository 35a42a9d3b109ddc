"""Pin the CPU oracle (oracle/) against the golden vectors generated from the unmodified reference
(tests/golden/make_golden.py) and against the reference's own KATs (test/lib/test_bitboard.py:11-112,
test/agent/test_player.py:11-75)."""
import hashlib
import json
import os

import numpy as np

from oracle import bitboard as bb
from oracle import mcts, nn

U64 = np.uint64


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def test_reference_kats(golden_dir):
    # integers quoted in SURVEY §8(c) from the reference's ASCII boards
    assert bb.find_correct_moves(0x00000000081d0603, 0x0002043814020100) == 0x0000780623000000
    assert bb.find_correct_moves(0x000700542a202020, 0x0088ffabd5dfdf5f) == 0x0870000000000080
    assert bb.find_correct_moves(0x008e868ef9fffd7c, 0xfe71797106000203) == 0x0100000000000000
    assert bb.find_correct_moves(0xfe71797106000203, 0x008e868ef9fffd7c) == 0x0000000000000080
    for k in _load(golden_dir, "kats.json"):
        assert bb.find_correct_moves(k["black"], k["white"]) == k["legal_black"]
        assert bb.find_correct_moves(k["white"], k["black"]) == k["legal_white"]


def test_bitboard_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "bitboard.npz"))
    own, enemy, pos = g["own"], g["enemy"], g["pos"]
    assert np.array_equal(bb.find_correct_moves_batch(own, enemy), g["legal"])
    assert np.array_equal(bb.calc_flip_batch(pos, own, enemy), g["flip"])
    for row, i in zip(g["flip_all"], g["sub"]):
        got = bb.calc_flip_batch(np.arange(64, dtype=np.uint8), np.full(64, own[i]), np.full(64, enemy[i]))
        assert np.array_equal(got, row)
    for name, t in (("flip_vertical", 4), ("rotate90", 1), ("rotate180", 2)):
        assert np.array_equal(bb.dihedral_batch(own, np.full(own.size, t, np.uint8)), g[name])
    assert all(bb.flip_diag_a1h8(int(x)) == int(y) for x, y in zip(own[:200], g["flip_diag"][:200]))
    assert all(bb.bit_count(int(x)) == int(c) for x, c in zip(own[:500], g["bit_count"][:500]))


def test_env_playouts(golden_dir):
    g = _load(golden_dir, "env.json")
    for game in g["games"] + g["special"]:
        env = bb.Env().reset()
        assert list(env.state()) == game["states"][0]
        for i, a in enumerate(game["actions"]):
            if "legals" in game:
                own, enemy = env.own_enemy()
                assert bb.find_correct_moves(own, enemy) == game["legals"][i]
            env.step(None if a < 0 else a)
            assert list(env.state()) == game["states"][i + 1], (game["tag"], i)
    for u in g["update"]:
        assert list(bb.Env().update(*u["args"]).state()) == u["state"]
    low = g["games"][0]
    assert low["actions"][:6] == [19, 18, 17, 9, 1, 0] and low["states"][-1][5] == 2  # SURVEY §8(c) regression vector


def test_symmetry_records(golden_dir):
    g = _load(golden_dir, "symmetry.json")
    for case in g["records"]:
        got = list(mcts.symmetries8(case["own"], case["enemy"], np.array(case["policy"])))
        assert len(got) == 8
        for (o, e, p), ((ro, re), rp) in zip(got, case["records"]):
            assert (o, e) == (ro, re)
            assert list(p) == rp
    for inv in g["inverse"]:
        t = inv["flip"] * 4 + inv["rot"]
        assert bb.dihedral(inv["board"], t) == inv["transformed"]
        assert list(mcts.inverse_policy(np.arange(64), t)) == inv["src_index"]


def test_reference_test_player_orientation():
    """test/agent/test_player.py:11-75 restated on the oracle's symmetries8."""
    idx = lambda x, y: y * 8 + x
    own = (1 << idx(0, 0)) | (1 << idx(1, 1))
    enemy = (1 << idx(7, 6)) | (1 << idx(7, 7))
    policy = np.zeros(64); policy[idx(7, 0)] = 0.8; policy[idx(0, 7)] = 0.2
    recs = list(mcts.symmetries8(own, enemy, policy))
    chk = lambda b, x, y: (b >> idx(x, y)) & 1 == 1
    o, e, p = recs[1]
    assert chk(o, 7, 0) and chk(o, 6, 1) and chk(e, 0, 7) and chk(e, 1, 7) and p[idx(7, 7)] == 0.8 and p[idx(0, 0)] == 0.2
    o, e, p = recs[2]
    assert chk(o, 7, 7) and chk(o, 6, 6) and chk(e, 0, 0) and chk(e, 0, 1) and p[idx(0, 7)] == 0.8 and p[idx(7, 0)] == 0.2
    o, e, p = recs[5]
    assert chk(o, 0, 0) and chk(o, 1, 1) and chk(e, 6, 7) and chk(e, 7, 7) and p[idx(0, 7)] == 0.8 and p[idx(7, 0)] == 0.2


def _oracle_game(sims, share, k=1, noise=0.0, tau=0, seed=7):
    pp = mcts.PlayParams(simulation_num_per_move=sims, parallel_search_num=k, noise_eps=noise, change_tau_turn=tau,
                         c_puct=5, thinking_loop=1, resign_threshold=None, share_mtcs_info_in_self_play=share)
    return mcts.SelfPlayGame(pp, nn.FakeNetAPI(), seed=seed, game_id=0).play()


def test_mcts_exact_vs_reference(golden_dir):
    """parallel_search_num = 1, deterministic evaluator, tau = 0: the reference is deterministic and the
    oracle must reproduce its root visit counts, W sums, moves and training records exactly."""
    g = _load(golden_dir, "mcts.json")
    for name in ("k1_s30_shared", "k1_s12_separate"):
        ref = g[name]
        game = _oracle_game(ref["sims"], ref["share"])
        assert len(game.plies) == len(ref["plies"])
        for mine, theirs in zip(game.plies, ref["plies"]):
            assert (mine["pid"], mine["own"], mine["enemy"]) == (theirs["pid"], theirs["own"], theirs["enemy"])
            assert list(mine["N"]) == theirs["N"], name
            assert mine["action"] == theirs["action"]
            assert abs(mine["q"] - theirs["q"]) < 1e-6
        assert game.black_z == ref["z"]
        recs = game.records()
        assert len(recs) == ref["n_records"]
        norm = [[[int(o), int(e)], [float(x) for x in p], int(z)] for (o, e), p, z in recs]
        assert norm[:16] == ref["records_head"]
        assert hashlib.sha256(json.dumps(norm).encode()).hexdigest() == ref["records_sha256"]
        assert game.n_expand == ref["expansions"]


def test_mcts_statistical_vs_reference(golden_dir):
    """K = 8 + Dirichlet noise: compare mean root visit fractions (reference: 150 repetitions)."""
    ref = _load(golden_dir, "mcts.json")["k8_noise_stat"]
    pp = mcts.PlayParams(simulation_num_per_move=ref["sims"], parallel_search_num=8, noise_eps=0.25, change_tau_turn=0,
                         c_puct=ref["c_puct"], resign_threshold=None)
    acc = np.zeros(64)
    reps = 150
    for r in range(reps):
        game = mcts.SelfPlayGame(pp, nn.FakeNetAPI(), seed=99, game_id=r)
        game.search(ref["own"], ref["enemy"], 1)
        n = game.table[(ref["own"], ref["enemy"])].N
        acc += n / n.sum()
    mine, theirs = acc / reps, np.array(ref["mean_visit_frac"])
    assert set(np.nonzero(mine)[0]) == set(np.nonzero(theirs)[0])
    assert np.abs(mine - theirs).max() < 0.04, np.abs(mine - theirs).max()


def test_oracle_nn_shapes():
    from reversi_zero_b200.agent import model as M
    mc = M.ModelConfig(cnn_filter_num=16, res_layer_num=1, value_fc_size=16)
    w = M.build_random_weights(mc, 1, perturb_bn=True)
    assert M.blob_to_weights(mc, M.weights_to_blob(mc, w)).keys() == w.keys()
    planes = nn.planes_from_bitboards([0x0000000810000000, 3], [0x0000001008000000, 12])
    p, v = nn.forward(w, planes, 1)
    assert p.shape == (2, 64) and v.shape == (2,) and np.allclose(p.sum(1), 1, atol=1e-5)
    assert planes[0, 0, 3, 4] == 1 and planes[0, 1, 3, 3] == 1  # bit 28 = (y=3,x=4)


def test_solver_vs_reference(golden_dir):
    """oracle/solver.py against the reference's Cython solver (lib/alt/reversi_solver_cython.pyx): the KATs of
    lib/reversi_solver.py:102-154 (q1 -> (57, +2), q2 -> (4|14, -2), q3 -> (3, +2)) and 120 random endgames, exact + WLD."""
    from oracle.solver import Solver
    g = _load(golden_dir, "solver.json")
    kat = {c["tag"]: c for c in g["positions"] if c["tag"].startswith("q")}
    assert (kat["q1"]["move"], kat["q1"]["score"]) == (57, 2) and kat["q2"]["score"] == -2 and kat["q2"]["move"] in (4, 14)
    assert (kat["q3"]["move"], kat["q3"]["score"]) == (3, 2)
    for c in g["positions"]:
        own, enemy = (c["black"], c["white"]) if c["next_player"] == 1 else (c["white"], c["black"])
        assert Solver().solve(own, enemy, c["exactly"]) == (c["move"], c["score"]), c["tag"]


def test_mcts_with_solver_exact_vs_reference(golden_dir):
    """whole games with the solver hooks on (agent/player.py:100-103,237-251), K = 1: every ply (searched or solved),
    root visit counts, training records and the number of network evaluations equal the reference's."""
    g = _load(golden_dir, "solver.json")["mcts"]
    for name, ref in g.items():
        pp = mcts.PlayParams(simulation_num_per_move=ref["sims"], parallel_search_num=1, noise_eps=0.0, change_tau_turn=0, c_puct=5,
                             thinking_loop=1, resign_threshold=None, use_solver_turn=ref["use_solver_turn"],
                             use_solver_turn_in_simulation=ref["use_solver_turn_in_simulation"])
        game = mcts.SelfPlayGame(pp, nn.FakeNetAPI(), seed=7, game_id=0).play()
        mine = sorted(game.plies + game.solved_plies, key=lambda r: r["turn"])
        assert len(mine) == len(ref["plies"]), name
        for a, b in zip(mine, ref["plies"]):
            assert (a["pid"], a["own"], a["enemy"], a["action"]) == (b["pid"], b["own"], b["enemy"], b["action"]), (name, a["turn"])
            if "N" in a:
                assert list(a["N"]) == b["N"], (name, a["turn"])
            assert abs(a["q"] - b["q"]) < 1e-6 and a["n"] == b["n"]
        recs = [[[int(o), int(e)], [float(x) for x in p], int(z)] for (o, e), p, z in game.records()]
        assert hashlib.sha256(json.dumps(recs).encode()).hexdigest() == ref["records_sha256"], name
        assert game.n_expand == ref["expansions"] and game.black_z == ref["z"]


def test_mcts_decision_features_exact_vs_reference(golden_dir):
    """whole games of the reference player with rethinking loops (agent/player.py:105-118), the resign rule (:123-130,
    resignation enabled), separate tables and the solver hooks -- one at a time and all together (K = 1, tau = 0,
    deterministic evaluator): every ply, every root visit vector, the resigned flags, the result, the training records
    and the number of network evaluations equal the reference's."""
    g = _load(golden_dir, "mcts_features.json")
    for name, ref in g.items():
        if "games" in ref:
            continue      # test_mcts_kept_table_exact_vs_reference
        kw = dict(simulation_num_per_move=ref["sims"], parallel_search_num=1, noise_eps=0.0, change_tau_turn=0, c_puct=5, thinking_loop=1,
                  resign_threshold=None, share_mtcs_info_in_self_play=ref["share"], use_solver_turn=0, use_solver_turn_in_simulation=0)
        kw.update(ref["play"])
        game = mcts.SelfPlayGame(mcts.PlayParams(**kw), nn.FakeNetAPI(), seed=7, game_id=0)
        game.enable_resign = ref["enable_resign"]      # the reference decides this per game with random() (self_play.py:144)
        game.play()
        mine = sorted(game.plies + game.solved_plies, key=lambda r: r["turn"])
        theirs = [p for p in ref["plies"] if p["action"] >= 0]
        assert len(mine) == len(theirs), name
        assert len(ref["plies"]) - len(theirs) == (1 if game.actions[-1] is None else 0), name      # the resignation itself
        for a, b in zip(mine, theirs):
            assert (a["pid"], a["own"], a["enemy"], a["action"]) == (b["pid"], b["own"], b["enemy"], b["action"]), (name, a["turn"])
            if "N" in a:
                assert list(a["N"]) == b["N"], (name, a["turn"])       # includes the visits of every rethinking loop
            assert abs(a["q"] - b["q"]) < 1e-6 and a["n"] == b["n"], (name, a["turn"])
        assert (game.resigned[1], game.resigned[2]) == (ref["resigned"]["black"], ref["resigned"]["white"]), name
        assert game.black_z == ref["z"] and game.env.turn == ref["turn"], name
        recs = [[[int(o), int(e)], [float(x) for x in p], int(z)] for (o, e), p, z in game.records()]
        assert len(recs) == ref["n_records"] and hashlib.sha256(json.dumps(recs).encode()).hexdigest() == ref["records_sha256"], name
        assert game.n_expand == ref["expansions"], name


def test_mcts_kept_table_exact_vs_reference(golden_dir):
    """reset_mtcs_info_per_game = 3 (worker/self_play.py:108-134): three consecutive reference games on ONE MCTSInfo; the
    oracle is handed the previous game's table and must reproduce games 2 and 3 as well."""
    for name in ("kept_table_3_games_s14", "kept_table_solver_3_games_s14"):
        _check_kept_table_games(_load(golden_dir, "mcts_features.json")[name], name)


def _check_kept_table_games(ref, name):
    pp = mcts.PlayParams(simulation_num_per_move=ref["sims"], parallel_search_num=1, noise_eps=0.0, change_tau_turn=0, c_puct=5, thinking_loop=1,
                         resign_threshold=None, share_mtcs_info_in_self_play=True, reset_mtcs_info_per_game=3,
                         use_solver_turn=ref.get("use_solver_turn", 0), use_solver_turn_in_simulation=ref.get("use_solver_turn_in_simulation", 0))
    table = None
    for i, rg in enumerate(ref["games"]):
        game = mcts.SelfPlayGame(pp, nn.FakeNetAPI(), seed=7, game_id=i, table=table).play()
        table = game.table
        mine = sorted(game.plies + game.solved_plies, key=lambda r: r["turn"])
        assert len(mine) == len(rg["plies"]), (name, i)
        for a, b in zip(mine, rg["plies"]):
            assert (a["pid"], a["own"], a["enemy"], a["action"]) == (b["pid"], b["own"], b["enemy"], b["action"]), (name, i, a["turn"])
            if "N" in a:
                assert list(a["N"]) == b["N"], (name, i, a["turn"])
            assert abs(a["q"] - b["q"]) < 1e-6 and a["n"] == b["n"], (name, i, a["turn"])
        recs = [[[int(o), int(e)], [float(x) for x in p], int(z)] for (o, e), p, z in game.records()]
        assert hashlib.sha256(json.dumps(recs).encode()).hexdigest() == rg["records_sha256"], i
        assert game.n_expand == rg["expansions"] and game.black_z == rg["z"] and game.env.turn == rg["turn"], i


def test_evaluation_match_exact_vs_reference(golden_dir):
    """worker/evaluate.py:66-96 run UNMODIFIED with two deterministic evaluators (the challenger's value negated) behind
    the reference's own ReversiModelAPI: every ply of the games -- position, move, root visit counts, solved plies, the
    resignation -- and the verdict equal the oracle's two-network game (api / api_b / black_net), which the engine's
    rz_engine_set_second_net path reproduces exactly on the GPU (tests/test_engine_gpu.py).  Variant "default_solver" is
    what an evaluation game does when nobody touches the solver settings: exact root solver from the eval play_config,
    WLD solver in simulations from the SELF-PLAY section (agent/player.py:100,237-238)."""
    for variant, g in _load(golden_dir, "eval_match.json").items():
        pl = g["play"]
        pp = mcts.PlayParams(simulation_num_per_move=pl["simulation_num_per_move"], parallel_search_num=pl["parallel_search_num"], c_puct=pl["c_puct"],
                             noise_eps=pl["noise_eps"], change_tau_turn=pl["change_tau_turn"], thinking_loop=pl["thinking_loop"],
                             resign_threshold=pl["resign_threshold"], allowed_resign_turn=pl["allowed_resign_turn"], disable_resignation_rate=0,
                             share_mtcs_info_in_self_play=False, use_solver_turn=pl["use_solver_turn"],
                             use_solver_turn_in_simulation=pl["use_solver_turn_in_simulation"])
        for i, ref in enumerate(g["games"]):
            game = mcts.SelfPlayGame(pp, nn.FakeNetAPI(), seed=3, game_id=i, api_b=nn.FakeNetAPI(sign=-1.0),
                                     black_net=0 if ref["best_is_black"] else 1)
            assert game.enable_resign                                    # disable_resignation_rate = 0 (config.py:110)
            game.play()
            mine = sorted(game.plies + game.solved_plies, key=lambda r: r["turn"])
            moves = [p for p in ref["plies"] if p["action"] >= 0]
            assert len(mine) == len(moves), (variant, i)
            for a, b in zip(mine, moves):
                assert (a["own"], a["enemy"], a["action"]) == (b["own"], b["enemy"], b["action"]), (variant, i, a["turn"])
                if "N" in a:
                    assert list(a["N"]) == b["N"], (variant, i, a["turn"])
            assert (game.actions[-1] is None) == (ref["plies"][-1]["action"] < 0), (variant, i)
            e = game.env
            assert [bin(e.black).count("1"), bin(e.white).count("1")] == ref["score"], (variant, i)
            ng_is_black = not ref["best_is_black"]
            ng_win = None if e.winner == 3 else int((e.winner == 1) == ng_is_black)
            assert ng_win == ref["ng_win"], (variant, i)


def test_mcts_k8_positions_vs_reference(golden_dir):
    """K = 8 + Dirichlet noise at every root selection: 12 roots from the opening to the endgame (tests/golden/mcts_k8.json,
    the UNMODIFIED ReversiPlayer, 100 repetitions each).  The reference's interleaving of its eight coroutines depends on
    asyncio timers, the oracle's wave model is deterministic, so the comparison is statistical: mean root visit fractions,
    mean Q of the most visited move, mean network evaluations per search."""
    ref = _load(golden_dir, "mcts_k8.json")["positions"]
    assert len(ref) >= 10
    reps = 50
    for p in ref:
        pp = mcts.PlayParams(simulation_num_per_move=p["sims"], parallel_search_num=8, noise_eps=0.25, change_tau_turn=0,
                             c_puct=p["c_puct"], resign_threshold=None)
        acc, q_top, rows = np.zeros(64), 0.0, 0
        for r in range(reps):
            game = mcts.SelfPlayGame(pp, nn.FakeNetAPI(), seed=1234, game_id=r)
            game.search(p["own"], p["enemy"], 1)
            node = game.table[(p["own"], p["enemy"])]
            acc += node.N / node.N.sum()
            a = int(np.argmax(node.N))
            q_top += float(node.W[a] / (node.N[a] + 1e-5))
            rows += game.n_expand
        mine, theirs = acc / reps, np.array(p["mean_visit_frac"])
        assert set(np.nonzero(mine)[0]) == set(np.nonzero(theirs)[0]), p["turn"]
        assert np.abs(mine - theirs).max() < 0.03, (p["turn"], np.abs(mine - theirs).max())
        assert abs(q_top / reps - p["mean_q_of_most_visited"]) < 0.05, (p["turn"], q_top / reps, p["mean_q_of_most_visited"])
        assert abs(rows / reps - p["mean_expansions"]) <= 0.03 * p["mean_expansions"] + 0.5, (p["turn"], rows / reps, p["mean_expansions"])


def test_mcts_k8_whole_game_distributions_vs_reference(golden_dir):
    """whole self-play games at K = 8 (50 simulations per move, tau turn 4, noise): plies per game, network evaluations per
    game, final disc difference and the winner split of the oracle against 80 games of the unmodified reference loop
    (worker/self_play.py:139-175 + agent/player.py) -- the quantities the throughput metric is built on."""
    ref = _load(golden_dir, "mcts_k8.json")["games"]
    g = ref["games"]
    pp = mcts.PlayParams(simulation_num_per_move=ref["sims"], parallel_search_num=8, noise_eps=ref["noise_eps"],
                         change_tau_turn=ref["change_tau_turn"], c_puct=ref["c_puct"], resign_threshold=None)
    n = 40
    mine = [mcts.SelfPlayGame(pp, nn.FakeNetAPI(), seed=4321, game_id=i).play() for i in range(n)]

    def stats(plies, exps, dd):
        return np.mean(plies), np.mean(exps), np.mean(dd), np.std(dd)
    m_pl, m_ex, m_dd, s_dd = stats([len(x.plies) for x in mine], [x.n_expand for x in mine],
                                   [bin(x.env.black).count("1") - bin(x.env.white).count("1") for x in mine])
    r_pl, r_ex, r_dd, r_sd = stats([x["plies"] for x in g], [x["expansions"] for x in g], [x["disc_diff"] for x in g])
    se_pl = np.std([x["plies"] for x in g]) * np.sqrt(1 / n + 1 / len(g))
    se_ex = np.std([x["expansions"] for x in g]) * np.sqrt(1 / n + 1 / len(g))
    assert abs(m_pl - r_pl) < 4 * se_pl + 0.5, (m_pl, r_pl)
    assert abs(m_ex - r_ex) < 4 * se_ex, (m_ex, r_ex)                       # expansions per game: what games/s is derived from
    assert abs(m_dd - r_dd) < 4 * r_sd * np.sqrt(1 / n + 1 / len(g)), (m_dd, r_dd)
    assert 0.6 < s_dd / r_sd < 1.6
