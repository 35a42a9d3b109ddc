"""GPU parity tests of the on-device MCTS engine against the CPU oracle (oracle/mcts.py, itself pinned
to the unmodified reference) with the deterministic evaluator: root statistics, whole games (moves,
visit counts, results) and the play_data file must be IDENTICAL; every engine game is additionally
replayed through the oracle's ReversiEnv restatement (replay parity)."""
import json
import os

import numpy as np
import pytest

from oracle import bitboard as ob
from oracle import mcts, nn as onn
from reversi_zero_b200 import engine as E

pytestmark = pytest.mark.gpu


def params(**kw):
    base = dict(simulation_num_per_move=40, parallel_search_num=8, noise_eps=0.0, change_tau_turn=4, c_puct=5,
                thinking_loop=1, resign_threshold=None, share_mtcs_info_in_self_play=True)
    base.update(kw)
    return mcts.PlayParams(**base)


def make_engine(pp, games, seed, **kw):
    kw.setdefault("max_searches_per_game", 60 * pp.thinking_loop)  # arena large enough that rethinking is never skipped
    cfg = E.engine_cfg_from_play_config(pp, games=games, seed=seed, eval_mode=E.EVAL_FAKE, **kw)
    return E.Engine(cfg)


@pytest.mark.parametrize("k,sims", [(1, 30), (8, 100), (4, 57)])
def test_search_root_exact(k, sims):
    own, enemy = 0x00000000081d0603, 0x0002043814020100
    pp = params(parallel_search_num=k, simulation_num_per_move=sims)
    eng = make_engine(pp, games=3, seed=11)
    for slot in range(3):
        n, w = eng.search_root(own, enemy, 1, slot)
        game = mcts.SelfPlayGame(pp, onn.FakeNetAPI(), seed=11, game_id=slot)
        game.search(own, enemy, 1)
        node = game.table[(own, enemy)]
        assert list(n) == list(node.N), (k, sims, slot)
        assert np.array_equal(w, node.W)
    eng.close()


def replay_check(g):
    """replay the engine's move list through the oracle env: every intermediate state must agree."""
    env = ob.Env().reset()
    for ply in g["plies"]:
        own, enemy = env.own_enemy()
        assert (ply["own"], ply["enemy"], ply["pid"]) == (own, enemy, env.next_player)
        legal = ob.find_correct_moves(own, enemy)
        assert all((legal >> int(a)) & 1 for a in np.nonzero(ply["N"])[0])
        assert ply["action"] == -1 or (legal >> ply["action"]) & 1
        env.step(None if ply["action"] < 0 else ply["action"])
    assert env.done and env.winner == g["winner"] and (env.black, env.white) == (g["black"], g["white"])
    assert env.turn == g["turn"]


@pytest.mark.parametrize("kw", [dict(), dict(share_mtcs_info_in_self_play=False, simulation_num_per_move=24),
                                dict(thinking_loop=3, required_visit_to_decide_action=30, start_rethinking_turn=2,
                                     simulation_num_per_move=20),
                                dict(resign_threshold=-0.3, allowed_resign_turn=10, disable_resignation_rate=0.5,
                                     simulation_num_per_move=20, change_tau_turn=60)])
def test_full_games_exact(kw, tmp_path):
    pp = params(**kw)
    n_games = 6
    eng = make_engine(pp, games=4, seed=21, max_games=n_games, overlap_groups=2)  # 4 slots (2 groups), 6 games: slots 0,1 play two games each
    eng.run(finished_target=n_games)
    raw_games, ng, raw_plies, _ = eng.poll_raw()
    assert ng == n_games
    path = str(tmp_path / "play_test.json")
    nrec = E.write_play_data(path, raw_games, ng, raw_plies, pp.save_policy_of_tau_1, pp.change_tau_turn)
    # re-read through poll()'s python view
    games = []
    for i in range(ng):
        g = raw_games[i]
        games.append(dict(game_id=int(g.game_id), winner=int(g.winner), black=int(g.black), white=int(g.white), turn=int(g.turn),
                          black_z=int(g.black_z), expansions=int(g.expansions), resigned_mask=int(g.resigned_mask),
                          plies=[dict(own=int(raw_plies[j].own), enemy=int(raw_plies[j].enemy), pid=int(raw_plies[j].player),
                                      action=int(raw_plies[j].action), N=np.array(raw_plies[j].n_visit[:]), loops=int(raw_plies[j].loops),
                                      recorded=int(raw_plies[j].recorded), q=float(raw_plies[j].q), n=float(raw_plies[j].n))
                                 for j in range(g.first_ply, g.first_ply + g.n_plies)]))
    expected_records = []
    for g in sorted(games, key=lambda x: x["game_id"]):
        replay_check(g)
        o = mcts.SelfPlayGame(pp, onn.FakeNetAPI(), seed=21, game_id=g["game_id"]).play()
        rec_plies = [p for p in g["plies"] if p["recorded"]]
        assert len(rec_plies) == len(o.plies)
        for mine, theirs in zip(rec_plies, o.plies):
            assert (mine["own"], mine["enemy"], mine["pid"]) == (theirs["own"], theirs["enemy"], theirs["pid"])
            assert list(mine["N"]) == list(theirs["N"])
            assert mine["action"] == theirs["action"] and mine["loops"] == theirs["loops"]
            assert abs(mine["q"] - theirs["q"]) < 1e-6 and mine["n"] == theirs["n"]
        assert g["black_z"] == o.black_z and g["winner"] == o.env.winner and g["expansions"] == o.n_expand
        assert g["resigned_mask"] == (1 if o.resigned[1] else 0) | (2 if o.resigned[2] else 0)
    # file content: same records, same text as json.dumps of the oracle's reference-format records
    for i in range(ng):  # file order = poll order
        g = games[i]
        o = mcts.SelfPlayGame(pp, onn.FakeNetAPI(), seed=21, game_id=g["game_id"]).play()
        expected_records += [[[int(a), int(b)], [float(x) for x in p], int(z)] for (a, b), p, z in o.records()]
    assert nrec == len(expected_records)
    text = open(path).read()
    assert text == json.dumps(expected_records)
    st = eng.stats()
    assert st["games_finished"] == n_games and st["expansions"] == sum(g["expansions"] for g in games)
    eng.close()


def test_noise_statistics(golden_dir):
    """Dirichlet root noise + K = 8: mean root visit fractions vs the reference (golden, 150 repetitions)."""
    ref = json.load(open(os.path.join(golden_dir, "mcts.json")))["k8_noise_stat"]
    pp = params(simulation_num_per_move=ref["sims"], noise_eps=0.25, c_puct=ref["c_puct"])
    eng = make_engine(pp, games=256, seed=5)
    acc = np.zeros(64)
    for slot in range(0, 256, 2):  # every call searches all slots; read a different one each time
        n, _ = eng.search_root(ref["own"], ref["enemy"], 1, slot)
        acc += n / n.sum()
    mine, theirs = acc / 128, np.array(ref["mean_visit_frac"])
    assert set(np.nonzero(mine)[0]) == set(np.nonzero(theirs)[0])
    assert np.abs(mine - theirs).max() < 0.04, np.abs(mine - theirs).max()
    eng.close()


def test_k8_positions_vs_reference(golden_dir):
    """K = 8 + Dirichlet noise on 12 roots from the opening to the endgame against the UNMODIFIED ReversiPlayer
    (tests/golden/mcts_k8.json, agent/player.py:189-215, 100 repetitions each): mean root visit fractions, mean Q of the most
    visited move and mean network evaluations per search of 64 independent engine slots."""
    ref = json.load(open(os.path.join(golden_dir, "mcts_k8.json")))["positions"]
    assert len(ref) >= 10
    slots = 64
    for p in ref:
        pp = params(simulation_num_per_move=p["sims"], noise_eps=0.25, c_puct=p["c_puct"], change_tau_turn=0)
        eng = make_engine(pp, games=slots, seed=17)
        acc, q_top = np.zeros(64), 0.0
        for slot in range(slots):   # every call searches all slots again with the same per-slot streams; read a different slot each time
            n, w = eng.search_root(p["own"], p["enemy"], 1, slot)
            acc += n / n.sum()
            a = int(np.argmax(n))
            q_top += float(w[a]) / (float(n[a]) + 1e-5)
            if slot == 0:
                exps = eng.stats()["expansions"] / slots
        reps = slots
        mine, theirs = acc / reps, np.array(p["mean_visit_frac"])
        assert set(np.nonzero(mine)[0]) == set(np.nonzero(theirs)[0]), p["turn"]
        assert np.abs(mine - theirs).max() < 0.04, (p["turn"], np.abs(mine - theirs).max())
        assert abs(q_top / reps - p["mean_q_of_most_visited"]) < 0.06, (p["turn"], q_top / reps, p["mean_q_of_most_visited"])
        assert abs(exps - p["mean_expansions"]) <= 0.03 * p["mean_expansions"] + 0.5, (p["turn"], exps, p["mean_expansions"])
        eng.close()


def test_k8_whole_game_distributions_vs_reference(golden_dir):
    """whole games at K = 8 (50 simulations per move, tau turn 4, noise 0.25) against 80 games of the unmodified reference
    loop: plies per game, network evaluations per game (the quantity games/s is derived from), final disc difference."""
    ref = json.load(open(os.path.join(golden_dir, "mcts_k8.json")))["games"]
    g = ref["games"]
    pp = params(simulation_num_per_move=ref["sims"], noise_eps=ref["noise_eps"], change_tau_turn=ref["change_tau_turn"], c_puct=ref["c_puct"])
    n = 512
    eng = make_engine(pp, games=n, seed=23, max_games=n)
    eng.run(finished_target=n)
    mine = eng.poll()
    eng.close()
    assert len(mine) == n
    for x in mine[:16]:
        replay_check(x)
    pl, ex = np.array([len(x["plies"]) for x in mine]), np.array([x["expansions"] for x in mine])
    dd = np.array([bin(x["black"]).count("1") - bin(x["white"]).count("1") for x in mine])
    r_pl, r_ex, r_dd = (np.array([x[k] for x in g]) for k in ("plies", "expansions", "disc_diff"))
    se = lambda a, b: np.sqrt(a.var() / a.size + b.var() / b.size)
    assert abs(pl.mean() - r_pl.mean()) < 4 * se(pl, r_pl) + 0.3, (pl.mean(), r_pl.mean())
    assert abs(ex.mean() - r_ex.mean()) < 4 * se(ex, r_ex), (ex.mean(), r_ex.mean())
    assert abs(dd.mean() - r_dd.mean()) < 4 * se(dd, r_dd), (dd.mean(), r_dd.mean())
    assert 0.7 < dd.std() / r_dd.std() < 1.4


def test_search_with_real_network_matches_oracle():
    """Integration of the pieces the deterministic evaluator cannot exercise: the dihedral transform of the leaf
    batch (K3), the network, and the inverse-dihedral policy gather + re-normalisation at expansion.  mini.yml-sized
    network on the fp32 generic kernel (2e-5 from torch): visit counts must be identical in almost every slot; a
    last-bit difference in a prior may flip one arg-max, so a small fraction of slots may differ slightly."""
    from reversi_zero_b200.agent import model as M
    from reversi_zero_b200 import net as N
    mc = M.ModelConfig(cnn_filter_num=16, res_layer_num=1, value_fc_size=16)
    w = M.build_random_weights(mc, 4, perturb_bn=True)
    net = N.Net(mc)
    net.load_weights(w)
    pp = params(simulation_num_per_move=48, parallel_search_num=4, noise_eps=0.0)
    cfg = E.engine_cfg_from_play_config(pp, games=24, seed=31, eval_mode=E.EVAL_NET, net_impl=N.IMPL_GENERIC)
    eng = E.Engine(cfg, net)
    own, enemy = 0x00000000081d0603, 0x0002043814020100
    same, dist = 0, []
    for slot in range(24):
        n, _ = eng.search_root(own, enemy, 1, slot)
        game = mcts.SelfPlayGame(pp, onn.OracleNetAPI(w, 1), seed=31, game_id=slot)
        game.search(own, enemy, 1)
        ref = game.table[(own, enemy)].N
        assert n.sum() == ref.sum()
        same += int(list(n) == list(ref))
        dist.append(np.abs(n - ref).sum())
    assert same >= 20, (same, dist)
    assert max(dist) <= 12, dist
    eng.close()
    net.close()


@pytest.mark.parametrize("kw", [dict(), dict(use_solver_turn=54, use_solver_turn_in_simulation=51, resign_threshold=-0.35, allowed_resign_turn=10,
                                             disable_resignation_rate=0, noise_eps=0.0)])
def test_evaluation_match_two_networks_exact(kw):
    """rz_engine_set_second_net (worker/evaluate.py:66-96): each search is evaluated by the mover's own network and the
    colours alternate with the game index.  Deterministic evaluators (the second one negates the value): whole games
    must equal the oracle driven with the same two evaluators (which equals the reference's EvaluateWorker.play_game,
    tests/test_oracle.py::test_evaluation_match_exact_vs_reference) -- also with the solver and the resign rule on, as an
    evaluation game has them by default."""
    pp = params(simulation_num_per_move=30, share_mtcs_info_in_self_play=False, change_tau_turn=0, **kw)
    eng = make_engine(pp, games=3, seed=41, max_games=6)
    eng.set_second_net(None, enable=True)
    eng.run(finished_target=6)
    games = sorted(eng.poll(), key=lambda g: g["game_id"])
    eng.close()
    assert [g["black_net"] for g in games] == [0, 1, 0, 1, 0, 1]
    for g in games:
        replay_check(g)
        o = mcts.SelfPlayGame(pp, onn.FakeNetAPI(), seed=41, game_id=g["game_id"], api_b=onn.FakeNetAPI(sign=-1.0),
                              black_net=g["black_net"]).play()
        theirs = sorted(o.plies + o.solved_plies, key=lambda r: r["turn"])
        played = [p for p in g["plies"] if p["action"] >= 0]                 # a resignation is logged as a ply with action -1
        assert [(p["own"], p["enemy"], p["action"], list(p["N"]) if p["recorded"] else None) for p in played] == \
               [(p["own"], p["enemy"], p["action"], list(p["N"]) if "N" in p else None) for p in theirs]
        assert g["winner"] == o.env.winner and (len(g["plies"]) > len(played)) == (o.actions[-1] is None)
    # the two evaluators really differ: the match is not symmetric
    assert len({(g["winner"], g["black_net"]) for g in games}) > 1 or len({g["black"] for g in games}) > 1


def test_evaluate_worker_with_real_networks(tmp_path):
    """EvaluateWorker mirror: identical weights -> a balanced match that does not promote the challenger at 0.9;
    files (best blob, next_generation dir) handled like worker/evaluate.py:33-43,115-121."""
    from reversi_zero_b200.config import Config
    from reversi_zero_b200.agent import model as M
    from reversi_zero_b200.worker import evaluate as EV
    cfg = Config(project_dir=str(tmp_path), data_dir=str(tmp_path / "data"))
    cfg.model.update(dict(cnn_filter_num=16, res_layer_num=1, value_fc_size=16))
    cfg.play.update(dict(c_puct=5, parallel_search_num=4))
    cfg.eval = dict(game_num=12, replace_rate=0.9, play_config=dict(simulation_num_per_move=16, c_puct=1))
    cfg.resource.create_directories()
    blob = M.weights_to_blob(cfg.model, M.build_random_weights(cfg.model, 1))
    np.save(cfg.resource.model_best_blob_path, blob)
    ng_dir = os.path.join(cfg.resource.next_generation_model_dir, cfg.resource.next_generation_model_dirname_tmpl % "20260922-000000.000000")
    os.makedirs(ng_dir)
    np.save(os.path.join(ng_dir, EV.NEXT_GENERATION_BLOB), blob)
    w = EV.EvaluateWorker(cfg)
    w.best_net = w._load(cfg.resource.model_best_blob_path)
    results, games = EV.play_match(cfg, w.best_net, w._load(os.path.join(ng_dir, EV.NEXT_GENERATION_BLOB)), 12)
    assert len(results) == 12 and all(r in (0, 1, None) for r in results)
    for g in games:
        replay_check(g)
    assert w.start(max_models=1) == 1 and not os.path.exists(ng_dir)      # evaluated, not promoted, directory removed


@pytest.mark.parametrize("kw", [dict(use_solver_turn=52, use_solver_turn_in_simulation=52, simulation_num_per_move=24),
                                dict(use_solver_turn=56, use_solver_turn_in_simulation=50, simulation_num_per_move=32, parallel_search_num=4),
                                dict(use_solver_turn=0, use_solver_turn_in_simulation=51, simulation_num_per_move=20),
                                # everything at once: solver at the root and inside the search, rethinking loops, separate tables
                                # per player, resignation (enabled for half of the games), late tau switch
                                dict(use_solver_turn=54, use_solver_turn_in_simulation=52, simulation_num_per_move=20, parallel_search_num=4,
                                     thinking_loop=2, required_visit_to_decide_action=30, start_rethinking_turn=2,
                                     share_mtcs_info_in_self_play=False, resign_threshold=-0.5, allowed_resign_turn=10,
                                     disable_resignation_rate=0.5, change_tau_turn=8)])
@pytest.mark.parametrize("budget_us", [None, 1])
def test_full_games_with_endgame_solver_exact(kw, budget_us, monkeypatch):
    """endgame solver hooks on the device (agent/player.py:100-103,150-161,237-251): exact root solves from
    use_solver_turn on (plies not recorded) and WLD-solved nodes inside the search; whole games equal the oracle (which
    equals the reference, tests/test_oracle.py::test_mcts_with_solver_exact_vs_reference).  Solves are resumable and
    advance for a bounded time per wave; budget_us = 1 makes every solve span many waves (slots wait, network results of
    the waiting slots are kept aside) -- the games must not change."""
    if budget_us is not None:
        monkeypatch.setenv("RZ_SOLVER_BUDGET_US", str(budget_us))
    pp = params(**kw)
    n_games = 5
    eng = make_engine(pp, games=3, seed=51, max_games=n_games)
    eng.run(finished_target=n_games)
    games = sorted(eng.poll(), key=lambda g: g["game_id"])
    eng.close()
    assert len(games) == n_games
    solved_total = 0
    for g in games:
        replay_check(g)
        o = mcts.SelfPlayGame(pp, onn.FakeNetAPI(), seed=51, game_id=g["game_id"]).play()
        theirs = sorted(o.plies + o.solved_plies, key=lambda r: r["turn"])
        played = [p for p in g["plies"] if p["action"] >= 0]     # a resignation is logged as a ply with action -1, not as a record
        assert len(played) == len(theirs)
        assert len(g["plies"]) - len(played) == (1 if g["plies"][-1]["action"] < 0 else 0)   # at most one, and it is the last ply
        for mine, ref in zip(played, theirs):
            assert (mine["own"], mine["enemy"], mine["pid"], mine["action"]) == (ref["own"], ref["enemy"], ref["pid"], ref["action"])
            assert mine["recorded"] == ("N" in ref)
            if "N" in ref:
                assert list(mine["N"]) == list(ref["N"]) and mine["loops"] == ref["loops"]
            else:
                solved_total += 1
                assert mine["n"] == 999.0 and mine["q"] == ref["q"]
        assert g["winner"] == o.env.winner and g["expansions"] == o.n_expand
        assert g["resigned_mask"] == (1 if o.resigned[1] else 0) | (2 if o.resigned[2] else 0)
    if kw["use_solver_turn"]:
        assert solved_total > 0


@pytest.mark.parametrize("solver", [False, True])
def test_reset_mtcs_info_per_game(solver):
    """PlayConfig.reset_mtcs_info_per_game = 3 (config/mini.yml:13): a slot keeps its statistics across three consecutive
    games (worker/self_play.py:111-134); whole games equal the oracle that is handed the previous game's table -- also
    with the solver hooks on, as mini.yml has them (solved positions of earlier games are met again in the kept table)."""
    pp = params(simulation_num_per_move=20, change_tau_turn=0, **(dict(use_solver_turn=54, use_solver_turn_in_simulation=51) if solver else {}))
    pp.reset_mtcs_info_per_game = 3
    n_games = 8  # 2 slots x 4 games: games 0,2,4 / 1,3,5 share a table, games 6 / 7 start a new one
    eng = make_engine(pp, games=2, seed=61, max_games=n_games)
    eng.run(finished_target=n_games)
    games = {g["game_id"]: g for g in eng.poll()}
    st = eng.stats()
    eng.close()
    assert len(games) == n_games
    for slot in range(2):
        table = None
        for k in range(4):
            gid = slot + 2 * k
            if k % 3 == 0:
                table = None
            o = mcts.SelfPlayGame(pp, onn.FakeNetAPI(), seed=61, game_id=gid, table=table).play()
            table = o.table
            g = games[gid]
            replay_check(g)
            theirs = sorted(o.plies + o.solved_plies, key=lambda r: r["turn"])
            assert [(p["own"], p["enemy"], p["action"], list(p["N"]) if p["recorded"] else None) for p in g["plies"]] == \
                   [(p["own"], p["enemy"], p["action"], list(p["N"]) if "N" in p else None) for p in theirs], gid
            assert g["expansions"] == o.n_expand and g["winner"] == o.env.winner
    assert st["max_nodes_used"] > 0


def test_full_size_selfplay_properties_and_sampled_exactness():
    """BASELINE config-2 shape -- 4096 resident games, 400 simulations per move, K = 8, Dirichlet noise, two slot groups --
    with the deterministic evaluator, every game played from the first to the last ply.  Size-independent properties
    over ALL games (ids, results, visit accounting, legality) + replay parity and EXACT equality with the oracle for a
    sample of game ids (a game's content depends only on (seed, game id), not on the 4095 games around it)."""
    pp = params(simulation_num_per_move=400, parallel_search_num=8)
    G = 4096
    eng = make_engine(pp, games=G, seed=20260922, max_games=G, overlap_groups=2, max_searches_per_game=60)
    eng.run(finished_target=G)
    games = eng.poll()
    st = eng.stats()
    eng.close()
    assert sorted(g["game_id"] for g in games) == list(range(G))                       # every id exactly once
    assert st["games_finished"] == G and st["expansions"] == sum(g["expansions"] for g in games)
    n_plies = 0
    for g in games:
        assert g["winner"] in (1, 2, 3) and g["black_z"] == {1: 1, 2: -1, 3: 0}[g["winner"]]
        nb, nw = bin(g["black"]).count("1"), bin(g["white"]).count("1")
        assert g["winner"] == (1 if nb > nw else 2 if nw > nb else 3) and not (g["black"] & g["white"])   # no resignation in this config
        for p in g["plies"]:
            n_plies += 1
            N = np.asarray(p["N"])
            legal = ob.find_correct_moves(p["own"], p["enemy"])
            turn = bin(p["own"] | p["enemy"]).count("1") - 4
            if turn == 0:
                continue                                                                 # first move is forced without a search (player.py:143-148)
            assert all((legal >> int(a)) & 1 for a in np.nonzero(N)[0]) and (legal >> p["action"]) & 1
            assert N.sum() >= 400 - 1 and N[p["action"]] > 0                              # this search's simulations (+ visits kept from earlier plies)
    assert 59 * G <= n_plies <= 61 * G
    exp = np.array([g["expansions"] for g in games])
    assert 20000 < exp.mean() < 22500                                                    # survey probe of the reference: ~21.3 k per game at S = 400
    by_id = {g["game_id"]: g for g in games}
    for gid in (0, 1234, 2047, 2048, 4095):                                             # both slot groups, first / last slots
        g = by_id[gid]
        replay_check(g)
        o = mcts.SelfPlayGame(pp, onn.FakeNetAPI(), seed=20260922, game_id=gid).play()
        assert len(g["plies"]) == len(o.plies)
        for mine, theirs in zip(g["plies"], o.plies):
            assert (mine["own"], mine["enemy"], mine["pid"], mine["action"]) == (theirs["own"], theirs["enemy"], theirs["pid"], theirs["action"])
            assert list(mine["N"]) == list(theirs["N"])
        assert g["winner"] == o.env.winner and g["expansions"] == o.n_expand
    for g in games[::64]:
        replay_check(g)
