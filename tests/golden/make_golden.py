"""Generate the committed golden vectors by running the UNMODIFIED reference (imported from
/root/reference/src through oracle/ref_shims).  Runs only in the build container; the outputs
(tests/golden/*.npz, *.json) are committed and are what tests/ (CPU and GPU) compare against.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

import oracle.ref_shims.install as shims  # noqa: E402

shims.install()

from reversi_zero.lib import bitboard as rb  # noqa: E402
from reversi_zero.env.reversi_env import ReversiEnv, Player, Winner  # noqa: E402
from reversi_zero.config import Config  # noqa: E402
from reversi_zero.agent.player import ReversiPlayer  # noqa: E402
from reversi_zero.lib.util import parse_to_bitboards  # noqa: E402

SEED = 20260922
U64 = np.uint64


def random_positions(rng, n):
    """SURVEY §8(d) config-5 recipe: densities 1/4, 1/2, 3/4; own/enemy disjoint."""
    a = rng.integers(0, 2 ** 64, size=n, dtype=U64)
    b = rng.integers(0, 2 ** 64, size=n, dtype=U64)
    r = rng.integers(0, 2 ** 64, size=n, dtype=U64)
    third = n // 3
    occ = a.copy()
    occ[:third] = a[:third] & b[:third]
    occ[2 * third:] = a[2 * third:] | b[2 * third:]
    own, enemy = occ & r, occ & ~r
    pos = rng.integers(0, 64, size=n, dtype=np.uint8)
    return own, enemy, pos


def env_state(env):
    w = 0 if env.winner is None else env.winner.value
    return [int(env.board.black), int(env.board.white), env.next_player.value, env.turn, int(env.done), w]


def gen_bitboard():
    rng = np.random.default_rng(SEED)
    own, enemy, pos = random_positions(rng, 3000)
    # reachable positions from random playouts
    r_own, r_enemy = [], []
    for g in range(40):
        env = ReversiEnv().reset()
        while not env.done:
            o, e = env.get_own_and_enemy()
            r_own.append(o); r_enemy.append(e)
            legal = rb.find_correct_moves(o, e)
            moves = [i for i in range(64) if legal >> i & 1]
            env.step(int(moves[rng.integers(len(moves))]))
    r_own = np.array(r_own, dtype=U64); r_enemy = np.array(r_enemy, dtype=U64)
    own = np.concatenate([own, r_own]); enemy = np.concatenate([enemy, r_enemy])
    pos = np.concatenate([pos, rng.integers(0, 64, size=r_own.size, dtype=np.uint8)])
    legal = np.array([rb.find_correct_moves(int(o), int(e)) for o, e in zip(own, enemy)], dtype=U64)
    flip = np.array([rb.calc_flip(int(p), int(o), int(e)) for p, o, e in zip(pos, own, enemy)], dtype=U64)
    # all 64 squares on a subset (illegal squares included)
    sub = np.arange(0, own.size, 37)
    flip_all = np.array([[rb.calc_flip(p, int(own[i]), int(enemy[i])) for p in range(64)] for i in sub], dtype=U64)
    fv = np.array([rb.flip_vertical(int(x)) for x in own], dtype=U64)
    fd = np.array([rb.flip_diag_a1h8(int(x)) for x in own], dtype=U64)
    r90 = np.array([rb.rotate90(int(x)) for x in own], dtype=U64)
    r180 = np.array([rb.rotate180(int(x)) for x in own], dtype=U64)
    cnt = np.array([rb.bit_count(int(x)) for x in own], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "bitboard.npz"), own=own, enemy=enemy, pos=pos, legal=legal, flip=flip,
                        sub=sub, flip_all=flip_all, flip_vertical=fv, flip_diag=fd, rotate90=r90, rotate180=r180,
                        bit_count=cnt)
    # reference KATs, test/lib/test_bitboard.py:11-112 (ASCII boards parsed with the reference parser)
    kats = []
    boards = {
        "kat1": "##########\n#OO      #\n#XOO     #\n#OXOOO   #\n#  XOX   #\n#   XXX  #\n#  X     #\n# X      #\n#        #\n##########",
        "kat2": "##########\n#OOOOOXO #\n#OOOOOXOO#\n#OOOOOXOO#\n#OXOXOXOO#\n#OOXOXOXO#\n#OOOOOOOO#\n#XXXO   O#\n#        #\n##########",
        "kat3": "##########\n#OOXXXXX #\n#XOXXXXXX#\n#XXXXXXXX#\n#XOOXXXXX#\n#OXXXOOOX#\n#OXXOOOOX#\n#OXXXOOOX#\n# OOOOOOO#\n##########",
    }
    for name, s in boards.items():
        b, w = parse_to_bitboards(s)
        kats.append(dict(name=name, black=b, white=w, legal_black=rb.find_correct_moves(b, w),
                         legal_white=rb.find_correct_moves(w, b)))
    return kats


def gen_env():
    rng = np.random.default_rng(SEED + 1)
    games = []

    def play(chooser, tag):
        env = ReversiEnv().reset()
        actions, states, legals = [], [env_state(env)], []
        while not env.done:
            o, e = env.get_own_and_enemy()
            legal = rb.find_correct_moves(o, e)
            legals.append(int(legal))
            a = chooser(legal)
            actions.append(a)
            env.step(a)
            states.append(env_state(env))
        games.append(dict(tag=tag, actions=actions, states=states, legals=legals))

    play(lambda legal: (legal & -legal).bit_length() - 1, "lowest")
    play(lambda legal: legal.bit_length() - 1, "highest")
    for g in range(60):
        play(lambda legal: int([i for i in range(64) if legal >> i & 1][rng.integers(bin(legal).count("1"))]), f"random{g}")
    # resign (None) and illegal-move cases, reversi_env.py:49-59
    special = []
    for (nm, acts) in (("resign_black", [None]), ("resign_white", [19, None]), ("illegal_black", [0]),
                       ("illegal_white", [19, 63]), ("occupied", [27])):
        env = ReversiEnv().reset()
        sts = [env_state(env)]
        for a in acts:
            env.step(a)
            sts.append(env_state(env))
        special.append(dict(tag=nm, actions=[-1 if a is None else a for a in acts], states=sts))
    # update(): turn computation + Board() zero quirk
    upd = []
    for b, w, p in ((0x00000000081d0603, 0x0002043814020100, 1), (0x0088ffabd5dfdf5f, 0x000700542a202020, 2), (0, 5, 1)):
        env = ReversiEnv().update(b, w, Player(p))
        upd.append(dict(args=[b, w, p], state=env_state(env)))
    return dict(games=games, special=special, update=upd)


def gen_symmetry():
    rng = np.random.default_rng(SEED + 2)
    cfg = Config()
    out = []
    for i in range(6):
        own, enemy, _ = random_positions(rng, 3)
        own, enemy = int(own[i % 3]), int(enemy[i % 3])
        policy = rng.random(64)
        policy /= policy.sum()
        pl = ReversiPlayer(cfg, None, api=object())
        pl.add_data_to_move_buffer_with_8_symmetries(own, enemy, policy)
        out.append(dict(own=own, enemy=enemy, policy=list(policy),
                        records=[[[int(o), int(e)], [float(x) for x in p]] for (o, e), p in pl.moves]))
    # inverse transform of the NN policy, player.py:300-321, all 8 (flip, rot) combinations
    inv = []
    for flip in (False, True):
        for rot in range(4):
            bw = 0x00000000081d0603
            t = bw
            if flip:
                t = rb.flip_vertical(t)
            for _ in range(rot):
                t = rb.rotate90(t)
            leaf_p = np.arange(64, dtype=np.float32)  # value = index in the transformed frame
            lp = leaf_p.reshape(8, 8)
            if rot > 0:
                lp = np.rot90(lp, k=rot)
            if flip:
                lp = np.flipud(lp)
            inv.append(dict(flip=int(flip), rot=rot, board=bw, transformed=int(t), src_index=[int(x) for x in lp.reshape(64)]))
    return dict(records=out, inverse=inv)


class FakeNet:
    """Same function as oracle.nn.FakeNetAPI (kept separate so the golden run depends on nothing of ours)."""

    def __init__(self):
        self.rows = 0

    def predict(self, x):
        x = np.asarray(x)
        n = x.shape[0]
        self.rows += n
        p = np.full((n, 64), 1.0 / 64, dtype=np.float32)
        cnt = x.reshape(n, 2, 64).astype(np.int32).sum(axis=2)
        v = ((cnt[:, 0] - cnt[:, 1]).astype(np.float32) / np.float32(64)).reshape(n, 1)
        return p, v


def ref_config(sims, k, noise_eps, change_tau_turn, c_puct=5, share=True):
    cfg = Config()
    pc = cfg.play
    pc.simulation_num_per_move = sims
    pc.parallel_search_num = k
    pc.noise_eps = noise_eps
    pc.change_tau_turn = change_tau_turn
    pc.c_puct = c_puct
    pc.thinking_loop = 1
    pc.use_solver_turn = 0
    pc.use_solver_turn_in_simulation = 0
    pc.resign_threshold = None
    pc.share_mtcs_info_in_self_play = share
    return cfg


def ref_selfplay_game(cfg, api):
    """worker/self_play.py:139-175 loop, driven directly (no process pool / files)."""
    env = ReversiEnv().reset()
    info = ReversiPlayer.create_mtcs_info() if cfg.play.share_mtcs_info_in_self_play else None
    black = ReversiPlayer(cfg, None, enable_resign=False, mtcs_info=info, api=api)
    white = ReversiPlayer(cfg, None, enable_resign=False, mtcs_info=info, api=api)
    plies = []
    from reversi_zero.agent.player import CounterKey
    while not env.done:
        if env.next_player == Player.black:
            pl, own, enemy = black, env.board.black, env.board.white
        else:
            pl, own, enemy = white, env.board.white, env.board.black
        a = pl.action_with_evaluation(own, enemy)
        key = CounterKey(own, enemy, Player.black.value)
        plies.append(dict(pid=env.next_player.value, own=int(own), enemy=int(enemy), action=int(a.action),
                          N=[int(x) for x in pl.var_n[key]], W=[float(x) for x in pl.var_w[key]],
                          n=float(a.n), q=float(a.q)))
        env.step(a.action)
    z = {Winner.black: 1, Winner.white: -1, Winner.draw: 0}[env.winner]
    black.finish_game(z)
    white.finish_game(-z)
    recs = [[[int(m[0][0]), int(m[0][1])], [float(x) for x in m[1]], int(m[2])] for m in black.moves + white.moves]
    return plies, recs, z


def gen_mcts():
    out = {}
    for name, sims, share in (("k1_s30_shared", 30, True), ("k1_s12_separate", 12, False)):
        api = FakeNet()
        cfg = ref_config(sims=sims, k=1, noise_eps=0, change_tau_turn=0, share=share)
        plies, recs, z = ref_selfplay_game(cfg, api)
        import hashlib
        digest = hashlib.sha256(json.dumps(recs).encode()).hexdigest()
        out[name] = dict(sims=sims, share=share, c_puct=5, plies=plies, z=z, n_records=len(recs), records_sha256=digest,
                         records_head=recs[:16], expansions=api.rows)
    # statistical: K=8 with Dirichlet noise, one mid-game root, many repetitions
    own, enemy = 0x00000000081d0603, 0x0002043814020100
    reps, sims = 150, 100
    acc = np.zeros(64)
    np.random.seed(12345)
    for r in range(reps):
        cfg = ref_config(sims=sims, k=8, noise_eps=0.25, change_tau_turn=0)
        pl = ReversiPlayer(cfg, None, enable_resign=False, api=FakeNet())
        pl.action_with_evaluation(own, enemy)
        from reversi_zero.agent.player import CounterKey
        n = pl.var_n[CounterKey(own, enemy, 1)]
        acc += n / n.sum()
    out["k8_noise_stat"] = dict(own=own, enemy=enemy, sims=sims, reps=reps, c_puct=5, mean_visit_frac=list(acc / reps))
    return out


def gen_solver():
    """lib/alt/reversi_solver (Cython, what agent/player.py:15 imports): KATs of lib/reversi_solver.py:102-154 and
    random endgame positions, exact and WLD, a fresh solver per call."""
    from reversi_zero.lib.alt.reversi_solver import ReversiSolver
    rng = np.random.default_rng(SEED + 3)
    out = []
    kat = [("q1", 0x80feafd2eaf20200, 0x7c00502d150d0d0f, 2, False), ("q2", 0xfffeadd0c0c00000, 0x0000522f3f3f2f0f, 1, False),
           ("q3", 0x3c1e8e9e9ad0a870, 0x40607161652f1504, 2, True)]
    for name, b, w, pl, exact in kat:
        mv, sc = ReversiSolver().solve(b, w, Player(pl), exactly=exact)
        out.append(dict(tag=name, black=b, white=w, next_player=pl, exactly=exact, move=int(mv), score=int(sc)))
    n = 0
    while n < 120:
        empties = int(rng.integers(1, 10))
        env = ReversiEnv().reset()
        while not env.done and 60 - env.turn > empties:
            o, e = env.get_own_and_enemy()
            legal = rb.find_correct_moves(o, e)
            ms = [i for i in range(64) if legal >> i & 1]
            env.step(int(ms[rng.integers(len(ms))]))
        if env.done:
            continue
        for exact in (True, False):
            mv, sc = ReversiSolver().solve(env.board.black, env.board.white, env.next_player, exactly=exact)
            out.append(dict(tag=f"rand{n}", black=int(env.board.black), white=int(env.board.white), next_player=env.next_player.value,
                            exactly=exact, move=int(mv), score=int(sc)))
        n += 1
    return out


def gen_mcts_solver():
    """whole games of the reference player with the solver hooks on (agent/player.py:100-103,237-251), K = 1, tau = 0"""
    out = {}
    for name, sims, ust, usim in (("solver_s16_t52", 16, 52, 52), ("solver_s12_t50_sim54", 12, 50, 54), ("solver_s40_t56_sim51", 40, 56, 51)):
        api = FakeNet()
        cfg = ref_config(sims=sims, k=1, noise_eps=0, change_tau_turn=0, share=True)
        cfg.play.use_solver_turn = ust
        cfg.play.use_solver_turn_in_simulation = usim
        plies, recs, z = ref_selfplay_game(cfg, api)
        import hashlib
        out[name] = dict(sims=sims, use_solver_turn=ust, use_solver_turn_in_simulation=usim, c_puct=5, plies=plies, z=z, n_records=len(recs),
                         records_sha256=hashlib.sha256(json.dumps(recs).encode()).hexdigest(), expansions=api.rows)
    return out


def ref_selfplay_game_full(cfg, api, enable_resign, info=None):
    """worker/self_play.py:139-175 + :219-238 driven directly, resignation included (action None -> env.step(None)).
    ``info``: an MCTSInfo kept from earlier games (reset_mtcs_info_per_game > 1, worker/self_play.py:108-134)."""
    from reversi_zero.agent.player import CounterKey
    env = ReversiEnv().reset()
    if info is None:
        info = ReversiPlayer.create_mtcs_info() if cfg.play.share_mtcs_info_in_self_play else None
    players = {Player.black: ReversiPlayer(cfg, None, enable_resign=enable_resign, mtcs_info=info, api=api),
               Player.white: ReversiPlayer(cfg, None, enable_resign=enable_resign, mtcs_info=info, api=api)}
    plies = []
    while not env.done:
        pl = players[env.next_player]
        own, enemy = (env.board.black, env.board.white) if env.next_player == Player.black else (env.board.white, env.board.black)
        a = pl.action_with_evaluation(own, enemy)
        n_now = pl.var_n[CounterKey(own, enemy, Player.black.value)]
        plies.append(dict(pid=env.next_player.value, own=int(own), enemy=int(enemy), action=-1 if a.action is None else int(a.action),
                          N=[int(x) for x in n_now], n=float(a.n), q=float(a.q)))
        env.step(a.action)
    z = {Winner.black: 1, Winner.white: -1, Winner.draw: 0}[env.winner]
    players[Player.black].finish_game(z)
    players[Player.white].finish_game(-z)
    recs = [[[int(m[0][0]), int(m[0][1])], [float(x) for x in m[1]], int(m[2])] for m in players[Player.black].moves + players[Player.white].moves]
    return plies, recs, z, dict(black=bool(players[Player.black].resigned), white=bool(players[Player.white].resigned)), int(env.turn)


def gen_mcts_features():
    """whole games of the reference player, K = 1, tau = 0, with the per-ply decision features on: rethinking loops
    (agent/player.py:105-118), the resign rule (:123-130) with resignation enabled, separate tables, and all of them
    together with the solver hooks."""
    import hashlib
    cases = {
        "rethink_s12": dict(sims=12, share=True, play=dict(thinking_loop=3, required_visit_to_decide_action=40, start_rethinking_turn=2)),
        "resign_s16": dict(sims=16, share=True, enable_resign=True, play=dict(resign_threshold=-0.4, allowed_resign_turn=10)),
        "all_s20": dict(sims=20, share=False, enable_resign=True,
                        play=dict(thinking_loop=2, required_visit_to_decide_action=30, start_rethinking_turn=2, resign_threshold=-0.5,
                                  allowed_resign_turn=10, use_solver_turn=54, use_solver_turn_in_simulation=52)),
        "all_no_resign_s20": dict(sims=20, share=False, enable_resign=False,
                                  play=dict(thinking_loop=2, required_visit_to_decide_action=30, start_rethinking_turn=2, resign_threshold=-0.5,
                                            allowed_resign_turn=10, use_solver_turn=54, use_solver_turn_in_simulation=52)),
    }
    out = {}
    for name, c in cases.items():
        api = FakeNet()
        cfg = ref_config(sims=c["sims"], k=1, noise_eps=0, change_tau_turn=0, share=c["share"])
        for k, v in c["play"].items():
            setattr(cfg.play, k, v)
        plies, recs, z, resigned, turn = ref_selfplay_game_full(cfg, api, c.get("enable_resign", False))
        out[name] = dict(sims=c["sims"], share=c["share"], enable_resign=c.get("enable_resign", False), play=c["play"], plies=plies, z=z,
                         resigned=resigned, turn=turn, n_records=len(recs),
                         records_sha256=hashlib.sha256(json.dumps(recs).encode()).hexdigest(), expansions=api.rows)
    # reset_mtcs_info_per_game = 3 (config/mini.yml:13): three consecutive games on ONE MCTSInfo; new players start with
    # expanded = set(var_p.keys()) (agent/player.py:44-47)
    cfg = ref_config(sims=14, k=1, noise_eps=0, change_tau_turn=0, share=True)
    info = ReversiPlayer.create_mtcs_info()
    games = []
    for _ in range(3):
        api = FakeNet()
        plies, recs, z, resigned, turn = ref_selfplay_game_full(cfg, api, False, info=info)
        games.append(dict(plies=plies, z=z, turn=turn, n_records=len(recs), records_sha256=hashlib.sha256(json.dumps(recs).encode()).hexdigest(),
                          expansions=api.rows))
    out["kept_table_3_games_s14"] = dict(sims=14, games=games, table_size=len(info.var_p))
    # the same with the solver hooks on (config/mini.yml: reset_mtcs_info_per_game 3 AND use_solver_turn* 50): every game
    # has fresh ReversiSolver objects (agent/player.py:60) but meets the priors / visits earlier solves left in the table
    cfg = ref_config(sims=14, k=1, noise_eps=0, change_tau_turn=0, share=True)
    cfg.play.use_solver_turn, cfg.play.use_solver_turn_in_simulation = 54, 51
    info = ReversiPlayer.create_mtcs_info()
    games = []
    for _ in range(3):
        api = FakeNet()
        plies, recs, z, resigned, turn = ref_selfplay_game_full(cfg, api, False, info=info)
        games.append(dict(plies=plies, z=z, turn=turn, n_records=len(recs), records_sha256=hashlib.sha256(json.dumps(recs).encode()).hexdigest(),
                          expansions=api.rows))
    out["kept_table_solver_3_games_s14"] = dict(sims=14, use_solver_turn=54, use_solver_turn_in_simulation=51, games=games,
                                                table_size=len(info.var_p))
    return out


def gen_eval_match():
    """The UNMODIFIED EvaluateWorker.play_game (worker/evaluate.py:66-96) with two different deterministic evaluators
    behind the reference's own ReversiModelAPI (agent/api.py:30-45): best = FakeNet, challenger = FakeNet with the value
    negated.  The players are the reference's ReversiPlayer, wrapped only to log what they were asked and answered."""
    import types
    import reversi_zero.worker.evaluate as ev
    from reversi_zero.agent.player import CounterKey

    def model(sign):
        def predict_on_batch(x):
            p, v = FakeNet().predict(x)
            return p, v * np.float32(sign)
        return types.SimpleNamespace(model=types.SimpleNamespace(predict_on_batch=predict_on_batch))

    log = []

    class LoggingPlayer(ReversiPlayer):
        def action(self, own, enemy, callback_in_mtcs=None):
            a = super().action(own, enemy, callback_in_mtcs)
            n = self.var_n[CounterKey(own, enemy, Player.black.value)]
            log.append(dict(own=int(own), enemy=int(enemy), action=-1 if a is None else int(a), N=[int(x) for x in n],
                            resigned=bool(self.resigned)))
            return a

    out = {}
    real_player, real_random = ev.ReversiPlayer, ev.random
    # "default_solver": the solver settings an evaluation game has when nobody touches them -- the exact root solver from
    # EvaluateConfig.play_config (a fresh PlayConfig: use_solver_turn = 50) and the WLD solver inside simulations from the
    # SELF-PLAY section, which ReversiPlayer reads through self.config.play (agent/player.py:237-238)
    for variant, root_solver, sim_solver in (("default_solver", 50, 50), ("no_solver", 0, 0)):
        cfg = Config()
        pc = cfg.eval.play_config
        pc.simulation_num_per_move, pc.parallel_search_num, pc.c_puct = 24, 1, 5
        pc.use_solver_turn = root_solver
        pc.use_solver_turn_in_simulation = 12345            # never read in an evaluation game
        cfg.play.use_solver_turn_in_simulation = sim_solver
        pc.resign_threshold = -0.35
        pc.allowed_resign_turn = 12345                      # never read either: the resign rule uses config.play (agent/player.py:127)
        cfg.play.allowed_resign_turn = 10
        worker = ev.EvaluateWorker.__new__(ev.EvaluateWorker)
        worker.config = cfg
        games = []
        try:
            ev.ReversiPlayer = LoggingPlayer
            for coin in (0.25, 0.75):               # random() < 0.5 -> the best model plays black (:70)
                ev.random = lambda: coin
                del log[:]
                ng_win, best_is_black, score = worker.play_game(model(1.0), model(-1.0))
                games.append(dict(best_is_black=bool(best_is_black), ng_win=ng_win, score=[int(score[0]), int(score[1])], plies=list(log)))
        finally:
            ev.ReversiPlayer, ev.random = real_player, real_random
        out[variant] = dict(play=dict(simulation_num_per_move=24, parallel_search_num=1, c_puct=5, resign_threshold=-0.35, allowed_resign_turn=10,
                                      thinking_loop=pc.thinking_loop, change_tau_turn=pc.change_tau_turn, noise_eps=pc.noise_eps,
                                      use_solver_turn=root_solver, use_solver_turn_in_simulation=sim_solver), games=games)
    return out


def gen_ingest():
    """Trainer-side ingest: whole reference games -> the reference's own play_data file
    (lib/data_helper.py:23-25) -> its own loader + OptimizeWorker.convert_to_training_data (worker/optimize.py:215-231).
    Stored: the compact rows (one per recorded ply, black's then white's as in worker/self_play.py:183) and the arrays
    the reference trainer builds."""
    import tempfile
    from reversi_zero.lib.data_helper import write_game_data_to_file, read_game_data_from_file
    from reversi_zero.worker.optimize import OptimizeWorker
    out = {}
    for name, tau1, ctt, sims in (("tau1", True, 4, 20), ("tau_rule", False, 4, 16), ("one_hot", False, 0, 10)):
        cfg = ref_config(sims=sims, k=1, noise_eps=0, change_tau_turn=ctt)
        cfg.play_data.save_policy_of_tau_1 = tau1
        plies, recs, z = ref_selfplay_game(cfg, FakeNet())
        rows = [pl for pid in (1, 2) for pl in plies if pl["pid"] == pid]
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "play_x.json")
            write_game_data_to_file(path, recs)
            states, policies, zs = OptimizeWorker.convert_to_training_data(read_game_data_from_file(path))
        assert states.shape == (8 * len(rows), 2, 8, 8) and states.dtype == np.uint8
        out[name + "_settings"] = np.array([int(tau1), ctt], np.int32)
        out[name + "_own"] = np.array([r["own"] for r in rows], U64)
        out[name + "_enemy"] = np.array([r["enemy"] for r in rows], U64)
        out[name + "_n_visit"] = np.array([r["N"] for r in rows], np.int32)
        out[name + "_row_z"] = np.array([z if r["pid"] == 1 else -z for r in rows], np.int32)
        out[name + "_states_packed"] = np.packbits(states.reshape(len(states), -1), axis=1, bitorder="little")
        out[name + "_policy"] = policies.astype(np.float64)
        out[name + "_z"] = zs.astype(np.int64)
    return out


def main():
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "features":   # only (re)generate the decision-feature MCTS fixtures
        with open(os.path.join(HERE, "mcts_features.json"), "w") as f:
            json.dump(gen_mcts_features(), f)
        print("mcts feature golden vectors written")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "eval":   # only (re)generate the evaluation-match fixture
        with open(os.path.join(HERE, "eval_match.json"), "w") as f:
            json.dump(gen_eval_match(), f)
        print("evaluation match golden vectors written")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "ingest":   # only (re)generate the trainer-ingest fixture
        np.savez_compressed(os.path.join(HERE, "ingest.npz"), **gen_ingest())
        print("ingest golden vectors written")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "solver":   # only (re)generate the solver fixtures
        with open(os.path.join(HERE, "solver.json"), "w") as f:
            json.dump(dict(positions=gen_solver(), mcts=gen_mcts_solver()), f)
        print("solver golden vectors written")
        return
    kats = gen_bitboard()
    env = gen_env()
    sym = gen_symmetry()
    mcts = gen_mcts()
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(kats, f)
    with open(os.path.join(HERE, "env.json"), "w") as f:
        json.dump(env, f)
    with open(os.path.join(HERE, "symmetry.json"), "w") as f:
        json.dump(sym, f)
    with open(os.path.join(HERE, "mcts.json"), "w") as f:
        json.dump(mcts, f)
    np.savez_compressed(os.path.join(HERE, "ingest.npz"), **gen_ingest())
    with open(os.path.join(HERE, "mcts_features.json"), "w") as f:
        json.dump(gen_mcts_features(), f)
    with open(os.path.join(HERE, "eval_match.json"), "w") as f:
        json.dump(gen_eval_match(), f)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
