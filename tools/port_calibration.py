"""Calibration of the CPU baseline: the oracle port of the reference worker (oracle/selfplay_cpu.py, the
`kind: "port"` arm of bench.py -- the only thing that can run on the GPU box) against the UNMODIFIED reference worker loop
(worker/self_play.py:139-175 driving agent/player.py ReversiPlayer through oracle/ref_shims) on the same cores of the build
container, same network evaluator (torch fp32 CPU forward of the random-init ch5 net, 1 thread per process), same play
parameters (BASELINE config 2: 400 simulations, K = 8, c_puct 5, noise 0.25, tau turn 4, thinking_loop 1, solver and
resignation off).  Each side: `procs` processes x `budget` seconds of one game from the opening.

    python tools/port_calibration.py [budget_s=60] [procs=nproc]   ->  profiles/port_calibration_r02.json
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "reversi-alpha-zero_b200"))

MODEL_KW = dict(cnn_filter_num=256, cnn_filter_size=3, res_layer_num=10, value_fc_size=256)
PLAY_KW = dict(simulation_num_per_move=400, parallel_search_num=8, c_puct=5, virtual_loss=3, noise_eps=0.25, dirichlet_alpha=0.5,
               change_tau_turn=4, thinking_loop=1, resign_threshold=None, share_mtcs_info_in_self_play=True)


def reference_worker(budget_s, seed):
    """one game of the unmodified reference loop for budget_s seconds"""
    import numpy as np
    import torch
    torch.set_num_threads(1)
    import oracle.ref_shims.install as shims
    shims.install()
    from reversi_zero.config import Config
    from reversi_zero.env.reversi_env import ReversiEnv, Player
    from reversi_zero.agent.player import ReversiPlayer
    from oracle import nn as onn
    from reversi_zero_b200.agent import model as M
    mc = M.ModelConfig(**MODEL_KW)
    api = onn.OracleNetAPI(M.build_random_weights(mc, 0), mc.res_layer_num)
    cfg = Config()
    for k, v in PLAY_KW.items():
        setattr(cfg.play, k, v)
    cfg.play.use_solver_turn = cfg.play.use_solver_turn_in_simulation = 0
    np.random.seed(seed)
    api.predict(np.zeros((8, 2, 8, 8), np.uint8))
    api.rows = api.calls = 0
    env = ReversiEnv().reset()
    info = ReversiPlayer.create_mtcs_info()
    black = ReversiPlayer(cfg, None, enable_resign=False, mtcs_info=info, api=api)
    white = ReversiPlayer(cfg, None, enable_resign=False, mtcs_info=info, api=api)
    t0 = time.perf_counter()
    plies = 0
    while not env.done and time.perf_counter() - t0 < budget_s:   # the budget is checked between moves: whole searches only
        if env.next_player == Player.black:
            a = black.action_with_evaluation(env.board.black, env.board.white)
        else:
            a = white.action_with_evaluation(env.board.white, env.board.black)
        env.step(a.action)
        plies += 1
    return dict(seconds=time.perf_counter() - t0, plies=plies, expansions=api.rows, nn_calls=api.calls)


def port_worker(budget_s, seed):
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import mcts, nn as onn
    from reversi_zero_b200.agent import model as M
    mc = M.ModelConfig(**MODEL_KW)
    api = onn.OracleNetAPI(M.build_random_weights(mc, 0), mc.res_layer_num)
    game = mcts.SelfPlayGame(mcts.PlayParams(**PLAY_KW), api, seed=20260922, game_id=seed)
    api.predict(np.zeros((8, 2, 8, 8), np.uint8))
    api.rows = api.calls = 0
    e = game.env
    t0 = time.perf_counter()
    while not e.done and time.perf_counter() - t0 < budget_s:     # same rule: whole searches only
        own, enemy = e.own_enemy()
        e.step(game.decide(own, enemy, e.next_player))
    return dict(seconds=time.perf_counter() - t0, plies=len(game.plies), expansions=api.rows, nn_calls=api.calls)


def run_side(which, budget_s, procs):
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", which, str(budget_s), str(i)], stdout=subprocess.PIPE,
                           text=True, env=env) for i in range(procs)]
    res = []
    for p in ps:
        out, _ = p.communicate()
        assert p.returncode == 0, which
        res.append(json.loads(out.strip().splitlines()[-1]))
    return dict(processes=procs, expansions=sum(r["expansions"] for r in res), plies=sum(r["plies"] for r in res),
                expansions_per_s=sum(r["expansions"] / r["seconds"] for r in res), plies_per_s=sum(r["plies"] / r["seconds"] for r in res),
                mean_nn_batch=sum(r["expansions"] for r in res) / max(1, sum(r["nn_calls"] for r in res)),
                seconds_per_process=[round(r["seconds"], 1) for r in res])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        fn = reference_worker if sys.argv[2] == "reference" else port_worker
        print(json.dumps(fn(float(sys.argv[3]), int(sys.argv[4]))))
        sys.exit(0)
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    ref = run_side("reference", budget, procs)
    port = run_side("port", budget, procs)
    ratio = port["expansions_per_s"] / ref["expansions_per_s"]
    out = dict(what=__doc__.split("\n\n")[0], budget_s=budget, reference_unmodified=ref, port=port,
               summary=dict(port_over_reference_expansions_per_s=ratio, port_over_reference_plies_per_s=port["plies_per_s"] / ref["plies_per_s"],
                            cores=procs, where="build container (no GPU); the reference cannot run on the GPU box",
                            reading="a ratio above 1 means the port is FASTER than the reference, i.e. the reported GPU/CPU speed-up is conservative"))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "port_calibration_r02.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out["summary"]))
