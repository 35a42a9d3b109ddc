"""CPU baseline runner: the oracle port of the reference's self-play worker (oracle/mcts.py + oracle/nn.py,
i.e. ReversiPlayer / ReversiEnv semantics with a torch-fp32 CPU forward) timed on the host cores.
Used ONLY by bench.py's cpu_baseline leg and `--impl reference` arm (the Python reference itself cannot
travel to the GPU box and Keras/TensorFlow are not installable)."""
import os
import time

import numpy as np


def _worker(args):
    (model_kw, weight_seed, play_kw, seed, game_id, n_search_plies, torch_threads) = args
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "reversi-alpha-zero_b200"))
    import torch
    torch.set_num_threads(torch_threads)
    from oracle import mcts, nn as onn
    from reversi_zero_b200.agent import model as M
    mc = M.ModelConfig(**model_kw)
    api = onn.OracleNetAPI(M.build_random_weights(mc, weight_seed), mc.res_layer_num)
    pp = mcts.PlayParams(**play_kw)
    game = mcts.SelfPlayGame(pp, api, seed=seed, game_id=game_id)
    t0 = time.perf_counter()
    searched = 0
    e = game.env
    while not e.done and searched < n_search_plies:
        own, enemy = e.own_enemy()
        turn = e.turn
        a = game.decide(own, enemy, e.next_player)
        if turn > 0:
            searched += 1
        e.step(a)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, searched_plies=searched, expansions=api.rows, nn_calls=api.calls, sims=game.n_sims)


def measure(model_kw, play_kw, n_search_plies=2, processes=None, torch_threads=1, seed=20260922, weight_seed=0):
    """Plays the first `n_search_plies` searched plies of one game per process, `processes` games in parallel
    (the reference's multi_process_num workers, worker/self_play.py:36-41).  Workers are plain subprocesses
    (no fork of a CUDA-initialised parent, no multiprocessing start-method pitfalls).  Returns aggregate rates."""
    import json
    import subprocess
    import sys
    processes = processes or os.cpu_count() or 1
    t0 = time.perf_counter()
    env = dict(os.environ, OMP_NUM_THREADS=str(torch_threads), MKL_NUM_THREADS=str(torch_threads), CUDA_VISIBLE_DEVICES="")
    procs = []
    for i in range(processes):
        job = json.dumps([model_kw, weight_seed, play_kw, seed, i, n_search_plies, torch_threads])
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), job], stdout=subprocess.PIPE, text=True, env=env))
    res = []
    for pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("cpu baseline worker failed")
        res.append(json.loads(out.strip().splitlines()[-1]))
    wall = max(r["seconds"] for r in res)
    plies = sum(r["searched_plies"] for r in res)
    exps = sum(r["expansions"] for r in res)
    return dict(wall_s=wall, total_wall_s=time.perf_counter() - t0, searched_plies=plies, expansions=exps,
                plies_per_s=plies / wall, expansions_per_s=exps / wall, processes=processes, torch_threads=torch_threads,
                mean_batch=exps / max(1, sum(r["nn_calls"] for r in res)))


if __name__ == "__main__":
    import json
    import sys
    print(json.dumps(_worker(tuple(json.loads(sys.argv[1])))))
