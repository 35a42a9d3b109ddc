"""CPU baseline runner: the oracle port of the reference's self-play worker (oracle/mcts.py + oracle/nn.py,
i.e. ReversiPlayer / ReversiEnv semantics with a torch-fp32 CPU forward) timed on the host cores.
Used ONLY by bench.py's cpu_baseline leg and `--impl reference` arm (the Python reference itself cannot
travel to the GPU box and Keras/TensorFlow are not installable)."""
import os
import time

import numpy as np


def usable_cores():
    """cores this process may actually use: affinity mask, capped by the cgroup CPU quota if any"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


def _worker(args):
    (model_kw, weight_seed, play_kw, seed, game_id, budget_s, torch_threads) = args
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
    api.predict(np.zeros((8, 2, 8, 8), np.uint8))  # warm-up (thread pools, first-touch)
    api.rows = api.calls = 0
    t0 = time.perf_counter()
    game.deadline = t0 + budget_s
    e = game.env
    try:
        while not e.done:
            own, enemy = e.own_enemy()
            a = game.decide(own, enemy, e.next_player)
            e.step(a)
    except mcts.TimeUp:
        pass
    dt = time.perf_counter() - t0
    return dict(seconds=dt, plies=len(game.plies), expansions=api.rows, nn_calls=api.calls, sims=game.n_sims)


class _WindowedAPI:
    """predict() of the wrapped evaluator + a snapshot of its counters whenever a window boundary has passed (the game
    loop is not touched: the reference worker has no notion of a bench step)."""

    def __init__(self, api, t0, window_s, windows):
        self.api, self.t0, self.window_s, self.windows = api, t0, window_s, windows
        self.marks = [(t0, 0, 0)]     # (time, rows, calls) at the first predict() after each boundary

    def predict(self, x):
        out = self.api.predict(x)
        now = time.perf_counter()
        if len(self.marks) <= self.windows and now >= self.t0 + len(self.marks) * self.window_s:
            self.marks.append((now, self.api.rows, self.api.calls))
        return out


def _worker_windows(args):
    """one process of `measure_windows`: games played back to back (the first from the opening) for windows x window_s
    seconds; per window: seconds, expansions (network rows), network calls"""
    (model_kw, weight_seed, play_kw, seed, game_id, stride, window_s, windows, torch_threads) = args
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "reversi-alpha-zero_b200"))
    import torch
    torch.set_num_threads(torch_threads)
    from oracle import mcts, nn as onn
    from reversi_zero_b200.agent import model as M
    mc = M.ModelConfig(**model_kw)
    inner = onn.OracleNetAPI(M.build_random_weights(mc, weight_seed), mc.res_layer_num)
    inner.predict(np.zeros((8, 2, 8, 8), np.uint8))  # warm-up (thread pools, first-touch)
    inner.rows = inner.calls = 0
    pp = mcts.PlayParams(**play_kw)
    t0 = time.perf_counter()
    api = _WindowedAPI(inner, t0, window_s, windows)
    games = plies = 0
    try:
        while True:
            game = mcts.SelfPlayGame(pp, api, seed=seed, game_id=game_id + games * stride)
            game.deadline = t0 + windows * window_s
            e = game.env
            while not e.done:
                own, enemy = e.own_enemy()
                e.step(game.decide(own, enemy, e.next_player))
                plies += 1
            games += 1
    except mcts.TimeUp:
        pass
    m = api.marks
    while len(m) <= windows:          # the deadline fell between two predict() calls: close the last window here
        m.append((time.perf_counter(), inner.rows, inner.calls))
    return dict(windows=[dict(seconds=m[i + 1][0] - m[i][0], expansions=m[i + 1][1] - m[i][1], nn_calls=m[i + 1][2] - m[i][2])
                         for i in range(windows)], games_finished=games, plies=plies)


def measure_windows(model_kw, play_kw, windows, window_s, processes=None, torch_threads=1, seed=20260922, weight_seed=0):
    """`processes` game streams in parallel like `measure`, started ONCE and sampled in `windows` consecutive windows of
    `window_s` seconds (the bench's warm-up + timed steps): returns one aggregate dict per window."""
    import json
    import subprocess
    import sys
    processes = processes or usable_cores()
    env = dict(os.environ, OMP_NUM_THREADS=str(torch_threads), MKL_NUM_THREADS=str(torch_threads), CUDA_VISIBLE_DEVICES="")
    procs = []
    for i in range(processes):
        job = json.dumps([model_kw, weight_seed, play_kw, seed, i, processes, window_s, windows, torch_threads])
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "windows", job], stdout=subprocess.PIPE, text=True, env=env))
    res = []
    for pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("cpu baseline worker failed")
        res.append(json.loads(out.strip().splitlines()[-1]))
    out = []
    for w in range(windows):
        ws = [r["windows"][w] for r in res]
        exps = sum(x["expansions"] for x in ws)
        out.append(dict(expansions=exps, expansions_per_s=sum(x["expansions"] / x["seconds"] for x in ws if x["seconds"] > 0),
                        mean_batch=exps / max(1, sum(x["nn_calls"] for x in ws)), processes=processes))
    return out, dict(games_finished=sum(r["games_finished"] for r in res), plies=sum(r["plies"] for r in res))


def measure(model_kw, play_kw, budget_s=15.0, processes=None, torch_threads=1, seed=20260922, weight_seed=0):
    """`processes` game streams in parallel (the reference's multi_process_num workers,
    worker/self_play.py:36-41), each playing one game from the start for `budget_s` seconds of wall clock.
    Workers are plain subprocesses (no fork of a CUDA-initialised parent).  Returns aggregate rates."""
    import json
    import subprocess
    import sys
    processes = processes or usable_cores()
    t0 = time.perf_counter()
    env = dict(os.environ, OMP_NUM_THREADS=str(torch_threads), MKL_NUM_THREADS=str(torch_threads), CUDA_VISIBLE_DEVICES="")
    procs = []
    for i in range(processes):
        job = json.dumps([model_kw, weight_seed, play_kw, seed, i, budget_s, torch_threads])
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), job], stdout=subprocess.PIPE, text=True, env=env))
    res = []
    for pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("cpu baseline worker failed")
        res.append(json.loads(out.strip().splitlines()[-1]))
    exps = sum(r["expansions"] for r in res)
    sims = sum(r["sims"] for r in res)
    return dict(budget_s=budget_s, total_wall_s=time.perf_counter() - t0, expansions=exps, simulations=sims,
                expansions_per_s=sum(r["expansions"] / r["seconds"] for r in res),
                simulations_per_s=sum(r["sims"] / r["seconds"] for r in res), processes=processes, torch_threads=torch_threads,
                mean_batch=exps / max(1, sum(r["nn_calls"] for r in res)))


if __name__ == "__main__":
    import json
    import sys
    if sys.argv[1] == "windows":
        print(json.dumps(_worker_windows(tuple(json.loads(sys.argv[2])))))
    else:
        print(json.dumps(_worker(tuple(json.loads(sys.argv[1])))))
