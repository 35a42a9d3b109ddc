"""Latency of the device endgame solver (rz_solve_dev): one batch = one request per lane of its grid (2 CTAs x 128 lanes
per SM), every lane running its own resumable stack machine to completion."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "reversi-alpha-zero_b200"))


def positions(empties, n, seed):
    from oracle import bitboard as ob
    rng = np.random.default_rng(seed)
    own, enemy = [], []
    while len(own) < n:
        e = ob.Env().reset()
        while not e.done and 60 - e.turn > empties:
            o, en = e.own_enemy(); legal = ob.find_correct_moves(o, en)
            ms = [i for i in range(64) if legal >> i & 1]; e.step(ms[rng.integers(len(ms))])
        if not e.done:
            o, en = e.own_enemy(); own.append(o); enemy.append(en)
    return np.array(own, np.uint64), np.array(enemy, np.uint64)


def main():
    import torch
    from reversi_zero_b200 import _cabi, device as D
    lib = _cabi.lib()
    n = 148 * 2 * 128  # lanes of the kernel's grid
    for empties in (8, 10):
        own, enemy = positions(empties, 400, empties)
        own = np.resize(own, n); enemy = np.resize(enemy, n)
        d_own, d_en = D.to_device(own), D.to_device(enemy)
        d_mv, d_sc = D.empty(n, np.int8), D.empty(n, np.int8)
        for exact in (0, 1):
            d_ex = D.to_device(np.full(n, exact, np.uint8))
            s = torch.cuda.current_stream()
            times = []
            for rep in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                _cabi.check(lib.rz_solve_dev(D.ptr(d_own), D.ptr(d_en), D.ptr(d_ex), D.ptr(d_mv), D.ptr(d_sc), n, D.stream_ptr(s)), "solve")
                e1.record(s); torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            print(json.dumps(dict(empties=empties, exact=exact, batch=n, ms_first=times[0], ms_warm=min(times[1:]))), flush=True)


if __name__ == "__main__":
    main()
