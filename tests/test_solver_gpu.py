"""GPU parity of the batched endgame solver (rz_solve) through the C ABI: the reference's KATs and golden
positions (exact + WLD, tests/golden/solver.json) and seeded random endgames against the oracle restatement."""
import json
import os

import numpy as np
import pytest

from oracle import bitboard as ob
from oracle.solver import Solver
from reversi_zero_b200.lib import reversi_solver as zs
from reversi_zero_b200.env.reversi_env import Player

pytestmark = pytest.mark.gpu


def test_golden_positions(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "solver.json")))["positions"]
    own = [c["black"] if c["next_player"] == 1 else c["white"] for c in g]
    enemy = [c["white"] if c["next_player"] == 1 else c["black"] for c in g]
    mv, sc = zs.solve_batch(own, enemy, [c["exactly"] for c in g])
    for c, m, s in zip(g, mv, sc):
        assert (int(m), int(s)) == (c["move"], c["score"]), c["tag"]
    # reference-style object API, KAT q1 (lib/reversi_solver.py:102-118): white to move, WLD -> (57, +2)
    assert zs.ReversiSolver().solve(0x80feafd2eaf20200, 0x7c00502d150d0d0f, Player.white, exactly=False) == (57, 2)


def test_random_endgames_vs_oracle():
    rng = np.random.default_rng(17)
    own, enemy = [], []
    while len(own) < 300:
        empties = int(rng.integers(1, 11))
        e = ob.Env().reset()
        while not e.done and 60 - e.turn > empties:
            o, en = e.own_enemy()
            legal = ob.find_correct_moves(o, en)
            ms = [i for i in range(64) if legal >> i & 1]
            e.step(ms[rng.integers(len(ms))])
        if not e.done:
            o, en = e.own_enemy()
            own.append(o); enemy.append(en)
    for exactly in (True, False):
        mv, sc = zs.solve_batch(own, enemy, [exactly] * len(own))
        for o, en, m, s in zip(own, enemy, mv, sc):
            assert (int(m), int(s)) == Solver().solve(o, en, exactly)


def test_refused_and_empty_cases():
    # opening position: 60 empties -> refused like a timeout; a position without a legal move -> no move
    mv, sc = zs.solve_batch([0x0000000810000000, 0xFFFFFFFFFFFFFF00], [0x0000001008000000, 0x00000000000000FE], [True, True])
    assert list(mv) == [-1, -1] and list(sc) == [0, 0]
    assert zs.ReversiSolver().solve(0x0000000810000000, 0x0000001008000000, Player.black, exactly=True) == (None, None)
