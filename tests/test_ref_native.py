"""The reference's OWN compiled code (lib/alt/bitboard_cython.pyx, lib/alt/reversi_solver_cython.pyx, built by
oracle/build_ref.py into oracle/_ref -- it travels to the GPU box) against
  * the oracle restatements (CPU, here), and
  * the CUDA operators through the C ABI (GPU): K1 move generation / flips on 1 M seeded positions of the
    SURVEY 8(d) config-5 recipe, and the batched endgame solver.
Skipped where oracle/_ref has not been built (it needs the reference checkout once)."""
import numpy as np
import pytest

from oracle import bitboard as ob, ref_native
from oracle.solver import Solver

pytestmark = pytest.mark.skipif(not ref_native.available(), reason="oracle/_ref not built")

U64 = np.uint64


def config5_positions(n, seed=20260922):
    """thirds of the set with occupancy a & b / a / a | b (densities 1/4, 1/2, 3/4), disjoint own/enemy."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2 ** 64, size=n, dtype=U64)
    b = rng.integers(0, 2 ** 64, size=n, dtype=U64)
    r = rng.integers(0, 2 ** 64, size=n, dtype=U64)
    occ = a.copy()
    third = n // 3
    occ[:third] = a[:third] & b[:third]
    occ[2 * third:] = a[2 * third:] | b[2 * third:]
    pos = rng.integers(0, 64, size=n, dtype=np.uint8)
    return occ & r, occ & ~r, pos


def reference_outputs(own, enemy, pos):
    bb, _ = ref_native.load()
    legal = np.fromiter((bb.find_correct_moves(int(o), int(e)) for o, e in zip(own, enemy)), dtype=U64, count=len(own))
    flip = np.fromiter((bb.calc_flip(int(p), int(o), int(e)) for p, o, e in zip(pos, own, enemy)), dtype=U64, count=len(own))
    return legal, flip


def random_endgames(n, seed, max_empties=10):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        empties = int(rng.integers(1, max_empties + 1))
        e = ob.Env().reset()
        while not e.done and 60 - e.turn > empties:
            o, en = e.own_enemy()
            legal = ob.find_correct_moves(o, en)
            ms = [i for i in range(64) if legal >> i & 1]
            e.step(ms[rng.integers(len(ms))])
        if not e.done:
            out.append(e.own_enemy())
    return out


def reference_solve(own, enemy, exactly):
    """ReversiSolver.solve(black, white, next_player, exactly) of the compiled reference, position given in the mover's frame"""
    _, sv = ref_native.load()
    mv, sc = sv.ReversiSolver().solve(own, enemy, ref_native.player_enum().black, exactly=exactly)
    return (-1, 0) if mv is None else (int(mv), int(sc))


def test_oracle_bitboard_matches_compiled_reference():
    own, enemy, pos = config5_positions(100_000)
    legal, flip = reference_outputs(own, enemy, pos)
    assert np.array_equal(ob.find_correct_moves_batch(own, enemy), legal)
    assert np.array_equal(ob.calc_flip_batch(pos, own, enemy), flip)


def test_oracle_solver_matches_compiled_reference():
    for own, enemy in random_endgames(120, seed=41):
        for exactly in (True, False):
            mv, sc = Solver().solve(own, enemy, exactly)
            assert ((-1, 0) if mv is None else (mv, sc)) == reference_solve(own, enemy, exactly)


@pytest.mark.gpu
def test_k1_kernels_match_compiled_reference_1m():
    from reversi_zero_b200.lib import bitboard as zb
    own, enemy, pos = config5_positions(1_000_000)
    legal, flip = reference_outputs(own, enemy, pos)
    assert np.array_equal(zb.find_correct_moves_batch(own, enemy), legal)      # bit-exact, 1 M positions
    assert np.array_equal(zb.calc_flip_batch(pos, own, enemy), flip)          # including occupied / illegal squares


@pytest.mark.gpu
def test_device_solver_matches_compiled_reference():
    from reversi_zero_b200.lib import reversi_solver as zs
    cases = random_endgames(200, seed=43)
    own = [c[0] for c in cases]
    enemy = [c[1] for c in cases]
    for exactly in (True, False):
        mv, sc = zs.solve_batch(own, enemy, [exactly] * len(cases))
        for o, e, m, s in zip(own, enemy, mv, sc):
            assert (int(m), int(s)) == reference_solve(o, e, exactly)
