"""GPU parity tests for K1 (batched bitboard kernels) through the C ABI: bit-exact against the golden
vectors from the reference and against the C oracle on seeded random inputs (SURVEY 8(d) config 5
recipe), plus size-independent properties at the full 10M-position size."""
import json
import os

import numpy as np
import pytest

from oracle import bitboard as ob
from reversi_zero_b200.lib import bitboard as zb

pytestmark = pytest.mark.gpu
U64 = np.uint64


def positions(seed, n):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2 ** 64, size=n, dtype=U64)
    b = rng.integers(0, 2 ** 64, size=n, dtype=U64)
    r = rng.integers(0, 2 ** 64, size=n, dtype=U64)
    third = n // 3
    occ = a.copy()
    occ[:third] = a[:third] & b[:third]
    occ[2 * third:] = a[2 * third:] | b[2 * third:]
    return occ & r, occ & ~r, rng.integers(0, 64, size=n, dtype=np.uint8)


def test_golden_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "bitboard.npz"))
    assert np.array_equal(zb.find_correct_moves_batch(g["own"], g["enemy"]), g["legal"])
    assert np.array_equal(zb.calc_flip_batch(g["pos"], g["own"], g["enemy"]), g["flip"])
    for row, i in zip(g["flip_all"], g["sub"]):
        got = zb.calc_flip_batch(np.arange(64, dtype=np.uint8), np.full(64, g["own"][i]), np.full(64, g["enemy"][i]))
        assert np.array_equal(got, row)


def test_dihedral_device_vs_golden(golden_dir):
    """rz_dihedral_dev (the DEVICE code: __byte_perm / __brevll paths of rz_bitboard.cuh, used by the engine's leaf gather
    and by the ingest kernel) against the reference's transforms, lib/bitboard.py:119-159, on the golden positions: all 8
    values of t, t = flip * 4 + rot meaning flip_vertical first, then rot x rotate90 (agent/player.py:166-179,300-305)."""
    g = np.load(os.path.join(golden_dir, "bitboard.npz"))
    for name in ("own", "enemy"):
        x = g[name]
        got = {t: zb.dihedral_batch(x, t) for t in range(8)}
        assert np.array_equal(got[0], x)
        if name == "own":   # the golden transforms were generated from `own` by the unmodified reference
            assert np.array_equal(got[1], g["rotate90"]) and np.array_equal(got[2], g["rotate180"])
            assert np.array_equal(got[4], g["flip_vertical"])
            # flip_diag_a1h8 (lib/bitboard.py:141-151) is not a member the players use directly; rotate90 is defined as
            # flip_diag(flip_vertical(x)) (:154), so flip_diag(x) = rotate90(flip_vertical(x)) = t 5
            assert np.array_equal(got[5], g["flip_diag"])
        # composition, element by element, through the oracle functions (pinned to the same goldens in tests/test_oracle.py)
        for t in range(8):
            want = np.array([_compose(int(v), t) for v in x[:512]], dtype=U64)
            assert np.array_equal(got[t][:512], want), t
        # group structure on the full set: four rotations are the identity, the flip is an involution, rot^-1 = rot^3
        assert np.array_equal(zb.dihedral_batch(got[3], 1), x) and np.array_equal(zb.dihedral_batch(got[4], 4), x)
        assert np.array_equal(zb.dihedral_batch(got[6], 6), x)       # flip then rotate180 is an involution as well
    # per-element t, ragged size, and agreement with the host twin the single-environment mirrors use
    rng = np.random.default_rng(3)
    x = g["own"][:1001]
    t = rng.integers(0, 8, size=x.size, dtype=np.uint8)
    assert np.array_equal(zb.dihedral_batch(x, t), np.array([zb.dihedral(int(v), int(tt)) for v, tt in zip(x, t)], dtype=U64))
    assert zb.dihedral_batch(x[:0], 0).size == 0


def _compose(v, t):
    if t & 4:
        v = ob.flip_vertical(v)
    for _ in range(t & 3):
        v = ob.rotate90(v)
    return v


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 256, 257, 100003])
def test_ragged_sizes_vs_oracle(n):
    own, enemy, pos = positions(n + 5, n)
    assert np.array_equal(zb.find_correct_moves_batch(own, enemy), ob.find_correct_moves_batch(own, enemy))
    assert np.array_equal(zb.calc_flip_batch(pos, own, enemy), ob.calc_flip_batch(pos, own, enemy))


def test_unaligned_views_vs_oracle():
    own, enemy, pos = positions(11, 4099)
    for off in (1, 3):  # odd element offsets break the 16-byte alignment the vector path wants
        o, e, p = own[off:], enemy[off:], pos[off:]
        assert np.array_equal(zb.find_correct_moves_batch(o, e), ob.find_correct_moves_batch(o, e))
        assert np.array_equal(zb.calc_flip_batch(p, o, e), ob.calc_flip_batch(p, o, e))


def test_overlapping_inputs_follow_the_reference_arithmetic():
    """own and enemy sharing squares is not a board, but lib/bitboard.py gives such input a definite answer (pure bit
    arithmetic; its env even produces such boards when a move onto an occupied square flips something, reversi_env.py:56-63).
    The bit-sliced kernels walk rays instead of rippling a carry, so they hand groups of 32 positions that contain an overlap
    to the scalar code: the results must still be the oracle's, bit for bit."""
    rng = np.random.default_rng(21)
    n = 50_000
    own = rng.integers(0, 2 ** 64, size=n, dtype=U64) & rng.integers(0, 2 ** 64, size=n, dtype=U64)
    enemy = rng.integers(0, 2 ** 64, size=n, dtype=U64) & rng.integers(0, 2 ** 64, size=n, dtype=U64)
    clean = rng.random(n) < 0.9                       # 90 % proper boards, the rest with shared squares: most groups of 32 are mixed
    enemy[clean] &= ~own[clean]
    pos = rng.integers(0, 64, size=n, dtype=np.uint8)
    assert (own & enemy).any()
    assert np.array_equal(zb.find_correct_moves_batch(own, enemy), ob.find_correct_moves_batch(own, enemy))
    assert np.array_equal(zb.calc_flip_batch(pos, own, enemy), ob.calc_flip_batch(pos, own, enemy))


def test_full_size_bit_exact_and_properties():
    n = 10_000_000  # BASELINE.json config 5
    own, enemy, pos = positions(20260922, n)
    legal = zb.find_correct_moves_batch(own, enemy)
    flip = zb.calc_flip_batch(pos, own, enemy)
    assert np.array_equal(legal, ob.find_correct_moves_batch(own, enemy))
    assert np.array_equal(flip, ob.calc_flip_batch(pos, own, enemy))
    # properties: moves only on empty squares; flips only opponent discs; on an EMPTY square the move is
    # legal iff something flips
    assert not np.any(legal & (own | enemy))
    assert not np.any(flip & ~enemy)
    bit = U64(1) << pos.astype(U64)
    empty = (bit & (own | enemy)) == 0
    assert np.array_equal(((legal & bit) != 0)[empty], (flip != 0)[empty])


def test_step_vs_oracle_random_playouts(golden_dir):
    rng = np.random.default_rng(5)
    n = 20000
    st = dict(black=np.full(n, 0x0000000810000000, U64), white=np.full(n, 0x0000001008000000, U64),
              next_player=np.ones(n, np.uint8), turn=np.zeros(n, np.uint8), done=np.zeros(n, np.uint8),
              winner=np.zeros(n, np.uint8))
    ref = {k: v.copy() for k, v in st.items()}
    for ply in range(64):
        own = np.where(st["next_player"] == 1, st["black"], st["white"])
        enemy = np.where(st["next_player"] == 1, st["white"], st["black"])
        legal = ob.find_correct_moves_batch(own, enemy)
        # random legal move (sometimes an illegal / occupied square, sometimes a resignation)
        action = np.empty(n, np.int8)
        r = rng.integers(0, 64, size=n)
        for i in range(n):
            m = int(legal[i])
            if m == 0 or rng.random() < 0.002:
                action[i] = r[i] if rng.random() < 0.5 else -1
            else:
                k = r[i] % m.bit_count()
                for _ in range(k):
                    m &= m - 1
                action[i] = (m & -m).bit_length() - 1
        live = st["done"] == 0
        got_legal = zb.step_batch(st["black"], st["white"], st["next_player"], st["turn"], st["done"], st["winner"],
                                  action, want_legal=True)
        ob.step_batch(ref["black"], ref["white"], ref["next_player"], ref["turn"], ref["done"], ref["winner"], action)
        for k in st:
            assert np.array_equal(st[k][live], ref[k][live]), (ply, k)
        own2 = np.where(ref["next_player"] == 1, ref["black"], ref["white"])
        en2 = np.where(ref["next_player"] == 1, ref["white"], ref["black"])
        exp_legal = np.where(ref["done"] == 1, U64(0), ob.find_correct_moves_batch(own2, en2))
        assert np.array_equal(got_legal[live], exp_legal[live])
        # finished games are frozen for the rest of the test
        for k in st:
            st[k][~live] = ref[k][~live]
        if not live.any():
            break
    # golden env playouts through the batched kernel, one game per lane
    g = json.load(open(os.path.join(golden_dir, "env.json")))
    games = g["games"]
    m = len(games)
    st = dict(black=np.full(m, 0x0000000810000000, U64), white=np.full(m, 0x0000001008000000, U64),
              next_player=np.ones(m, np.uint8), turn=np.zeros(m, np.uint8), done=np.zeros(m, np.uint8),
              winner=np.zeros(m, np.uint8))
    for ply in range(max(len(x["actions"]) for x in games)):
        idx = [i for i, x in enumerate(games) if ply < len(x["actions"])]
        sub = {k: np.ascontiguousarray(v[idx]) for k, v in st.items()}
        act = np.array([games[i]["actions"][ply] for i in idx], np.int8)
        zb.step_batch(sub["black"], sub["white"], sub["next_player"], sub["turn"], sub["done"], sub["winner"], act)
        for j, i in enumerate(idx):
            assert [int(sub[k][j]) for k in ("black", "white", "next_player", "turn", "done", "winner")] == games[i]["states"][ply + 1]
        for k in st:
            st[k][idx] = sub[k]
