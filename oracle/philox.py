"""Philox4x32-10 counter-based RNG -- the engine's RNG, restated in Python so that oracle and CUDA
engine draw identical streams (tests compare them exactly).  Not a reference algorithm: the reference
uses numpy's MT19937 (agent/player.py:112,300-301; lib/bitboard.py:164), which a per-thread device
RNG cannot reproduce; parity with the reference on RNG-driven choices is statistical."""

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF

# stream purposes (counter word 2)
P_DIHEDRAL, P_MOVE, P_NOISE, P_GAME = 0, 1, 2, 3


def philox4x32(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def draw(seed, game_id, seq, purpose, idx=0):
    return philox4x32((game_id & MASK, seq & MASK, purpose & MASK, idx & MASK), (seed & MASK, (seed >> 32) & MASK))


def u01(x):
    """u32 -> double in (0,1)."""
    return (x + 0.5) * (1.0 / 4294967296.0)


def u53(a, b):
    """two u32 -> double in [0,1) with 53 random bits."""
    return ((a >> 5) * 67108864.0 + (b >> 6)) * (1.0 / 9007199254740992.0)
