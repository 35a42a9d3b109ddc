"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's trainer-side ingest,
SURVEY 8(f).4 -- from compact play rows (one per recorded ply) to the arrays the reference trainer builds.

Follows, in order:
  * agent/player.py:132,366-385   stored policy: N / sum(N) (float64) when save_policy_of_tau_1 or turn < change_tau_turn,
                                  else one-hot at np.argmax(N)
  * agent/player.py:166-179       8 symmetries, flip in (False, True) x rot_right in 0..3 (oracle.mcts.symmetries8)
  * agent/player.py:357-364       z appended to every record
  * worker/optimize.py:215-231    convert_to_training_data: bit_to_array(own/enemy, 64).reshape(8, 8), np.array(...)
Pinned against the reference itself by tests/golden/ingest.npz (tests/golden/make_golden.py ingest).
"""
import numpy as np

from . import bitboard as bb
from .mcts import symmetries8


def rows_to_training_arrays(own, enemy, n_visit, row_z, save_policy_of_tau_1, change_tau_turn):
    """-> (states uint8 [8n,2,8,8], policy float64 [8n,64], z int64 [8n]); record 8*r + t = row r, symmetry t."""
    states, policies, zs = [], [], []
    for o, e, n, z in zip(own, enemy, n_visit, row_z):
        o, e = int(o), int(e)
        n = np.asarray(n, dtype=np.float64)
        turn = bin(o).count("1") + bin(e).count("1") - 4
        if save_policy_of_tau_1 or turn < change_tau_turn:
            pol = n / np.sum(n)
        else:
            pol = np.zeros(64)
            pol[int(np.argmax(n))] = 1
        for o_s, e_s, p_s in symmetries8(o, e, pol):
            states.append([bb.bit_to_array(o_s, 64).reshape(8, 8), bb.bit_to_array(e_s, 64).reshape(8, 8)])
            policies.append(p_s)
            zs.append(int(z))
    if not states:
        return np.zeros((0, 2, 8, 8), np.uint8), np.zeros((0, 64)), np.zeros((0,), np.int64)
    return np.array(states, dtype=np.uint8), np.array(policies, dtype=np.float64), np.array(zs, dtype=np.int64)
