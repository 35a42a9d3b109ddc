"""Mirror of the reference's ``ReversiSolver`` (lib/alt/reversi_solver_cython.pyx:32-61, imported by
agent/player.py:15) over the device solver (csrc/rz_solver.cuh): ``solve(black, white, next_player, timeout,
exactly) -> (move, score)`` or ``(None, None)``.  ``solve_batch`` solves many positions in one launch."""
import numpy as np

from .. import _cabi


def solve_batch(own, enemy, exactly):
    """own / enemy: uint64 arrays in the mover's frame; exactly: bool array.  -> (move int8[], score int8[]), move -1 = none."""
    own = np.ascontiguousarray(own, dtype=np.uint64)
    enemy = np.ascontiguousarray(enemy, dtype=np.uint64)
    ex = np.ascontiguousarray(np.broadcast_to(np.asarray(exactly, dtype=np.uint8), own.shape))
    move = np.empty(own.shape, np.int8)
    score = np.empty(own.shape, np.int8)
    _cabi.check(_cabi.lib().rz_solve(own.ctypes.data_as(_cabi.u64p), enemy.ctypes.data_as(_cabi.u64p), ex.ctypes.data_as(_cabi.u8p),
                                      move.ctypes.data_as(_cabi.i8p), score.ctypes.data_as(_cabi.i8p), own.size), "rz_solve")
    return move, score


class ReversiSolver:
    def solve(self, black, white, next_player, timeout=30, exactly=False):
        """next_player: Player enum (or its value: 1 black, 2 white).  `timeout` is accepted for compatibility; the device
        solver refuses positions with more than 12 empty squares instead (returns (None, None) like a timeout)."""
        p = getattr(next_player, "value", next_player)
        own, enemy = (black, white) if p == 1 else (white, black)
        mv, sc = solve_batch([own], [enemy], [exactly])
        if mv[0] < 0:
            return None, None
        return int(mv[0]), int(sc[0])
