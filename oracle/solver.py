"""CPU restatement of the reference's endgame solver (lib/alt/reversi_solver_cython.pyx:32-127, the variant
agent/player.py:15 imports).  TEST INFRASTRUCTURE ONLY.

Plain minimax without pruning over moves in ascending square order, first strictly-better move wins ties
(:92-95); the score is the final disc difference from the point of view of the side to move (:122); when the
opponent has no move the same side moves again without a sign flip (:90-91,117-119); `exactly=False`
(win / loss / draw mode) stops a node as soon as it has a move with a positive score (:99).  The reference keeps
one cache dict per player and clears it when an exact call follows a WLD call (:47-49); within one game a
player's WLD calls (turn < use_solver_turn) all precede its exact calls, so the two modes never see each other's
entries -- here each mode simply has its own dict.  No timeout (the reference gives up after 30 s, :78-79).
"""
from . import bitboard as bb


class Solver:
    def __init__(self):
        self.cache = {True: {}, False: {}}
        self.nodes = 0

    MAX_EMPTIES = 12  # the device solver refuses larger positions (csrc/rz_solver.cuh), the analogue of the reference's timeout

    def solve(self, own, enemy, exactly=False):
        """-> (move, score) for the side to move (own), or (None, None) if it has no move / the position is refused."""
        if 64 - bb.bit_count(own | enemy) > self.MAX_EMPTIES:
            return None, None
        move, score = self._f(own, enemy, exactly)
        return (None, None) if move < 0 else (move, score)

    def _f(self, own, enemy, exactly):
        cache = self.cache[exactly]
        key = (own, enemy)
        hit = cache.get(key)
        if hit is not None:
            return hit
        self.nodes += 1
        legal = bb.find_correct_moves(own, enemy)
        best_move, best_score = -1, -100
        m = legal
        while m:
            if not exactly and best_score > 0:
                break
            a = (m & -m).bit_length() - 1
            m &= m - 1
            fl = bb.calc_flip(a, own, enemy)
            own2, en2 = (own ^ fl) | (1 << a), enemy ^ fl
            if bb.find_correct_moves(en2, own2):
                score = -self._f(en2, own2, exactly)[1]
            elif bb.find_correct_moves(own2, en2):
                score = self._f(own2, en2, exactly)[1]
            else:
                score = bb.bit_count(own2) - bb.bit_count(en2)
            if best_score < score:
                best_move, best_score = a, score
        cache[key] = (best_move, best_score)
        return best_move, best_score
