"""``ReversiEnv`` / ``Board`` / ``Player`` / ``Winner`` with the reference's interface
(env/reversi_env.py:9-143), implemented as a thin view over the engine's game-state struct
(``rz_env_state``): every transition is computed by the engine's rule code (csrc/rz_bitboard.cuh, host
twin), not by Python big-int arithmetic."""
import enum
from logging import getLogger

from .. import _cabi
from ..lib.bitboard import board_to_string

logger = getLogger(__name__)
Player = enum.Enum("Player", "black white")          # values 1, 2 (env/reversi_env.py:9)
Winner = enum.Enum("Winner", "black white draw")     # values 1, 2, 3 (env/reversi_env.py:11)

_START_BLACK = 0b00010000 << 24 | 0b00001000 << 32
_START_WHITE = 0b00001000 << 24 | 0b00010000 << 32


def another_player(player):
    return Player(3 - player.value)


class Board:
    """Two bitboards; a zero / None bitboard means "start stones" as in the reference (env/reversi_env.py:133-140)."""

    def __init__(self, black=None, white=None, init_type=0):
        b, w = black or _START_BLACK, white or _START_WHITE
        self.black, self.white = (w, b) if init_type else (b, w)

    @property
    def number_of_black_and_white(self):
        return int(self.black).bit_count(), int(self.white).bit_count()


class ReversiEnv:
    def __init__(self):
        self.board = None
        self.next_player = None
        self.turn = 0
        self.done = False
        self.winner = None

    # -- state <-> rz_env_state ----------------------------------------------------------------------------
    def _pack(self):
        return _cabi.EnvState(int(self.board.black), int(self.board.white), self.next_player.value, self.turn, int(self.done),
                              self.winner.value if self.winner else 0)

    def _unpack(self, s):
        self.board.black, self.board.white = int(s.black), int(s.white)
        self.next_player, self.turn = Player(s.next_player), int(s.turn)
        self.done, self.winner = bool(s.done), (Winner(s.winner) if s.winner else None)

    # -- reference interface -------------------------------------------------------------------------------
    def reset(self):
        return self.update(None, None, Player.black)

    def update(self, black, white, next_player):
        """env/reversi_env.py:34-40: turn = stones - 4."""
        self.board = Board(black, white)
        self.next_player = next_player
        self.turn = sum(self.board.number_of_black_and_white) - 4
        self.done, self.winner = False, None
        return self

    def step(self, action):
        """env/reversi_env.py:42-74.  action in 0..63, None = resign; an action that flips nothing loses."""
        assert action is None or 0 <= action <= 63, f"Illegal action={action}"
        s = self._pack()
        turn_before = s.turn
        _cabi.lib().rz_env_step_host(s, -1 if action is None else int(action))
        if action is not None and s.done and s.turn == turn_before:
            logger.warning(f"Illegal action={action}, No Flipped!")
        self._unpack(s)
        return self.board, {}

    def get_own_and_enemy(self):
        b = self.board
        return (b.black, b.white) if self.next_player == Player.black else (b.white, b.black)

    def set_own_and_enemy(self, own, enemy):
        if self.next_player == Player.black:
            self.board.black, self.board.white = own, enemy
        else:
            self.board.white, self.board.black = own, enemy

    def render(self):
        b, w = self.board.number_of_black_and_white
        print(f"next={self.next_player.name} turn={self.turn} B={b} W={w}")
        print(board_to_string(self.board.black, self.board.white, with_edge=True))

    @property
    def observation(self):
        return self.board
