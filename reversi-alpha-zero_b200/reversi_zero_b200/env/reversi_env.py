"""Mirror of the reference's ``reversi_zero.env.reversi_env`` (env/reversi_env.py): same classes,
attributes and semantics, state transitions computed by the engine's rule code (csrc/rz_bitboard.cuh,
host twin) instead of Python big-int arithmetic."""
import enum
from logging import getLogger

from .. import _cabi
from ..lib.bitboard import board_to_string, bit_count

logger = getLogger(__name__)
Player = enum.Enum("Player", "black white")          # env/reversi_env.py:9
Winner = enum.Enum("Winner", "black white draw")     # env/reversi_env.py:11


def another_player(player):
    return Player.white if player == Player.black else Player.black


class Board:
    """env/reversi_env.py:133-143"""

    def __init__(self, black=None, white=None, init_type=0):
        self.black = black or (0b00010000 << 24 | 0b00001000 << 32)
        self.white = white or (0b00001000 << 24 | 0b00010000 << 32)
        if init_type:
            self.black, self.white = self.white, self.black

    @property
    def number_of_black_and_white(self):
        return bit_count(self.black), bit_count(self.white)


class ReversiEnv:
    """env/reversi_env.py:18-130"""

    def __init__(self):
        self.board = None
        self.next_player = None
        self.turn = 0
        self.done = False
        self.winner = None

    def reset(self):
        self.board = Board()
        self.next_player = Player.black
        self.turn = 0
        self.done = False
        self.winner = None
        return self

    def update(self, black, white, next_player):
        self.board = Board(black, white)
        self.next_player = next_player
        self.turn = sum(self.board.number_of_black_and_white) - 4
        self.done = False
        self.winner = None
        return self

    def step(self, action):
        """action: 0..63, or None to resign (env/reversi_env.py:42-74)."""
        assert action is None or 0 <= action <= 63, f"Illegal action={action}"
        s = _cabi.EnvState(int(self.board.black), int(self.board.white), self.next_player.value, self.turn,
                           int(self.done), 0 if self.winner is None else self.winner.value)
        _cabi.lib().rz_env_step_host(s, -1 if action is None else int(action))
        if s.done and not self.done and action is not None and s.turn == self.turn:
            logger.warning(f"Illegal action={action}, No Flipped!")
        self.board.black, self.board.white = int(s.black), int(s.white)
        self.next_player = Player(s.next_player)
        self.turn = int(s.turn)
        self.done = bool(s.done)
        self.winner = Winner(s.winner) if s.winner else None
        return self.board, {}

    def get_own_and_enemy(self):
        if self.next_player == Player.black:
            return self.board.black, self.board.white
        return self.board.white, self.board.black

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
