"""ctypes front-end of the C oracle (oracle/rz_oracle.c).  Test infrastructure only."""
import ctypes as C
import numpy as np

from .build import build

_lib = C.CDLL(build())
_u64, _u64p, _u8p, _i8p = C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_int8)
_lib.rzo_find_correct_moves.restype = _u64
_lib.rzo_find_correct_moves.argtypes = [_u64, _u64]
_lib.rzo_calc_flip.restype = _u64
_lib.rzo_calc_flip.argtypes = [C.c_int, _u64, _u64]
for _n in ("rzo_flip_vertical", "rzo_flip_diag_a1h8", "rzo_rotate90", "rzo_rotate180"):
    getattr(_lib, _n).restype = _u64
    getattr(_lib, _n).argtypes = [_u64]
_lib.rzo_dihedral.restype = _u64
_lib.rzo_dihedral.argtypes = [_u64, C.c_int]
_lib.rzo_bit_count.restype = C.c_int
_lib.rzo_bit_count.argtypes = [_u64]


class EnvStruct(C.Structure):
    _fields_ = [("black", _u64), ("white", _u64), ("next_player", C.c_uint8), ("turn", C.c_uint8),
                ("done", C.c_uint8), ("winner", C.c_uint8)]


_lib.rzo_env_reset.argtypes = [C.POINTER(EnvStruct)]
_lib.rzo_env_update.argtypes = [C.POINTER(EnvStruct), _u64, _u64, C.c_int]
_lib.rzo_env_step.argtypes = [C.POINTER(EnvStruct), C.c_int]

find_correct_moves = _lib.rzo_find_correct_moves
flip_vertical = _lib.rzo_flip_vertical
flip_diag_a1h8 = _lib.rzo_flip_diag_a1h8
rotate90 = _lib.rzo_rotate90
rotate180 = _lib.rzo_rotate180
dihedral = _lib.rzo_dihedral
bit_count = _lib.rzo_bit_count


def calc_flip(pos, own, enemy):
    return _lib.rzo_calc_flip(int(pos), own, enemy)


def _p(a, t):
    return a.ctypes.data_as(t)


def find_correct_moves_batch(own, enemy):
    own = np.ascontiguousarray(own, np.uint64); enemy = np.ascontiguousarray(enemy, np.uint64)
    out = np.empty_like(own)
    _lib.rzo_find_correct_moves_batch(_p(own, _u64p), _p(enemy, _u64p), _p(out, _u64p), C.c_size_t(own.size))
    return out


def calc_flip_batch(pos, own, enemy):
    pos = np.ascontiguousarray(pos, np.uint8)
    own = np.ascontiguousarray(own, np.uint64); enemy = np.ascontiguousarray(enemy, np.uint64)
    out = np.empty_like(own)
    _lib.rzo_calc_flip_batch(_p(pos, _u8p), _p(own, _u64p), _p(enemy, _u64p), _p(out, _u64p), C.c_size_t(own.size))
    return out


def step_batch(black, white, next_player, turn, done, winner, action):
    """In-place SoA step; arrays must be contiguous uint64/uint8, action int8 (-1 = resign)."""
    _lib.rzo_step_batch(_p(black, _u64p), _p(white, _u64p), _p(next_player, _u8p), _p(turn, _u8p),
                        _p(done, _u8p), _p(winner, _u8p), _p(action, _i8p), C.c_size_t(black.size))


def dihedral_batch(x, t):
    x = np.ascontiguousarray(x, np.uint64); t = np.ascontiguousarray(t, np.uint8)
    out = np.empty_like(x)
    _lib.rzo_dihedral_batch(_p(x, _u64p), _p(t, _u8p), _p(out, _u64p), C.c_size_t(x.size))
    return out


def bit_to_array(x, size=64):
    """lib/bitboard.py:136-138: bit i -> array[i] (uint8)."""
    return ((int(x) >> np.arange(size, dtype=np.uint64).astype(object)) & 1).astype(np.uint8) if size > 64 else \
        ((np.uint64(x) >> np.arange(size, dtype=np.uint64)) & np.uint64(1)).astype(np.uint8)


class Env:
    """Oracle game state (env/reversi_env.py:18-130) over the C struct.  Player: 1 black, 2 white;
    winner: 0 None, 1 black, 2 white, 3 draw."""

    def __init__(self):
        self.s = EnvStruct()
        _lib.rzo_env_reset(self.s)

    def reset(self):
        _lib.rzo_env_reset(self.s)
        return self

    def update(self, black, white, next_player):
        _lib.rzo_env_update(self.s, black, white, int(next_player))
        return self

    def step(self, action):
        _lib.rzo_env_step(self.s, -1 if action is None else int(action))
        return self

    def copy(self):
        e = Env.__new__(Env)
        e.s = EnvStruct(self.s.black, self.s.white, self.s.next_player, self.s.turn, self.s.done, self.s.winner)
        return e

    black = property(lambda self: self.s.black)
    white = property(lambda self: self.s.white)
    next_player = property(lambda self: self.s.next_player)
    turn = property(lambda self: self.s.turn)
    done = property(lambda self: bool(self.s.done))
    winner = property(lambda self: self.s.winner)

    def own_enemy(self):
        return (self.s.black, self.s.white) if self.s.next_player == 1 else (self.s.white, self.s.black)

    def state(self):
        s = self.s
        return (s.black, s.white, s.next_player, s.turn, s.done, s.winner)
