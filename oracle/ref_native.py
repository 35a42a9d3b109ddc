"""TEST INFRASTRUCTURE ONLY.  Loader of the reference's own compiled Cython modules built by ``oracle/build_ref.py``
into ``oracle/_ref`` -- the real reference code for move generation, flips and the endgame solver, usable on the GPU
box where /root/reference does not exist.

    from oracle import ref_native
    if ref_native.available():
        bb, solver_mod = ref_native.load()
        bb.find_correct_moves(own, enemy); bb.calc_flip(pos, own, enemy); solver_mod.ReversiSolver().solve(...)
"""
import enum
import importlib
import os
import sys
import types

from . import build_ref

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return build_ref.built()


def _ensure_player_module():
    """reversi_solver_cython.pyx:3 does ``from reversi_zero.env.reversi_env import Player`` and only uses
    ``Player.black`` / ``Player.white`` (env/reversi_env.py:9: black = 1, white = 2).  Where the reference package is not
    importable (GPU box) a three-module stand-in with just that enum is registered."""
    from .ref_shims import install as shims
    if shims.available():   # build container: use the real reference package (and never shadow it with the stand-in)
        shims.install()
        importlib.import_module("reversi_zero.env.reversi_env")
        return

    class Player(enum.Enum):
        black = 1
        white = 2

    for name in ("reversi_zero", "reversi_zero.env", "reversi_zero.env.reversi_env"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["reversi_zero.env.reversi_env"].Player = Player


def load():
    """-> (bitboard_cython module, reversi_solver_cython module)"""
    if not available():
        raise RuntimeError("oracle/_ref is not built (python oracle/build_ref.py in the build container)")
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    _ensure_player_module()
    bb = importlib.import_module("rzref.alt.bitboard_cython")
    sv = importlib.import_module("rzref.alt.reversi_solver_cython")
    return bb, sv


def player_enum():
    """The ``Player`` class the loaded solver module compares against (the reference's own, or the stand-in)."""
    _ensure_player_module()
    return sys.modules["reversi_zero.env.reversi_env"].Player
