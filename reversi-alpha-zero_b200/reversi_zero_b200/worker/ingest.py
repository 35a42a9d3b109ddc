"""Trainer-side ingest: what ``OptimizeWorker.load_play_data`` / ``convert_to_training_data``
(worker/optimize.py:165-231) do with ``play_*.json`` -- here from the compact row files the self-play worker can
write next to them (``play_*.rzrows``, 280 bytes per recorded ply instead of ~5 KB of JSON per ply), expanded to
training arrays on the device by ``rz_ingest`` (csrc/rz_ingest.cu).  No CPU fallback.

    rows, tau1, ctt = read_play_rows("data/play_data/play_20260922-101500.123456.rzrows")
    states, policy, z = to_training_arrays(rows, tau1, ctt)              # numpy, as the reference trainer holds them
    states_d, policy_d, z_d = to_training_tensors(rows, tau1, ctt, 0)    # torch tensors on cuda:0 for a device-side trainer
"""
import ctypes as C
import os
from glob import glob

import numpy as np

from .. import _cabi

ROW_DTYPE = np.dtype([("own", "<u8"), ("enemy", "<u8"), ("n_visit", "<i4", (64,)), ("z", "<i4"), ("pad", "<i4")])
assert ROW_DTYPE.itemsize == C.sizeof(_cabi.PlayRow) == 280

ROWS_SUFFIX = ".rzrows"


def rows_path_of(json_path):
    """play_<ts>.json -> play_<ts>.rzrows (not matched by the trainer's ``play_*.json`` glob, lib/data_helper.py:11-14)."""
    return os.path.splitext(json_path)[0] + ROWS_SUFFIX


def write_play_rows(path, games, n_games, plies, save_policy_of_tau_1=True, change_tau_turn=4):
    """games / plies: ctypes arrays as returned by Engine.poll_raw(); same games and order as write_play_data."""
    n = C.c_size_t()
    _cabi.check(_cabi.lib().rz_write_play_rows(path.encode(), games, n_games, plies, int(bool(save_policy_of_tau_1)),
                                                int(change_tau_turn), C.byref(n)), "rz_write_play_rows")
    return n.value


def read_play_rows(path):
    """-> (rows: structured numpy array of ROW_DTYPE, save_policy_of_tau_1: bool, change_tau_turn: int)"""
    L = _cabi.lib()
    n, tau1, ctt = C.c_size_t(), C.c_int(), C.c_int()
    _cabi.check(L.rz_read_play_rows(path.encode(), None, 0, C.byref(n), C.byref(tau1), C.byref(ctt)), "rz_read_play_rows")
    rows = np.zeros(n.value, ROW_DTYPE)
    if n.value:
        _cabi.check(L.rz_read_play_rows(path.encode(), rows.ctypes.data_as(C.c_void_p), n.value, C.byref(n), None, None), "rz_read_play_rows")
    return rows, bool(tau1.value), int(ctt.value)


def make_rows(own, enemy, n_visit, z):
    rows = np.zeros(len(own), ROW_DTYPE)
    rows["own"], rows["enemy"], rows["n_visit"], rows["z"] = own, enemy, n_visit, z
    return rows


def to_training_arrays(rows, save_policy_of_tau_1=True, change_tau_turn=4):
    """numpy twin of convert_to_training_data: (states uint8 [N,2,8,8], policy float32 [N,64], z float32 [N]), N = 8 * rows."""
    rows = np.ascontiguousarray(rows, ROW_DTYPE)
    n = len(rows)
    states = np.empty((8 * n, 2, 8, 8), np.uint8)
    policy = np.empty((8 * n, 64), np.float32)
    z = np.empty((8 * n,), np.float32)
    _cabi.check(_cabi.lib().rz_ingest(rows.ctypes.data_as(C.c_void_p), n, int(bool(save_policy_of_tau_1)), int(change_tau_turn),
                                       states.ctypes.data_as(_cabi.u8p), policy.ctypes.data_as(_cabi.f32p), z.ctypes.data_as(_cabi.f32p)),
                "rz_ingest")
    return states, policy, z


def to_training_tensors(rows, save_policy_of_tau_1=True, change_tau_turn=4, device=0):
    """Rows -> torch tensors that stay on the device (one H2D copy of the compact rows, the expansion happens in HBM)."""
    import torch
    dev = torch.device("cuda", device)
    rows = np.ascontiguousarray(rows, ROW_DTYPE)
    n = len(rows)
    with torch.cuda.device(dev):
        d_rows = torch.from_numpy(rows.view(np.uint8).reshape(-1)).to(dev)
        states = torch.empty((8 * n, 2, 8, 8), dtype=torch.uint8, device=dev)
        policy = torch.empty((8 * n, 64), dtype=torch.float32, device=dev)
        z = torch.empty((8 * n,), dtype=torch.float32, device=dev)
        _cabi.check(_cabi.lib().rz_ingest_dev(d_rows.data_ptr(), n, int(bool(save_policy_of_tau_1)), int(change_tau_turn), states.data_ptr(),
                                               policy.data_ptr(), z.data_ptr(), torch.cuda.current_stream().cuda_stream), "rz_ingest_dev")
    return states, policy, z


def load_play_data_dir(play_data_dir, device=0):
    """All row files of a play_data directory (the trainer's ``load_play_data``, worker/optimize.py:165-180) as one
    device-resident dataset: (states, policy, z) torch tensors, files in sorted order like get_game_data_filenames."""
    import torch
    parts = []
    for path in sorted(glob(os.path.join(play_data_dir, "play_*" + ROWS_SUFFIX))):
        rows, tau1, ctt = read_play_rows(path)
        if len(rows):
            parts.append(to_training_tensors(rows, tau1, ctt, device))
    if not parts:
        return None
    return tuple(torch.cat([p[i] for p in parts]) for i in range(3))
