"""The device endgame solver's resumable stack machine (csrc/rz_solver.cuh, host/device code) compiled for the HOST and
checked against the golden positions of the reference solver (tests/golden/solver.json) and the oracle on seeded random
endgames -- in one go and time-sliced (suspended and resumed every few node steps, as the engine does per wave).  No GPU
needed; the same cases run on the device in tests/test_solver_gpu.py."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import bitboard as ob
from oracle.solver import Solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "reversi-alpha-zero_b200", "csrc")


@pytest.fixture(scope="module")
def host_solver(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("solver_host") / "solver_host_check")
    subprocess.run([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-I", CSRC,
                    os.path.join(ROOT, "tests", "support", "solver_host_check.cu"), "-o", exe], check=True)
    return exe


def _cases(golden_dir):
    cases = []
    for c in json.load(open(os.path.join(golden_dir, "solver.json")))["positions"]:
        black_to_move = c["next_player"] == 1
        cases.append((c["black"] if black_to_move else c["white"], c["white"] if black_to_move else c["black"], int(c["exactly"]),
                      (c["move"], c["score"])))
    rng = np.random.default_rng(29)
    n = 0
    while n < 150:
        empties = int(rng.integers(1, 11))
        e = ob.Env().reset()
        while not e.done and 60 - e.turn > empties:
            o, en = e.own_enemy()
            legal = ob.find_correct_moves(o, en)
            ms = [i for i in range(64) if legal >> i & 1]
            e.step(ms[rng.integers(len(ms))])
        if e.done:
            continue
        o, en = e.own_enemy()
        for exactly in (1, 0):
            mv, sc = Solver().solve(o, en, bool(exactly))
            cases.append((o, en, exactly, (-1, 0) if mv is None else (mv, sc)))
        n += 1
    cases.append((0x0000000810000000, 0x0000001008000000, 1, (-1, 0)))  # 60 empties: refused
    cases.append((0xFFFFFFFFFFFFFF00, 0x00000000000000FE, 1, (-1, 0)))  # no legal move
    return cases


@pytest.mark.parametrize("slice_ticks", [0, 1])
def test_stack_machine_matches_reference_and_oracle(host_solver, golden_dir, slice_ticks):
    cases = _cases(golden_dir)
    text = "".join(f"{o:x} {e:x} {x}\n" for o, e, x, _ in cases)
    r = subprocess.run([host_solver, str(slice_ticks)], input=text, capture_output=True, text=True, check=True)
    got = [tuple(int(v) for v in line.split()) for line in r.stdout.strip().split("\n")]
    assert len(got) == len(cases)
    for (o, e, x, want), g in zip(cases, got):
        assert g == want, (hex(o), hex(e), x)
    resumes = int(r.stderr.split()[-1])
    assert (resumes > 1000) if slice_ticks else (resumes == 0)
