"""Capacity of the host side of the output path (CPU only, no GPU): what ONE writer thread of the self-play worker sustains
when every finished game is a complete game -- SURVEY section 7's hard part (the reference writes ~0.5 MB of JSON per game).

The bench's `e2e` leg runs for ~30 s from a warm start, so the games that finish inside it recorded only the plies they
searched after their warm-start turn (profiles/bench_r02_1gpu.json: 47.6 KB of play data per game); in a long run every game
records all ~60 plies.  This tool feeds the unmodified `SelfPlayWorker._harvest()` -- the writer thread's loop body: copy of
the ctypes records, per-game bookkeeping (draw dropping, resign statistics, GGF text, game-index file, simulation schedule),
`rz_write_play_data` (the C JSON writer with the 8-symmetry expansion), pruning to `max_file_num` files -- with complete
games from a stand-in engine (random legal playouts, visit counts drawn like a 400-simulation search) and config/ch5.yml's
output settings, and reports games/s, MB/s and bytes per game on one host thread, plus the C writer alone.

    python tools/writer_bench.py [games] > profiles/writer_bench_r02.json"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "reversi-alpha-zero_b200"))


def synth_games(n_games, sims=400, seed=20260923):
    """ctypes arrays (Game[n], Ply[..]) of complete games: random legal playouts from the opening; at every ply the visit
    counts of a `sims`-simulation search are drawn from a Dirichlet-multinomial over the legal moves (turn 0: the single
    forced visit of bypass_first_move, agent/player.py:143-148)."""
    from oracle import bitboard as ob
    from reversi_zero_b200 import _cabi
    rng = np.random.default_rng(seed)
    games, plies = [], []
    for gid in range(n_games):
        e = ob.Env().reset()
        first = len(plies)
        while not e.done:
            own, enemy = e.own_enemy()
            legal = ob.find_correct_moves(own, enemy)
            ms = [i for i in range(64) if legal >> i & 1]
            p = _cabi.Ply()
            p.own, p.enemy, p.player, p.recorded, p.loops, p.waves = own, enemy, e.next_player, 1, 1, 50
            if e.turn == 0:
                a = ms[0]
                p.n_visit[a] = 1
            else:
                counts = rng.multinomial(sims, rng.dirichlet(np.full(len(ms), 0.3)))
                for m, c in zip(ms, counts):
                    p.n_visit[m] = int(c)
                a = ms[int(np.argmax(counts))] if e.turn >= 4 else ms[int(rng.choice(len(ms), p=counts / counts.sum()))]
            p.action, p.n, p.q = a, float(p.n_visit[a]), float(rng.uniform(-1, 1))
            plies.append(p)
            e.step(a)
        g = _cabi.Game()
        g.game_id, g.first_ply, g.n_plies, g.winner = gid, first, len(plies) - first, e.winner
        g.black_z = 1 if e.winner == 1 else (-1 if e.winner == 2 else 0)
        g.black, g.white, g.turn, g.resign_enabled = e.black, e.white, e.turn, 1
        g.expansions, g.simulations, g.table_nodes = 21199, 23574, 21000
        games.append(g)
    G = (_cabi.Game * n_games)(*games)
    P = (_cabi.Ply * len(plies))(*plies)
    return G, P


class StandInEngine:
    """poll_raw() of the real engine: up to 256 finished games per call, here cut from a pre-built pool (game ids renumbered)"""

    def __init__(self, G, P, total):
        from reversi_zero_b200 import _cabi
        self.G, self.P, self.total, self.given = G, P, total, 0
        self.n_pool = len(G)
        self._empty = ((_cabi.Game * 1)(), (_cabi.Ply * 1)())

    def poll_raw(self):
        n = min(256, self.n_pool, self.total - self.given)
        if n <= 0:
            return self._empty[0], 0, self._empty[1], 0
        for i in range(n):
            self.G[i].game_id = self.given + i
        self.given += n
        return self.G, n, self.P, self.G[n - 1].first_ply + self.G[n - 1].n_plies


def run(total=2000, pool=256):
    from reversi_zero_b200 import _cabi
    from reversi_zero_b200.config import Config
    from reversi_zero_b200.engine import write_play_data
    from reversi_zero_b200.worker.self_play import SelfPlayWorker
    G, P = synth_games(pool)
    n_plies = G[pool - 1].first_ply + G[pool - 1].n_plies
    tmp = tempfile.mkdtemp(prefix="rz_writer_bench_")
    try:
        # (1) the C writer alone: one file with `pool` complete games
        path = os.path.join(tmp, "all.json")
        t0 = time.perf_counter()
        n_rec = write_play_data(path, G, pool, P, True, 4)
        dt_c = time.perf_counter() - t0
        c_bytes = os.path.getsize(path)
        os.remove(path)
        # (2) the writer thread's loop body with config/ch5.yml's output settings (ch5.yml:3-7; GGF on, config.py:123-124)
        cfg = Config(project_dir=tmp, data_dir=os.path.join(tmp, "data"))
        cfg.play_data.update(dict(nb_game_in_file=1, max_file_num=800, drop_draw_game_rate=0.5, enable_ggf_data=True, nb_game_in_ggf_file=100))
        cfg.play.schedule_of_simulation_num_per_move = [(0, 400)]
        cfg.play.simulation_num_per_move = 400
        cfg.resource.create_directories()
        w = SelfPlayWorker(cfg)
        w.engine = StandInEngine(G, P, total)
        t0 = time.process_time()
        t0w = time.perf_counter()
        n = 0
        while True:
            k = w._harvest()
            if k == 0:
                break
            n += k
        w._flush_files(force=True)
        cpu_s, wall_s = time.process_time() - t0, time.perf_counter() - t0w
        return dict(what="one writer thread of SelfPlayWorker on complete games (stand-in engine, CPU only), config/ch5.yml output settings",
                    games=n, plies_per_game=n_plies / pool, files_written=len(w.files_written), bytes_written=w.bytes_written,
                    bytes_per_written_game=w.bytes_written / max(1, len(w.files_written)),
                    wall_seconds=wall_s, cpu_seconds=cpu_s, games_per_sec=n / wall_s, mb_per_sec=w.bytes_written / wall_s / 1e6,
                    ms_per_game=wall_s / n * 1e3,
                    c_writer_alone=dict(games=pool, records=int(n_rec), bytes=c_bytes, seconds=dt_c, mb_per_sec=c_bytes / dt_c / 1e6,
                                        games_per_sec=pool / dt_c),
                    headroom_over_one_b200=dict(selfplay_games_per_sec_per_gpu=49.0, writer_thread_capacity_over_it=n / wall_s / 49.0),
                    host=dict(cores_usable=len(os.sched_getaffinity(0)), note="build container, tmpfs/overlay /tmp"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(json.dumps(run(*(int(a) for a in sys.argv[1:])), indent=1))
