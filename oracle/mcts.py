"""CPU restatement of the reference's MCTS player and self-play game loop.  TEST INFRASTRUCTURE ONLY.

Follows agent/player.py (ReversiPlayer) and worker/self_play.py:139-175,219-238 of the reference.
It is at the same time the executable specification of the CUDA engine's search: the engine must
reproduce this file's visit counts / moves / records EXACTLY when both are driven by the same
evaluator and the same Philox streams (tests/test_engine_parity.py), and this file in turn is pinned
against the unmodified reference by tests/golden/make_golden.py (exact at parallel_search_num = 1
with a deterministic evaluator; statistical otherwise -- the reference's own visit counts depend on
asyncio timer interleaving and numpy's MT19937 stream, agent/player.py:253-254,300-301).

Differences from the reference, all deliberate and documented in DESIGN.md:
  * statistics are stored once per position in the side-to-move's frame; the reference stores them
    under CounterKey(black, white, next_player) AND under the colour-swapped mirror key with the
    sign of W flipped (player.py:276-280,388-393).  The two are the same table (proof in DESIGN.md).
  * asyncio coroutines (player.py:189-215) become explicit waves: up to parallel_search_num
    simulations are in flight; a simulation that reaches a position which another in-flight
    simulation is expanding parks with its virtual loss applied (player.py:253-254) and resumes after
    the batch evaluation.
  * W is accumulated in float32 (reference: float64); selection arithmetic follows numpy's dtype
    promotion on the reference's expressions (float32 priors, float64 Q/U).
  * RNG: Philox4x32-10 streams (oracle/philox.py) instead of numpy MT19937.
"""
import time

import numpy as np

from . import bitboard as bb
from . import philox as px
from .solver import Solver

f32, f64 = np.float32, np.float64


class PlayParams:
    """Subset of PlayConfig / PlayDataConfig (config.py:116-166) that the self-play path reads."""

    def __init__(self, **kw):
        self.simulation_num_per_move = 200
        self.thinking_loop = 1
        self.required_visit_to_decide_action = 400
        self.start_rethinking_turn = 8
        self.c_puct = 1
        self.noise_eps = 0.25
        self.dirichlet_alpha = 0.5
        self.change_tau_turn = 4
        self.virtual_loss = 3
        self.parallel_search_num = 8
        self.resign_threshold = None
        self.allowed_resign_turn = 20
        self.disable_resignation_rate = 0.1
        self.share_mtcs_info_in_self_play = True
        self.save_policy_of_tau_1 = True
        self.use_solver_turn = 0                # config.py:154; 0 = solver off (agent/player.py:100)
        self.use_solver_turn_in_simulation = 0  # config.py:155 (agent/player.py:237-238)
        self.reset_mtcs_info_per_game = 1       # config.py:131
        self.max_sims_per_wave = 0  # engine knob (0 = 2 * parallel_search_num): simulations started per game per wave
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class Node:
    __slots__ = ("legal", "legal_arr", "P", "pn", "N", "W", "exp", "wld")

    def __init__(self, legal):
        self.legal = legal
        self.legal_arr = bb.bit_to_array(legal, 64)
        self.P = np.zeros(64, f32)
        self.pn = np.zeros(64, f32)
        self.N = np.zeros(64, np.int64)
        self.W = np.zeros(64, f32)
        self.wld = None  # cached (action, score) of the in-simulation WLD solve of this position
        self.exp = 0  # bit (pid-1) set once player pid expanded it (each ReversiPlayer has its own
        #               `expanded` set even when N/W/P dicts are shared, player.py:44-47)


def np_sum_f32_64(a):
    """numpy's float32 add.reduce over 64 contiguous elements: 8 running column sums, then a fixed tree
    (numpy/core/src/umath/loops_utils.h.src pairwise_sum, n <= 128 branch).  The engine uses the same order."""
    r = a[0:8].astype(f32).copy()
    for i in range(8, 64, 8):
        r = (r + a[i:i + 8]).astype(f32)
    return f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))


def normalize_prior(P, legal_arr):
    """player.py:406-413 with temperature == 1 (turn <= policy_decay_turn = 60 always holds)."""
    p = (P * legal_arr).astype(f32)
    s = np_sum_f32_64(p)
    if s > 0:
        p = (p / s).astype(f32)
    return p


def square_map(t):
    """T_t[s] = square that s moves to under dihedral transform t (flip_vertical if t&4, then t&3 x rotate90)."""
    return np.array([bb.dihedral(1 << s, t).bit_length() - 1 for s in range(64)], dtype=np.int64)


_SQMAP = [square_map(t) for t in range(8)]


def inverse_policy(p_t, t):
    """player.py:315-321: policy predicted on the transformed board -> original frame."""
    return np.asarray(p_t)[_SQMAP[t]]


def select_action(node, is_root, pp, noise=None):
    """player.py:395-428 in the mover's frame.  Returns the first index of the maximum."""
    N = node.N
    xx = max(np.sqrt(f64(int(N.sum()))), 1.0)
    if is_root and pp.noise_eps > 0:
        t32 = (f32(1 - pp.noise_eps) * node.pn).astype(f32)
        p = t32.astype(f64) + pp.noise_eps * noise
        u = pp.c_puct * p * xx / (1.0 + N)
    else:
        c32 = (f32(pp.c_puct) * node.pn).astype(f32)
        u = c32.astype(f64) * xx / (1.0 + N)
    q = node.W.astype(f64) / (N + 1e-5)
    v = (q + u + 1000.0) * node.legal_arr
    return int(np.argmax(v))


class Descent:
    __slots__ = ("env", "path")

    def __init__(self, env):
        self.env = env
        self.path = []  # (node, action, mover_is_root)


class TimeUp(Exception):
    """raised inside a search when SelfPlayGame.deadline (time.perf_counter value) has passed"""


class SelfPlayGame:
    """One self-play game: both players, shared (or separate) statistics, compact per-ply log."""
    deadline = None  # optional wall-clock bound used by the CPU-baseline timing (oracle/selfplay_cpu.py)

    def __init__(self, pp, api, seed=0, game_id=0, noise_rng=None, api_b=None, black_net=0, table=None):
        """api_b / black_net: evaluation matches (worker/evaluate.py:66-96) -- the player of colour `pid` is
        evaluated by `api` when it plays for network 0 and by `api_b` when it plays for network 1; black_net says
        which network has the black stones."""
        self.pp, self.api, self.seed, self.game_id = pp, api, seed, game_id
        self.api_b, self.black_net = api_b, black_net
        self.noise_rng = noise_rng or np.random.default_rng((seed, game_id))
        self.env = bb.Env().reset()
        # reset_mtcs_info_per_game > 1 (worker/self_play.py:111-134): statistics carried over from the previous game of the
        # same worker; the new players treat every position with a prior as expanded (player.py:47)
        self.table = table if table is not None else {}
        for node in self.table.values():
            node.exp = 3
        self.n_expand = 0      # leaves sent to the evaluator (== "node expansions")
        self.n_rootsel = 0
        self.n_sims = 0
        self.n_waves = 0
        self.plies = []        # dicts: pid, own, enemy, N, policy, saved_policy, action, n, q, loops
        self.resigned = {1: False, 2: False}
        self.solver = {1: Solver(), 2: Solver()}  # one ReversiSolver per ReversiPlayer (player.py:60,435-436)
        self.n_solves = 0
        self.solved_plies = []
        self.enable_resign = pp.disable_resignation_rate <= px.u01(px.draw(seed, game_id, 0, px.P_GAME)[0])

    # -- keys ------------------------------------------------------------------------------------
    def _key(self, own, enemy, pid):
        return (own, enemy) if self.pp.share_mtcs_info_in_self_play else (own, enemy, pid)

    def _dirichlet(self, legal):
        n = bb.bit_count(legal)
        g = self.noise_rng.standard_gamma(self.pp.dirichlet_alpha, n)
        g = g / g.sum()
        out = np.zeros(64, f64)
        out[np.nonzero(bb.bit_to_array(legal, 64))[0]] = g  # ascending bit order, lib/bitboard.py:162-171
        return out

    # -- one simulation step machine ----------------------------------------------------------------
    def _run(self, d, pid, pending, pending_keys):
        """Advance descent d until it terminates ('done'), needs an evaluation ('pending') or parks."""
        pp, vl = self.pp, self.pp.virtual_loss
        while True:
            e = d.env
            if e.done:  # player.py:226-232
                v = 0.0 if e.winner == 3 else (1.0 if e.winner == pid else -1.0)
                self._backup(d.path, v)
                return "done"
            own, enemy = e.own_enemy()
            key = self._key(own, enemy, pid)
            if pp.use_solver_turn_in_simulation and e.turn >= pp.use_solver_turn_in_simulation:  # player.py:237-251
                node = self.table.get(key)
                if node is not None and node.wld is not None:
                    if self._apply_wld(d, node, e.next_player == pid):
                        return "done"
                elif key in pending_keys:
                    return "parked"
                else:  # the engine solves on the device between two waves: request, like a network evaluation
                    pending_keys.add(key)
                    pending.append((d, key, own, enemy, -1, e.next_player == pid))
                    return "pending"
            if key in pending_keys:  # player.py:253-254
                return "parked"
            node = self.table.get(key)
            if node is None or not (node.exp >> (pid - 1)) & 1:  # player.py:257
                t = px.draw(self.seed, self.game_id, self.n_expand, px.P_DIHEDRAL)
                flip = 4 if px.u01(t[0]) < 0.5 else 0          # player.py:300
                rot = int(px.u01(t[1]) * 4)                     # player.py:301
                self.n_expand += 1
                pending_keys.add(key)
                pending.append((d, key, own, enemy, flip | rot, e.next_player == pid))
                return "pending"
            is_root = not d.path
            noise = None
            if is_root and pp.noise_eps > 0:
                noise = self._dirichlet(node.legal)
                self.n_rootsel += 1
            a = select_action(node, is_root, pp, noise)
            node.N[a] += vl                                   # player.py:270-271 (mover's frame)
            node.W[a] = f32(node.W[a] - f32(vl))
            d.path.append((node, a, e.next_player == pid))
            e.step(a)

    def _apply_wld(self, d, node, mover_is_root):
        """player.py:239-251 with the cached result (action, score) of the WLD solve; returns False when the reference's
        `if action:` test fails (no result, or the solved move is square 0)."""
        action, score = node.wld
        if not action:
            return False
        sgn = f32(np.sign(score))                      # value for the side to move at this node
        node.N[action] += 1
        node.W[action] = f32(node.W[action] + sgn)
        node.P = np.zeros(64, f32)
        node.P[action] = 1
        node.pn = node.P.copy()
        self._backup(d.path, float(sgn) if mover_is_root else -float(sgn))
        return True

    def _backup(self, path, v_root):
        """player.py:276-280."""
        vl = self.pp.virtual_loss
        for node, a, mover_is_root in path:
            node.N[a] += 1 - vl
            node.W[a] = f32(node.W[a] + f32(f32(vl) + f32(v_root if mover_is_root else -v_root)))

    def search(self, own, enemy, pid):
        """player.py:189-215: simulation_num_per_move simulations, <= parallel_search_num in flight."""
        pp = self.pp
        S, K = pp.simulation_num_per_move, pp.parallel_search_num
        started, parked = 0, []
        cap = pp.max_sims_per_wave or 2 * K
        while True:
            pending, pending_keys, still_parked = [], set(), []
            started_wave = 0
            for d in parked:
                r = self._run(d, pid, pending, pending_keys)
                if r == "parked":
                    still_parked.append(d)
            while started < S and len(pending) + len(still_parked) < K and started_wave < cap:
                started += 1
                started_wave += 1
                d = Descent(bb.Env().update(own if pid == 1 else enemy, enemy if pid == 1 else own, pid))
                r = self._run(d, pid, pending, pending_keys)
                if r == "parked":
                    still_parked.append(d)
            parked = still_parked
            if not pending:
                assert not parked
                if started >= S:
                    break
                continue
            self.n_waves += 1
            parked = parked + self._evaluate(pending, pid)
            if self.deadline is not None and time.perf_counter() > self.deadline:
                self.n_sims += started
                raise TimeUp()
        self.n_sims += started

    def _evaluate(self, pending, pid):
        """player.py:283-327 for a whole wave (network requests) + the WLD solves requested in this wave; results are
        consumed in request order."""
        nn = [x for x in pending if x[4] >= 0]
        policy = value = None
        resume = []
        if nn:
            t_own = np.array([bb.dihedral(o, t) for (_, _, o, e, t, _) in nn], dtype=np.uint64)
            t_en = np.array([bb.dihedral(e, t) for (_, _, o, e, t, _) in nn], dtype=np.uint64)
            sh = np.arange(64, dtype=np.uint64)
            planes = np.stack([((t_own[:, None] >> sh) & np.uint64(1)), ((t_en[:, None] >> sh) & np.uint64(1))],
                              axis=1).astype(np.uint8).reshape(-1, 2, 8, 8)
            net = 0 if self.api_b is None else (self.black_net if pid == 1 else 1 - self.black_net)
            policy, value = (self.api_b if net else self.api).predict(planes)
        i = 0
        for (d, key, own, enemy, t, mover_is_root) in pending:
            node = self.table.get(key)
            if node is None:
                node = self.table[key] = Node(bb.find_correct_moves(own, enemy))
            if t < 0:  # WLD solve (player.py:239)
                self.n_solves += 1
                mv, sc = self.solver[pid].solve(own, enemy, exactly=False)
                node.wld = (mv, sc)
                if not self._apply_wld(d, node, mover_is_root):
                    resume.append(d)  # `if action:` failed: the simulation goes on from this node in the next wave
                continue
            node.P = inverse_policy(np.asarray(policy[i], f32), t).astype(f32)
            node.pn = normalize_prior(node.P, node.legal_arr)
            node.exp |= 1 << (pid - 1)
            v = float(np.asarray(value[i]).reshape(-1)[0])
            i += 1
            self._backup(d.path, v if mover_is_root else -v)
        return resume

    # -- per-ply decision ---------------------------------------------------------------------------
    def decide(self, own, enemy, pid):
        """player.py:82-134 (solver disabled).  Returns action or None (resign)."""
        pp = self.pp
        turn = bb.bit_count(own) + bb.bit_count(enemy) - 4
        key = self._key(own, enemy, pid)
        ply = len(self.plies)
        loops = 0
        if pp.use_solver_turn and turn >= pp.use_solver_turn:  # action_by_searching, player.py:100-103,150-161
            mv, sc = self.solver[pid].solve(own, enemy, exactly=True)
            if mv is not None:
                root = self.table.get(key)
                if root is None:
                    root = self.table[key] = Node(bb.find_correct_moves(own, enemy))
                sgn = float(np.sign(sc))
                root.N[mv] = 999
                root.W[mv] = f32(sgn * 999)
                root.P = np.zeros(64, f32)
                root.P[mv] = 1
                root.pn = root.P.copy()
                self.solved_plies.append(dict(pid=pid, own=own, enemy=enemy, action=mv, n=999.0, q=sgn, turn=turn))
                return mv  # not saved as play data (player.py:102)
        for tl in range(pp.thinking_loop):
            loops += 1
            if turn > 0:
                self.search(own, enemy, pid)
                root = self.table[key]
            else:  # bypass_first_move, player.py:143-148
                root = self.table.get(key)
                if root is None:
                    root = self.table[key] = Node(bb.find_correct_moves(own, enemy))
                a0 = int(np.argmax(root.legal_arr))
                root.N[a0] = 1
                root.W[a0] = 0
                root.P = (root.legal_arr / np.sum(root.legal_arr)).astype(f32)
                root.pn = normalize_prior(root.P, root.legal_arr)
            N = root.N
            tau1 = N / np.sum(N)                                           # player.py:384-385
            if turn < pp.change_tau_turn:
                policy = tau1
            else:
                policy = np.zeros(64)
                policy[int(np.argmax(N))] = 1                            # player.py:378-382
            r = px.draw(self.seed, self.game_id, ply * 16 + tl, px.P_MOVE)
            cdf = np.cumsum(policy)
            cdf /= cdf[-1]
            action = int(np.searchsorted(cdf, px.u53(r[0], r[1]), side="right"))  # np.random.choice, player.py:112
            q = root.W.astype(f64) / (N + 1e-5)
            abv = int(np.argmax(q + (N > 0) * 100))
            value_diff = q[action] - q[abv]
            if turn <= pp.start_rethinking_turn or \
                    (value_diff > -0.01 and N[action] >= pp.required_visit_to_decide_action):
                break
        rec = dict(pid=pid, own=own, enemy=enemy, N=N.copy(), policy=policy, action=action,
                   n=float(N[action]), q=float(q[action]), loops=loops, turn=turn)
        if pp.resign_threshold is not None and np.max(q - (N == 0) * 10) <= pp.resign_threshold:  # player.py:123-130
            self.resigned[pid] = True
            if self.enable_resign and turn >= pp.allowed_resign_turn:
                return None
        rec["saved_policy"] = tau1 if pp.save_policy_of_tau_1 else policy
        self.plies.append(rec)
        return action

    def play(self):
        """worker/self_play.py:139-175 game loop; returns self."""
        e = self.env
        self.actions = []
        while not e.done:
            own, enemy = e.own_enemy()
            a = self.decide(own, enemy, e.next_player)
            self.actions.append(a)
            e.step(a)
        self.black_z = {1: 1, 2: -1, 3: 0}[e.winner]  # self_play.py:219-231
        return self

    def records(self):
        """Reference-format training records: all of black's then all of white's (self_play.py:183),
        each ply expanded to 8 symmetries (player.py:166-179), z appended (player.py:357-364)."""
        out = []
        for pid, z in ((1, self.black_z), (2, -self.black_z)):
            for rec in self.plies:
                if rec["pid"] != pid:
                    continue
                for own_s, enemy_s, pol_s in symmetries8(rec["own"], rec["enemy"], rec["saved_policy"]):
                    out.append([[own_s, enemy_s], list(pol_s), z])
        return out


def symmetries8(own, enemy, policy):
    """player.py:166-179: order flip in (F,T) x rot_right in 0..3; policy flipud then rot90(k=-rot)."""
    for flip in (False, True):
        for rot in range(4):
            o, e, p = own, enemy, np.asarray(policy, dtype=f64).reshape(8, 8)
            if flip:
                o, e, p = bb.flip_vertical(o), bb.flip_vertical(e), np.flipud(p)
            for _ in range(rot):
                o, e = bb.rotate90(o), bb.rotate90(e)
            if rot:
                p = np.rot90(p, k=-rot)
            yield o, e, p.reshape(64)
