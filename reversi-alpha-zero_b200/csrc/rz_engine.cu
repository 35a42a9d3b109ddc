// rz_engine.cu -- K2/K3/K6/K7: on-device MCTS self-play (agent/player.py ReversiPlayer,
// worker/self_play.py game loop) for thousands of concurrent games.
//
// Data layout (HBM, per game slot, flat arrays):
//   * transposition table: open-addressing hash (uint32 slots, generation-tagged so a new game needs no
//     clearing) -> node index; keys are (own, enemy) in the side-to-move's frame.  The reference keeps
//     statistics under CounterKey(black, white, next_player) and under the colour-swapped mirror key
//     with W negated (player.py:276-280,388-393): that is one table in the mover's frame.
//   * nodes (32 B): key, legal-move mask, first-edge index, per-player "expanded" bits
//     (each ReversiPlayer has its own `expanded` set even when statistics are shared, player.py:44-47).
//   * edges (16 B, legal moves only, ascending square order): visit count N, value sum W (fp32, mover's
//     frame), prior P already re-normalised over the legal moves (player.py:406-413).
//   * K descent slots (parallel_search_num) with the search path, K-entry pending / parked lists.
//   * two ply logs (double-buffered mailboxes) from which the host harvests finished games.
// One wave = one `tick` kernel (consume the previous evaluations: expand + backup; decide moves, step
// games, start new games; run up to K descents per game with virtual loss; gather the leaves with a
// warp-scan into one compact batch, already dihedral-transformed) + one network launch on the batch
// whose size the network kernel reads from device memory (no host round trip inside a wave).
// The arithmetic of selection / backup / move choice follows oracle/mcts.py operation by operation
// (fp32 priors and W, fp64 Q/U, numpy summation order), so that parity tests can demand exact equality.
// This translation unit is compiled with -fmad=false for that reason.
#include <deque>
#include <new>
#include <vector>
#include <stdlib.h>
#include <string.h>
#include "rz_bitboard.cuh"
#include "rz_net.cuh"
#include "rz_solver.cuh"

namespace rz {
namespace solver {
constexpr int kBlockThreads = 128;
// The engine's solver step: advance every unfinished request of a slot group for at most `budget_ns`, then queue what is
// still unfinished for the group's next wave.  `active` holds request-context indices: the unfinished ones of the last
// wave followed by those the tick kernel just added.  One CTA per SM, so all of them are resident beside the other
// group's network kernel and the step costs at most about `budget_ns` of stream time.
__global__ void __launch_bounds__(kBlockThreads) solve_active_kernel(SolveCtx* __restrict__ ctx, const uint32_t* __restrict__ active,
                                                                     const uint32_t* __restrict__ n_active, uint32_t* __restrict__ next,
                                                                     uint32_t* __restrict__ n_next, u64* tt_base, long long budget_ns) {
    const uint32_t n = *n_active;
    if (n == 0) return;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, total = gridDim.x * blockDim.x, n_warps = total >> 5;
    const long long deadline = global_ns() + budget_ns;
    const TT tt{tt_base + (size_t)tid * kTtEntries * kTtWordsPerEntry};
    // request r -> lane r / n_warps of warp r % n_warps: spreads a short list over all warps (few divergent lanes per warp)
    for (uint32_t r = (tid & 31u) * n_warps + (tid >> 5); r < n; r += total) {
        const uint32_t idx = active[r];
        if (!solve_advance(ctx + idx, tt, deadline)) next[atomicAdd(n_next, 1u)] = idx;
    }
}
}  // namespace solver

namespace eng {

// ---- Philox4x32-10 (same streams as oracle/philox.py) ------------------------------------------------
enum { P_DIHEDRAL = 0, P_MOVE = 1, P_NOISE = 2, P_GAME = 3 };

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}
__device__ __forceinline__ U4 draw(uint64_t seed, uint64_t game_id, uint32_t seq, uint32_t purpose, uint32_t idx) {
    return philox((uint32_t)game_id, seq, purpose, idx, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// ---- device-side structures ---------------------------------------------------------------------------
struct Node {  // 32 B
    u64 own, enemy, legal;
    uint32_t edge_base;
    uint8_t exp;   // bit (pid-1): expanded by that player
    uint8_t kpid;  // 0 in shared mode, else the player the table entry belongs to
    uint16_t pad;
};
struct __align__(16) Edge { int32_t n; float w; float p; int32_t pad; };

enum : uint8_t { D_FREE = 0, D_PENDING = 1, D_PARKED = 2 };
constexpr int kMaxK = 16;
constexpr int kMaxPath = 64;

struct Descent {
    u64 black, white;        // position reached so far
    u64 leaf_own, leaf_enemy;
    uint32_t leaf_index;     // row in the evaluation batch
    uint8_t next_player, status, dihedral, path_len;
    uint8_t leaf_mover_is_root;
    uint8_t kept;            // 1: the network result of this leaf was copied to keep_policy / keep_value
    uint8_t pad[2];
    uint32_t path[kMaxPath];  // edge index within the slot's arena | (mover_is_root << 31)
};

enum : uint8_t { PH_IDLE = 0, PH_SEARCH = 1, PH_DECIDE = 2, PH_NEWGAME = 3, PH_SOLVE = 4 /* waiting for the exact root solve */ };
constexpr uint8_t kSolveMarker = 0xFF;  // Descent::dihedral of a descent that waits for a WLD solve instead of a network evaluation

struct Slot {
    EnvState env;            // the real game
    u64 game_id;
    u64 games_played;        // games this slot has started
    uint32_t gen;            // hash generation tag of the current game (1..4095)
    uint32_t n_nodes, n_edges;
    uint32_t n_expand, n_rootsel, n_sims, sims_started, sims_target;
    uint32_t ply;            // decided plies recorded in the log so far
    uint8_t phase, tl, log_sel, enable_resign;
    uint8_t resigned_mask, search_only, n_pending, n_parked;
    uint8_t pending[kMaxK], parked[kMaxK];
    u64 root_own, root_enemy;
    uint8_t root_pid, black_net, cur_net;  // black_net / cur_net: evaluation matches (two networks)
    uint8_t root_req;         // 1: an exact root solve has to be put into this wave's solver batch
    uint8_t pad[4];
    uint32_t n_solves, n_searched_plies;
};

struct Status {
    unsigned long long games_started, games_finished, expansions, simulations, plies, idle_slots;
    unsigned long long max_nodes, max_edges;
    int error;  // 0 or RZ_E*
    int pad;
};

struct DevCfg {
    int G, S, K, vl, change_tau_turn, thinking_loop, required_visit, start_rethinking_turn, allowed_resign_turn;
    int use_resign, share, max_plies, warm_start, sims_cap, two_nets, solver_turn, solver_sim_turn, keep_games;
    float c_puct, noise_eps, alpha, resign_threshold, disable_resignation_rate;
    u64 seed, first_game_id, game_id_stride, max_games;
    uint32_t nodes_cap, edges_cap, hash_cap;  // per slot (hash_cap is a power of two)
};

struct DevPtrs {
    Slot* slots;
    Descent* desc;         // [G][K]
    uint32_t* hash;        // [G][hash_cap]
    Node* nodes;           // [G][nodes_cap]
    Edge* edges;           // [G][edges_cap]
    rz_ply* plies;         // [G][2][max_plies]
    rz_game* mail_hdr;     // [G][2]
    uint8_t* mail_flag;    // [G][2]  1 = finished game waiting for the host
    Status* status;
    uint32_t* batch_count; // leaves in the current batch
    u64* batch_own;        // [G*K] transformed, side-to-move frame
    u64* batch_enemy;
    float* policy;         // [G*K][64]
    float* value;          // [G*K]
    // endgame solver: one resumable request context per descent (+ one per slot for the exact root solve), the lists of
    // unfinished requests (per group, double-buffered by wave parity) and the network results a waiting slot has to keep
    solver::SolveCtx* sctx;  // [G][K + 1]
    uint32_t* sactive;       // [2 groups][2 parities][G * (K + 1)]
    uint32_t* solve_count;   // [2 groups][2 parities] (64 words apart)
    float* keep_policy;      // [G*K][64]
    float* keep_value;       // [G*K]
    u64* solver_tt;          // per-lane transposition tables of the solver kernel, one set per slot group
};

__device__ __forceinline__ uint32_t hash_key(u64 own, u64 enemy, uint32_t kpid) {
    u64 h = own * 0x9E3779B97F4A7C15ULL ^ (enemy + 0x7F4A7C15ULL) * 0xC2B2AE3D27D4EB4FULL ^ (u64)kpid * 0x165667B19E3779F9ULL;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
    return (uint32_t)h;
}

struct Ctx {
    const DevCfg& c;
    const DevPtrs& p;
    int s;  // slot index
    Slot& sl;
    Descent* desc;
    uint32_t* hash;
    Node* nodes;
    Edge* edges;
    __device__ Ctx(const DevCfg& c_, const DevPtrs& p_, int s_)
        : c(c_), p(p_), s(s_), sl(p_.slots[s_]), desc(p_.desc + (size_t)s_ * c_.K), hash(p_.hash + (size_t)s_ * c_.hash_cap),
          nodes(p_.nodes + (size_t)s_ * c_.nodes_cap), edges(p_.edges + (size_t)s_ * c_.edges_cap) {}

    __device__ void fail(int code) { atomicMin(&p.status->error, code); }

    __device__ int find_node(u64 own, u64 enemy, uint32_t kpid) const {
        const uint32_t mask = c.hash_cap - 1;
        uint32_t h = hash_key(own, enemy, kpid) & mask;
        for (uint32_t probe = 0; probe < c.hash_cap; ++probe) {
            const uint32_t e = hash[h];
            if ((e >> 20) != sl.gen) return -1;  // empty (or stale generation)
            const uint32_t idx = (e & 0xFFFFFu) - 1;
            const Node& nd = nodes[idx];
            if (nd.own == own && nd.enemy == enemy && nd.kpid == kpid) return (int)idx;
            h = (h + 1) & mask;
        }
        return -1;
    }

    // creates the node (N = W = 0 for every legal move); returns its index or -1 on arena overflow
    __device__ int create_node(u64 own, u64 enemy, uint32_t kpid) {
        const u64 legal = find_correct_moves(own, enemy);
        const uint32_t nl = (uint32_t)popc64(legal);
        if (sl.n_nodes >= c.nodes_cap || sl.n_edges + nl > c.edges_cap || sl.n_nodes >= 0xFFFFEu) { fail(RZ_ECAPACITY); return -1; }
        const uint32_t idx = sl.n_nodes++;
        Node nd; nd.own = own; nd.enemy = enemy; nd.legal = legal; nd.edge_base = sl.n_edges; nd.exp = 0; nd.kpid = (uint8_t)kpid; nd.pad = 0;
        nodes[idx] = nd;
        for (uint32_t i = 0; i < nl; ++i) edges[sl.n_edges + i] = Edge{0, 0.f, 0.f, 0};
        sl.n_edges += nl;
        const uint32_t mask = c.hash_cap - 1;
        uint32_t h = hash_key(own, enemy, kpid) & mask;
        while ((hash[h] >> 20) == sl.gen) h = (h + 1) & mask;
        hash[h] = (sl.gen << 20) | (idx + 1);
        return (int)idx;
    }

    __device__ uint32_t kpid_of(int pid) const { return c.share ? 0u : (uint32_t)pid; }

    // ---- player.py:276-280 --------------------------------------------------------------------------
    __device__ void backup(const Descent& d, float v_root) {
        const float vl = (float)c.vl;
        for (int i = 0; i < d.path_len; ++i) {
            Edge& e = edges[d.path[i] & 0x7FFFFFFFu];
            e.n += 1 - c.vl;
            const float sv = (d.path[i] >> 31) ? v_root : -v_root;
            e.w = e.w + (vl + sv);
        }
    }

    // ---- lib/bitboard.py:162-171 dirichlet_noise_of_mask: Gamma(alpha) draws, normalised ---------------
    __device__ double gamma_draw(uint32_t rootsel, uint32_t child) {
        const double alpha = (double)c.alpha;
        const double a = alpha < 1.0 ? alpha + 1.0 : alpha;  // Marsaglia-Tsang with the alpha < 1 boost
        const double dd = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
        for (uint32_t att = 0; att < 16; ++att) {
            const U4 r = draw(c.seed, sl.game_id, rootsel, P_NOISE, child * 16 + att);
            const double z = sqrt(-2.0 * log(u01(r.x))) * cospi(2.0 * u01(r.y));
            const double t = 1.0 + cc * z;
            if (t <= 0.0) continue;
            const double v = t * t * t, u = u01(r.z);
            if (log(u) < 0.5 * z * z + dd - dd * v + dd * log(v)) {
                double g = dd * v;
                if (alpha < 1.0) g *= pow(u01(r.w), 1.0 / alpha);
                return g;
            }
        }
        return dd;
    }

    // ---- player.py:395-428 select_action_q_and_u (mover's frame); returns child rank -------------------
    __device__ int select(const Node& nd, bool is_root) {
        const int nl = popc64(nd.legal);
        const Edge* ed = edges + nd.edge_base;
        long long sum_n = 0;
        for (int i = 0; i < nl; ++i) sum_n += ed[i].n;
        const double xx = fmax(sqrt((double)sum_n), 1.0);
        const bool noisy = is_root && c.noise_eps > 0.f;
        double gsum = 0.0;
        double g[40];
        uint32_t rootsel = 0;
        if (noisy) {
            rootsel = sl.n_rootsel++;
            for (int i = 0; i < nl; ++i) { g[i] = gamma_draw(rootsel, (uint32_t)i); gsum += g[i]; }
        }
        const float keep = (float)(1.0 - (double)c.noise_eps);
        const double eps = (double)c.noise_eps, cp = (double)c.c_puct;
        int best = 0;
        double best_v = -1.0;
        for (int i = 0; i < nl; ++i) {
            const double n = (double)ed[i].n;
            double u;
            if (noisy) {
                const float t32 = keep * ed[i].p;
                const double pr = (double)t32 + eps * (g[i] / gsum);
                u = cp * pr * xx / (1.0 + n);
            } else {
                const float c32 = c.c_puct * ed[i].p;
                u = (double)c32 * xx / (1.0 + n);
            }
            const double q = (double)ed[i].w / (n + 1e-5);
            const double v = q + u + 1000.0;
            if (v > best_v) { best_v = v; best = i; }
        }
        return best;
    }

    __device__ static int nth_set_bit(u64 m, int k) {
        for (int i = 0; i < k; ++i) m &= m - 1;
        return ctz64(m);
    }
    __device__ static int rank_of(u64 legal, int action) { return popc64(legal & ((1ULL << action) - 1)); }

    // ---- one simulation until it terminates / needs an evaluation / parks (oracle SelfPlayGame._run) ----
    // returns 0 done, 1 pending, 2 parked
    __device__ int run(int di) {
        Descent& d = desc[di];
        const int pid = sl.root_pid;
        while (true) {
            const bool black_to_move = d.next_player == 1;
            const u64 own = black_to_move ? d.black : d.white, enemy = black_to_move ? d.white : d.black;
            const u64 legal_here = find_correct_moves(own, enemy);
            if (legal_here == 0) {
                // terminal (the step that led here found no move for either side): player.py:226-232
                const uint8_t w = winner_by_count(d.black, d.white);
                const float v = w == 3 ? 0.f : (w == pid ? 1.f : -1.f);
                backup(d, v);
                return 0;
            }
            const uint32_t kp = kpid_of(pid);
            for (int j = 0; j < sl.n_pending; ++j) {  // player.py:253-254 now_expanding
                const Descent& o = desc[sl.pending[j]];
                if (o.leaf_own == own && o.leaf_enemy == enemy) return 2;
            }
            const int ni = find_node(own, enemy, kp);
            if (ni < 0 || !((nodes[ni].exp >> (pid - 1)) & 1)) {  // leaf for this player, player.py:257
                const U4 r = draw(c.seed, sl.game_id, sl.n_expand, P_DIHEDRAL, 0);
                const int flip = u01(r.x) < 0.5 ? 4 : 0;          // player.py:300
                const int rot = (int)(u01(r.y) * 4.0);             // player.py:301
                sl.n_expand++;
                d.dihedral = (uint8_t)(flip | rot);
                d.leaf_own = own; d.leaf_enemy = enemy;
                d.leaf_mover_is_root = (uint8_t)(d.next_player == pid);
                return 1;
            }
            const Node nd = nodes[ni];
            const int r = select(nd, d.path_len == 0);
            Edge& e = edges[nd.edge_base + r];
            e.n += c.vl;                                   // player.py:270-271
            e.w = e.w - (float)c.vl;
            if (d.path_len >= kMaxPath) { fail(RZ_ECAPACITY); return 0; }
            d.path[d.path_len++] = (nd.edge_base + (uint32_t)r) | ((uint32_t)(d.next_player == pid) << 31);
            // env.step (the move is legal by construction)
            const int a = nth_set_bit(nd.legal, r);
            const u64 fl = calc_flip(a, own, enemy);
            const u64 own2 = own ^ fl | (1ULL << a), en2 = enemy ^ fl;
            d.black = black_to_move ? own2 : en2;
            d.white = black_to_move ? en2 : own2;
            if (find_correct_moves(en2, own2)) d.next_player = black_to_move ? 2 : 1;
            // else: pass (same player) or game over -- resolved at the top of the loop
            else if (!find_correct_moves(own2, en2)) {
                const uint8_t w = winner_by_count(d.black, d.white);
                const float v = w == 3 ? 0.f : (w == pid ? 1.f : -1.f);
                backup(d, v);
                return 0;
            }
        }
    }

    // ---- player.py:283-327 for one evaluated leaf -------------------------------------------------------
    __device__ void consume(int di) {
        Descent& d = desc[di];
        const int pid = sl.root_pid;
        const uint32_t kp = kpid_of(pid);
        int ni = find_node(d.leaf_own, d.leaf_enemy, kp);
        if (ni < 0) ni = create_node(d.leaf_own, d.leaf_enemy, kp);
        if (ni >= 0) {
            Node& nd = nodes[ni];
            const float* pol = p.policy + (size_t)d.leaf_index * 64;
            // inverse dihedral (player.py:315-321) + re-normalisation over legal moves (player.py:406-413),
            // float32, numpy's summation order for 64 contiguous elements (8 column sums, fixed tree)
            float col[8];
            for (int j = 0; j < 8; ++j) col[j] = 0.f;
            for (int sq = 0; sq < 64; ++sq) {
                const float v = ((nd.legal >> sq) & 1ULL) ? pol[dihedral_square(sq, d.dihedral)] : 0.f;
                col[sq & 7] = col[sq & 7] + v;
            }
            const float sum = ((col[0] + col[1]) + (col[2] + col[3])) + ((col[4] + col[5]) + (col[6] + col[7]));
            Edge* ed = edges + nd.edge_base;
            u64 m = nd.legal;
            for (int i = 0; m; ++i, m &= m - 1) {
                const float v = pol[dihedral_square(ctz64(m), d.dihedral)];
                ed[i].p = sum > 0.f ? v / sum : v;
            }
            nd.exp |= (uint8_t)(1u << (pid - 1));
        }
        const float v = p.value[d.leaf_index];
        backup(d, d.leaf_mover_is_root ? v : -v);
        d.status = D_FREE;
    }

    __device__ void start_descent(int di) {
        Descent& d = desc[di];
        const bool black_root = sl.root_pid == 1;
        d.black = black_root ? sl.root_own : sl.root_enemy;
        d.white = black_root ? sl.root_enemy : sl.root_own;
        d.next_player = sl.root_pid;
        d.path_len = 0;
    }

    // ---- oracle SelfPlayGame.search, one wave's worth ---------------------------------------------------
    // returns true if evaluations are pending (the slot must wait for the network)
    __device__ bool search_wave() {
        while (true) {
            uint8_t still[kMaxK];
            int n_still = 0;
            int started_wave = 0;
            sl.n_pending = 0;
            for (int j = 0; j < sl.n_parked; ++j) {
                const int di = sl.parked[j];
                const int r = run(di);
                if (r == 2) still[n_still++] = (uint8_t)di;
                else if (r == 1) { desc[di].status = D_PENDING; sl.pending[sl.n_pending++] = (uint8_t)di; }
                else desc[di].status = D_FREE;
            }
            while (sl.sims_started < sl.sims_target && sl.n_pending + n_still < c.K && started_wave < c.sims_cap) {
                int di = 0;
                ++started_wave;
                while (desc[di].status != D_FREE) ++di;
                sl.sims_started++;
                start_descent(di);
                desc[di].status = D_PARKED;  // reserve while running
                const int r = run(di);
                if (r == 2) still[n_still++] = (uint8_t)di;
                else if (r == 1) { desc[di].status = D_PENDING; sl.pending[sl.n_pending++] = (uint8_t)di; }
                else desc[di].status = D_FREE;
            }
            sl.n_parked = (uint8_t)n_still;
            for (int j = 0; j < n_still; ++j) sl.parked[j] = still[j];
            if (sl.n_pending > 0) return true;
            if (sl.sims_started >= sl.sims_target) return false;
            return true;  // cap reached with nothing to evaluate: continue in the next wave
        }
    }

    __device__ void begin_search(u64 own, u64 enemy, int pid) {
        sl.root_own = own; sl.root_enemy = enemy; sl.root_pid = (uint8_t)pid;
        sl.cur_net = c.two_nets ? (uint8_t)(pid == 1 ? sl.black_net : 1 - sl.black_net) : (uint8_t)0;
        sl.sims_started = 0; sl.sims_target = (uint32_t)c.S;
        sl.n_pending = 0; sl.n_parked = 0;
        sl.phase = PH_SEARCH;
    }

    __device__ void new_game() {
        const u64 local = (u64)s + sl.games_played * (u64)c.G;
        if (c.max_games && local >= c.max_games) {
            sl.phase = PH_IDLE;
            atomicAdd(&p.status->idle_slots, 1ULL);
            return;
        }
        sl.game_id = c.first_game_id + local * c.game_id_stride;
        sl.black_net = c.two_nets ? (uint8_t)(local & 1) : (uint8_t)0;
        sl.games_played++;
        env_reset(sl.env);
        sl.gen = sl.gen + 1;
        if (sl.gen >= 4096) {  // generation tags wrapped: clear this slot's table once
            for (uint32_t i = 0; i < c.hash_cap; ++i) hash[i] = 0;
            sl.gen = 1;
        }
        sl.n_nodes = 0; sl.n_edges = 0; sl.n_expand = 0; sl.n_rootsel = 0; sl.n_sims = 0; sl.ply = 0; sl.tl = 0;
        sl.resigned_mask = 0; sl.search_only = 0;
        for (int k = 0; k < c.K; ++k) desc[k].status = D_FREE;
        sl.enable_resign = (uint8_t)((double)c.disable_resignation_rate <= u01(draw(c.seed, sl.game_id, 0, P_GAME, 0).x));
        atomicAdd(&p.status->games_started, 1ULL);
        sl.phase = PH_DECIDE;  // turn 0: bypass_first_move decides without a search
        if (c.warm_start && sl.games_played == 1) {
            // steady-state start: advance this slot's first game by a random number of random legal plies
            const U4 r0 = draw(c.seed, sl.game_id, 1, P_GAME, 0);
            const int pre = (int)(u01(r0.x) * 58.0);
            for (int i = 0; i < pre && !sl.env.done; ++i) {
                const bool b = sl.env.next_player == 1;
                const u64 legal = find_correct_moves(b ? sl.env.black : sl.env.white, b ? sl.env.white : sl.env.black);
                const U4 r = draw(c.seed, sl.game_id, 2 + (uint32_t)i, P_GAME, 0);
                env_step(sl.env, nth_set_bit(legal, (int)(u01(r.x) * (double)popc64(legal))));
            }
            if (sl.env.done) env_reset(sl.env);
            else if (sl.env.turn > 0) {
                const bool b = sl.env.next_player == 1;
                begin_search(b ? sl.env.black : sl.env.white, b ? sl.env.white : sl.env.black, sl.env.next_player);
            }
        }
    }

    __device__ void finish_game() {
        rz_game& g = p.mail_hdr[(size_t)s * 2 + sl.log_sel];
        g.game_id = sl.game_id; g.black = sl.env.black; g.white = sl.env.white;
        g.first_ply = 0; g.n_plies = (int32_t)sl.ply; g.expansions = (int32_t)sl.n_expand; g.simulations = (int32_t)sl.n_sims;
        g.winner = sl.env.winner; g.black_z = sl.env.winner == 1 ? 1 : (sl.env.winner == 2 ? -1 : 0);
        g.resign_enabled = sl.enable_resign; g.resigned_mask = sl.resigned_mask; g.turn = sl.env.turn;
        g.black_net = sl.black_net; g.pad[0] = g.pad[1] = 0;
        atomicMax(&p.status->max_nodes, (unsigned long long)sl.n_nodes);
        atomicMax(&p.status->max_edges, (unsigned long long)sl.n_edges);
        __threadfence();
        p.mail_flag[(size_t)s * 2 + sl.log_sel] = 1;
        atomicAdd(&p.status->games_finished, 1ULL);
        sl.log_sel ^= 1;
        sl.phase = PH_NEWGAME;
    }

    // ---- player.py:82-134 action_with_evaluation (solver disabled) + self_play.py:155-162 ---------------
    __device__ void decide() {
        const bool black_to_move = sl.env.next_player == 1;
        const int pid = sl.env.next_player;
        const u64 own = black_to_move ? sl.env.black : sl.env.white, enemy = black_to_move ? sl.env.white : sl.env.black;
        const int turn = popc64(own) + popc64(enemy) - 4;
        const uint32_t kp = kpid_of(pid);
        int ni = find_node(own, enemy, kp);
        if (turn == 0) {  // bypass_first_move, player.py:143-148
            if (ni < 0) ni = create_node(own, enemy, kp);
            if (ni < 0) { sl.phase = PH_IDLE; return; }
            const Node& nd0 = nodes[ni];
            const int nl0 = popc64(nd0.legal);
            Edge* ed0 = edges + nd0.edge_base;
            ed0[0].n = 1; ed0[0].w = 0.f;
            for (int i = 0; i < nl0; ++i) ed0[i].p = 1.0f / (float)nl0;  // legal / sum(legal), renormalised (== itself)
        }
        if (ni < 0) { fail(RZ_ESTATE); sl.phase = PH_IDLE; return; }
        const Node nd = nodes[ni];
        const int nl = popc64(nd.legal);
        const Edge* ed = edges + nd.edge_base;
        long long sum_n = 0;
        int arg_n = 0;
        for (int i = 0; i < nl; ++i) { sum_n += ed[i].n; if (ed[i].n > ed[arg_n].n) arg_n = i; }
        // calc_policy, player.py:366-385
        const bool tau1 = turn < c.change_tau_turn;
        const U4 r = draw(c.seed, sl.game_id, sl.ply * 16 + sl.tl, P_MOVE, 0);
        const double uu = u53(r.x, r.y);
        int choice = arg_n;
        if (tau1) {  // np.random.choice(range(64), p = N / sum N): cumsum, normalise, searchsorted(side='right')
            double total = 0.0;
            for (int i = 0; i < nl; ++i) total += (double)ed[i].n / (double)sum_n;
            double cum = 0.0;
            choice = nl - 1;
            for (int i = 0; i < nl; ++i) {
                cum += (double)ed[i].n / (double)sum_n;
                if (cum / total > uu) { choice = i; break; }
            }
        }
        // rethinking rule, player.py:113-118
        int abv = -1;
        double q_abv = 0.0;
        double max_q_visited = -10.0;
        for (int i = 0; i < nl; ++i) {
            if (ed[i].n > 0) {
                const double q = (double)ed[i].w / ((double)ed[i].n + 1e-5);
                if (abv < 0 || q + 100.0 > q_abv + 100.0) { abv = i; q_abv = q; }
                if (q > max_q_visited) max_q_visited = q;
            }
        }
        const double q_choice = (double)ed[choice].w / ((double)ed[choice].n + 1e-5);
        const double value_diff = q_choice - q_abv;
        sl.tl++;
        const bool accept = turn <= c.start_rethinking_turn || (value_diff > -0.01 && ed[choice].n >= c.required_visit);
        // (rethinking is skipped when the node arena could not also hold one search for every remaining ply)
        const bool room = (u64)sl.n_nodes + (u64)c.S * (u64)(61 - turn) + 64 <= (u64)c.nodes_cap;
        if (!accept && sl.tl < c.thinking_loop && turn > 0 && room) {  // think again: another simulation_num_per_move
            begin_search(own, enemy, pid);
            return;
        }
        const int action = nth_set_bit(nd.legal, choice);
        // log the ply
        if ((int)sl.ply >= c.max_plies) { fail(RZ_ECAPACITY); sl.phase = PH_IDLE; return; }
        rz_ply& pl = p.plies[((size_t)s * 2 + sl.log_sel) * c.max_plies + sl.ply];
        pl.own = own; pl.enemy = enemy;
        for (int i = 0; i < 64; ++i) pl.n_visit[i] = 0;
        { u64 m = nd.legal; for (int i = 0; m; ++i, m &= m - 1) pl.n_visit[ctz64(m)] = ed[i].n; }
        pl.player = (uint8_t)pid; pl.loops = sl.tl; pl.pad[0] = pl.pad[1] = pl.pad[2] = 0;
        pl.n = (float)ed[choice].n; pl.q = (float)q_choice;
        bool resign = false;
        if (c.use_resign && max_q_visited <= (double)c.resign_threshold) {  // player.py:123-130
            sl.resigned_mask |= (uint8_t)(1u << (pid - 1));
            if (sl.enable_resign && turn >= c.allowed_resign_turn) resign = true;
        }
        pl.action = resign ? (int16_t)-1 : (int16_t)action;
        pl.recorded = resign ? 0 : 1;
        sl.ply++;
        sl.tl = 0;
        atomicAdd(&p.status->plies, 1ULL);
        env_step(sl.env, resign ? -1 : action);  // self_play.py:162
        if (sl.env.done) { finish_game(); return; }
        const bool b2 = sl.env.next_player == 1;
        begin_search(b2 ? sl.env.black : sl.env.white, b2 ? sl.env.white : sl.env.black, sl.env.next_player);
    }
};

// ---- the per-wave kernel: one thread per game slot ------------------------------------------------------
constexpr int kTickThreads = 64;

__global__ void __launch_bounds__(kTickThreads) tick_kernel(const DevCfg c, const DevPtrs p, const int slot0, const int slot_end,
                                                            const int group) {
    const int s = slot0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int n_leaves = 0;
    if (s < slot_end) {
        Ctx x(c, p, s);
        Slot& sl = x.sl;
        if (p.status->error == 0 && sl.phase != PH_IDLE) {
            // 1. consume the evaluations requested in the previous wave, in request order
            if (sl.phase == PH_SEARCH) {
                uint32_t consumed = sl.n_pending;
                for (int j = 0; j < sl.n_pending; ++j) x.consume(sl.pending[j]);
                sl.n_pending = 0;
                if (consumed) atomicAdd(&p.status->expansions, (unsigned long long)consumed);
            }
            // 2. advance the slot's state machine until it needs the network again
            for (int guard = 0; guard < 100000; ++guard) {
                if (p.status->error != 0) break;
                if (sl.phase == PH_SEARCH) {
                    if (x.search_wave()) break;
                    sl.n_sims += sl.sims_started;
                    atomicAdd(&p.status->simulations, (unsigned long long)sl.sims_started);
                    if (sl.search_only) { sl.phase = PH_IDLE; atomicAdd(&p.status->idle_slots, 1ULL); break; }
                    sl.phase = PH_DECIDE;
                } else if (sl.phase == PH_DECIDE) {
                    x.decide();
                } else if (sl.phase == PH_NEWGAME) {
                    // the other ply log must have been harvested by the host before it is reused
                    if (p.mail_flag[(size_t)s * 2 + sl.log_sel]) break;
                    x.new_game();
                } else {
                    break;
                }
            }
            if (sl.phase == PH_SEARCH) n_leaves = sl.n_pending;
        }
    }
    // 3. gather: warp-scan the per-slot leaf counts, one atomic per warp, coalesced-ish batch writes
    int incl = n_leaves;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (!c.two_nets && lane == 31 && total > 0) base = atomicAdd(p.batch_count + group * 64, (uint32_t)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (n_leaves > 0) {
        Slot& sl = p.slots[s];
        Descent* desc = p.desc + (size_t)s * c.K;
        // each (network, group) owns the batch rows [net * G * K + slot0 * K, ...); leaf_index is the absolute row
        uint32_t at = (uint32_t)slot0 * (uint32_t)c.K + base + (uint32_t)(incl - n_leaves);
        if (c.two_nets)  // slots of one warp may belong to different networks: one atomic per slot
            at = (uint32_t)sl.cur_net * (uint32_t)c.G * (uint32_t)c.K + (uint32_t)slot0 * (uint32_t)c.K +
                 atomicAdd(p.batch_count + (sl.cur_net * 2 + group) * 64, (uint32_t)n_leaves);
        for (int j = 0; j < n_leaves; ++j, ++at) {
            Descent& d = desc[sl.pending[j]];
            d.leaf_index = at;
            p.batch_own[at] = dihedral(d.leaf_own, d.dihedral);      // K3: NN input already transformed
            p.batch_enemy[at] = dihedral(d.leaf_enemy, d.dihedral);
        }
    }
}

#include "rz_engine_warp.cuh"

// RZ_EVAL_FAKE: policy 1/64, value (#own - #enemy)/64 (oracle/nn.py FakeNetAPI)
__global__ void fake_eval_kernel(const u64* __restrict__ own, const u64* __restrict__ enemy, const uint32_t* __restrict__ count,
                                 float* __restrict__ policy, float* __restrict__ value, float sign) {
    const uint32_t n = *count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n * 64; i += gridDim.x * blockDim.x) {
        policy[i] = 1.0f / 64.0f;
        if ((i & 63) == 0) value[i >> 6] = sign * ((float)(popc64(own[i >> 6]) - popc64(enemy[i >> 6])) / 64.0f);
    }
}

__global__ void init_slots_kernel(const DevCfg c, const DevPtrs p) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= c.G) return;
    Slot& sl = p.slots[s];
    memset(&sl, 0, sizeof(Slot));
    sl.phase = PH_NEWGAME;
    p.mail_flag[(size_t)s * 2] = 0; p.mail_flag[(size_t)s * 2 + 1] = 0;
}

// test hook: every slot searches the same root once (no game loop)
__global__ void setup_search_root_kernel(const DevCfg c, const DevPtrs p, u64 own, u64 enemy, int pid, int keep_tree) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= c.G) return;
    Ctx x(c, p, s);
    Slot& sl = x.sl;
    sl.game_id = c.first_game_id + (u64)s * c.game_id_stride;
    if (!keep_tree || sl.gen == 0) {
        sl.gen = sl.gen + 1;
        if (sl.gen >= 4096) { for (uint32_t i = 0; i < c.hash_cap; ++i) x.hash[i] = 0; sl.gen = 1; }
        sl.n_nodes = 0; sl.n_edges = 0; sl.n_expand = 0; sl.n_rootsel = 0; sl.n_sims = 0;
    }
    sl.ply = 0; sl.tl = 0;
    for (int k = 0; k < c.K; ++k) x.desc[k].status = D_FREE;
    env_update(sl.env, pid == 1 ? own : enemy, pid == 1 ? enemy : own, pid);
    x.begin_search(own, enemy, pid);
    sl.search_only = 1;
}
__global__ void read_root_kernel(const DevCfg c, const DevPtrs p, int s, int32_t* n_out, float* w_out) {
    Ctx x(c, p, s);
    for (int i = 0; i < 64; ++i) { n_out[i] = 0; w_out[i] = 0.f; }
    const int ni = x.find_node(x.sl.root_own, x.sl.root_enemy, x.kpid_of(x.sl.root_pid));
    if (ni < 0) return;
    const Node& nd = x.nodes[ni];
    u64 m = nd.legal;
    for (int i = 0; m; ++i, m &= m - 1) { n_out[ctz64(m)] = x.edges[nd.edge_base + i].n; w_out[ctz64(m)] = x.edges[nd.edge_base + i].w; }
}

}  // namespace eng
}  // namespace rz

using namespace rz;
using namespace rz::eng;

struct FinishedGame {
    rz_game hdr;
    std::vector<rz_ply> plies;
};

struct rz_engine {
    rz_engine_cfg cfg;
    DevCfg dc;
    DevPtrs dp;
    rz_net* net;
    rz_net* net_b;  // evaluation matches: the second network (NULL in self-play)
    int device;
    cudaStream_t stream;      // group 0 + all host<->device traffic
    cudaStream_t stream2;     // group 1 (tick of one group overlaps the network launch of the other)
    int n_groups;
    int group_slot0[3];
    size_t solver_tt_words_per_group;
    int solve_parity[2];      // per group: which of its two unfinished-solve lists the next wave reads
    long long solver_budget_ns;  // time the solver step may take per wave and group (RZ_SOLVER_BUDGET_US)
    int tick_impl;            // 0 = warp-per-game kernel (default), 1 = thread-per-slot cross-check (RZ_TICK_IMPL=thread)
    void* arena[32];
    int n_arena;
    Status* h_status;     // pinned
    uint8_t* h_flags;     // pinned [G*2]
    uint64_t waves, nn_launches, mcts_launches;
    uint64_t finished_total;
    std::deque<FinishedGame> queue;
    // device timing: 3 events per queued wave (before tick, between tick and evaluation, after evaluation)
    cudaEvent_t ev[2 * 3 * 8];  // [group][wave in burst][3]
    cudaEvent_t ev_run[3];     // run start, run end, group-1 join
    int ev_used;
    double nn_ms, mcts_ms, run_ms;
};

static int collect_timing(rz_engine* e) {  // call after the stream has been synchronised
    for (int g = 0; g < e->n_groups; ++g)
        for (int i = 0; i < e->ev_used; ++i) {
            float a = 0.f, b = 0.f;
            cudaEvent_t* ev = e->ev + (g * 8 + i) * 3;
            RZ_CUDA_TRY(cudaEventElapsedTime(&a, ev[0], ev[1]));
            RZ_CUDA_TRY(cudaEventElapsedTime(&b, ev[1], ev[2]));
            e->mcts_ms += a; e->nn_ms += b;
        }
    e->ev_used = 0;
    return RZ_OK;
}

static int dev_alloc(rz_engine* e, void** ptr, size_t bytes, bool zero) {
    if (e->n_arena >= (int)(sizeof(e->arena) / sizeof(e->arena[0]))) { set_error("rz_engine: allocation table full"); return RZ_ESTATE; }
    cudaError_t ce = cudaMalloc(ptr, bytes);
    if (ce != cudaSuccess) {
        set_error("rz_engine: cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(ce));
        cudaGetLastError();
        return RZ_ENOMEM;
    }
    e->arena[e->n_arena++] = *ptr;
    if (zero) RZ_CUDA_TRY(cudaMemsetAsync(*ptr, 0, bytes, e->stream));
    return RZ_OK;
}

static int drain_mailboxes(rz_engine* e) {
    const int G = e->dc.G;
    RZ_CUDA_TRY(cudaMemcpyAsync(e->h_flags, e->dp.mail_flag, (size_t)G * 2, cudaMemcpyDeviceToHost, e->stream));
    RZ_CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < G * 2; ++i) {
        if (!e->h_flags[i]) continue;
        FinishedGame fg;
        RZ_CUDA_TRY(cudaMemcpyAsync(&fg.hdr, e->dp.mail_hdr + i, sizeof(rz_game), cudaMemcpyDeviceToHost, e->stream));
        RZ_CUDA_TRY(cudaStreamSynchronize(e->stream));
        const int np = fg.hdr.n_plies;
        if (np < 0 || np > e->dc.max_plies) { set_error("rz_engine: corrupt mailbox (n_plies=%d)", np); return RZ_ESTATE; }
        fg.plies.resize((size_t)np);
        if (np) RZ_CUDA_TRY(cudaMemcpyAsync(fg.plies.data(), e->dp.plies + (size_t)i * e->dc.max_plies, (size_t)np * sizeof(rz_ply),
                                            cudaMemcpyDeviceToHost, e->stream));
        RZ_CUDA_TRY(cudaMemsetAsync(e->dp.mail_flag + i, 0, 1, e->stream));
        RZ_CUDA_TRY(cudaStreamSynchronize(e->stream));
        e->queue.push_back(std::move(fg));
        e->finished_total++;
    }
    return RZ_OK;
}

static int launch_wave(rz_engine* e) {
    const DevCfg& c = e->dc;
    const bool timed = e->ev_used < 8;
    for (int g = 0; g < e->n_groups; ++g) {
        cudaStream_t st = g == 0 ? e->stream : e->stream2;
        const int s0 = e->group_slot0[g], s1 = e->group_slot0[g + 1];
        cudaEvent_t* ev = e->ev + (g * 8 + e->ev_used) * 3;
        const size_t rows = (size_t)(s1 - s0) * c.K;
        if (timed) RZ_CUDA_TRY(cudaEventRecord(ev[0], st));
        for (int net = 0; net <= c.two_nets; ++net) RZ_CUDA_TRY(cudaMemsetAsync(e->dp.batch_count + (net * 2 + g) * 64, 0, sizeof(uint32_t), st));
        const bool solving = c.solver_turn > 0 || c.solver_sim_turn > 0;
        const int par = e->solve_parity[g];  // the unfinished-solve list this wave's tick appends to
        if (solving) RZ_CUDA_TRY(cudaMemsetAsync(e->dp.solve_count + (g * 2 + (1 - par)) * 64, 0, sizeof(uint32_t), st));
        if (e->tick_impl == 1)
            tick_kernel<<<(s1 - s0 + kTickThreads - 1) / kTickThreads, kTickThreads, 0, st>>>(c, e->dp, s0, s1, g);
        else
            tick_warp_kernel<<<(s1 - s0 + 1) / 2, kWarpTickThreads, 0, st>>>(c, e->dp, s0, s1, g, par);
        RZ_LAUNCH_CHECK();
        e->mcts_launches++;
        if (solving) {  // advance the group's unfinished solves for a bounded time (underneath the other group's tower)
            const size_t list = (size_t)c.G * (c.K + 1);
            solver::solve_active_kernel<<<num_sms(), solver::kBlockThreads, 0, st>>>(
                e->dp.sctx, e->dp.sactive + (g * 2 + par) * list, e->dp.solve_count + (g * 2 + par) * 64,
                e->dp.sactive + (g * 2 + (1 - par)) * list, e->dp.solve_count + (g * 2 + (1 - par)) * 64,
                e->dp.solver_tt + (size_t)g * e->solver_tt_words_per_group, e->solver_budget_ns);
            RZ_LAUNCH_CHECK();
            e->mcts_launches++;
            e->solve_parity[g] = 1 - par;
        }
        if (timed) RZ_CUDA_TRY(cudaEventRecord(ev[1], st));
        for (int net = 0; net <= c.two_nets; ++net) {
            uint32_t* count = e->dp.batch_count + (net * 2 + g) * 64;
            const size_t row0 = (size_t)net * c.G * c.K + (size_t)s0 * c.K;
            if (e->cfg.eval_mode == RZ_EVAL_FAKE) {
                fake_eval_kernel<<<num_sms() * 4, 256, 0, st>>>(e->dp.batch_own + row0, e->dp.batch_enemy + row0, count, e->dp.policy + row0 * 64,
                                                              e->dp.value + row0, net ? -1.f : 1.f);
                RZ_LAUNCH_CHECK();
                e->mcts_launches++;
            } else {
                RZ_TRY(net_forward_counted(net ? e->net_b : e->net, e->dp.batch_own + row0, e->dp.batch_enemy + row0, e->dp.policy + row0 * 64,
                                           e->dp.value + row0, count, rows, e->cfg.net_impl, st));
                e->nn_launches++;
            }
        }
        if (timed) RZ_CUDA_TRY(cudaEventRecord(ev[2], st));
    }
    if (timed) e->ev_used++;
    e->waves++;
    return RZ_OK;
}

static int sync_all(rz_engine* e) {
    if (e->n_groups > 1) RZ_CUDA_TRY(cudaStreamSynchronize(e->stream2));
    RZ_CUDA_TRY(cudaStreamSynchronize(e->stream));
    return RZ_OK;
}

static int read_status(rz_engine* e) {
    RZ_TRY(sync_all(e));
    RZ_CUDA_TRY(cudaMemcpyAsync(e->h_status, e->dp.status, sizeof(Status), cudaMemcpyDeviceToHost, e->stream));
    RZ_CUDA_TRY(cudaStreamSynchronize(e->stream));
    RZ_TRY(collect_timing(e));
    if (e->h_status->error != 0) {
        set_error("rz_engine: device-side failure %d (%s)", e->h_status->error,
                  e->h_status->error == RZ_ECAPACITY ? "node/edge/ply arena overflow" : "inconsistent search state");
        return e->h_status->error;
    }
    return RZ_OK;
}

extern "C" {

int rz_engine_create(const rz_engine_cfg* cfg, rz_net* net, int device, rz_engine** out) {
    RZ_REQUIRE(cfg && out, "rz_engine_create: null pointer");
    RZ_REQUIRE(cfg->games >= 1 && cfg->games <= (1 << 20), "games out of range (%d)", cfg->games);
    RZ_REQUIRE(cfg->simulation_num_per_move >= 1 && cfg->simulation_num_per_move <= 100000, "simulation_num_per_move out of range");
    RZ_REQUIRE(cfg->parallel_search_num >= 1 && cfg->parallel_search_num <= kMaxK, "parallel_search_num must be 1..%d", kMaxK);
    RZ_REQUIRE(cfg->thinking_loop >= 1 && cfg->thinking_loop <= 255, "thinking_loop must be 1..255");
    RZ_REQUIRE(cfg->eval_mode == RZ_EVAL_FAKE || net, "rz_engine_create: a network is required unless eval_mode == RZ_EVAL_FAKE");
    RZ_REQUIRE(cfg->game_id_stride >= 1, "game_id_stride must be >= 1");
    RZ_CUDA_TRY(cudaSetDevice(device));
    rz_engine* e = new (std::nothrow) rz_engine();
    if (!e) { set_error("out of host memory"); return RZ_ENOMEM; }
    e->cfg = *cfg; e->net = net; e->net_b = nullptr; e->device = device; e->n_arena = 0;
    e->waves = e->nn_launches = e->mcts_launches = e->finished_total = 0;
    e->h_status = nullptr; e->h_flags = nullptr; e->stream = nullptr; e->stream2 = nullptr;
    e->ev_used = 0; e->nn_ms = e->mcts_ms = e->run_ms = 0.0;
    for (int i = 0; i < 48; ++i) e->ev[i] = nullptr;
    e->ev_run[0] = e->ev_run[1] = e->ev_run[2] = nullptr;
    // two slot groups on two streams: the MCTS tick of one group runs underneath the network launch of the other
    e->n_groups = cfg->overlap_groups == 1 ? 1 : (cfg->overlap_groups == 2 ? 2 : (cfg->games >= 256 ? 2 : 1));
    if (cfg->games < 2) e->n_groups = 1;
    e->group_slot0[0] = 0;
    e->group_slot0[1] = e->n_groups == 2 ? (cfg->games + 1) / 2 : cfg->games;
    e->group_slot0[2] = cfg->games;
    const char* ti = getenv("RZ_TICK_IMPL");
    e->tick_impl = (ti && strcmp(ti, "thread") == 0) ? 1 : 0;
    if (e->tick_impl == 1 && (cfg->use_solver_turn > 0 || cfg->use_solver_turn_in_simulation > 0 || cfg->reset_mtcs_info_per_game > 1)) {
        set_error("the thread-per-slot cross-check kernel implements neither the endgame solver hooks nor reset_mtcs_info_per_game > 1");
        delete e;
        return RZ_EINVAL;
    }
    DevCfg& c = e->dc;
    c.G = cfg->games; c.S = cfg->simulation_num_per_move; c.K = cfg->parallel_search_num; c.vl = cfg->virtual_loss;
    c.change_tau_turn = cfg->change_tau_turn; c.thinking_loop = cfg->thinking_loop; c.required_visit = cfg->required_visit_to_decide_action;
    c.start_rethinking_turn = cfg->start_rethinking_turn; c.allowed_resign_turn = cfg->allowed_resign_turn;
    c.use_resign = cfg->use_resign_threshold; c.share = cfg->share_mtcs_info; c.max_plies = cfg->max_plies > 0 ? cfg->max_plies : 64;
    c.warm_start = cfg->warm_start;
    c.two_nets = 0;
    c.keep_games = cfg->reset_mtcs_info_per_game > 1 ? cfg->reset_mtcs_info_per_game : 1;
    c.solver_turn = cfg->use_solver_turn; c.solver_sim_turn = cfg->use_solver_turn_in_simulation;
    c.sims_cap = cfg->max_sims_per_wave > 0 ? cfg->max_sims_per_wave : 2 * cfg->parallel_search_num;
    c.c_puct = cfg->c_puct; c.noise_eps = cfg->noise_eps; c.alpha = cfg->dirichlet_alpha; c.resign_threshold = cfg->resign_threshold;
    c.disable_resignation_rate = cfg->disable_resignation_rate;
    c.seed = cfg->seed; c.first_game_id = cfg->first_game_id; c.game_id_stride = cfg->game_id_stride; c.max_games = cfg->max_games;
    // every simulation creates at most one node; a game has at most 60 searched plies
    const uint64_t searches = cfg->max_searches_per_game > 0 ? (uint64_t)cfg->max_searches_per_game
                                                              : (uint64_t)60 * (c.thinking_loop > 2 ? 2 : c.thinking_loop);
    uint64_t nodes = searches * c.S * (uint64_t)c.keep_games + 64;
    if (nodes > 0xFFFF0) nodes = 0xFFFF0;
    c.nodes_cap = (uint32_t)nodes;
    c.edges_cap = c.nodes_cap * 14;
    uint32_t h = 1024;
    while (h < 2 * c.nodes_cap) h <<= 1;
    c.hash_cap = h;
    int rc = RZ_OK;
    cudaError_t ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking);
    if (ce != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(ce)); delete e; return RZ_ECUDA; }
    const size_t G = c.G, B = 2 * G * c.K;  // rows for two networks (evaluation matches); self-play uses the first half
    DevPtrs& p = e->dp;
    rc = dev_alloc(e, (void**)&p.slots, G * sizeof(Slot), true);
    if (!rc) rc = dev_alloc(e, (void**)&p.desc, G * c.K * sizeof(Descent), true);
    if (!rc) rc = dev_alloc(e, (void**)&p.hash, G * c.hash_cap * sizeof(uint32_t), true);
    if (!rc) rc = dev_alloc(e, (void**)&p.nodes, G * c.nodes_cap * sizeof(Node), false);
    if (!rc) rc = dev_alloc(e, (void**)&p.edges, G * (size_t)c.edges_cap * sizeof(Edge), false);
    if (!rc) rc = dev_alloc(e, (void**)&p.plies, G * 2 * c.max_plies * sizeof(rz_ply), false);
    if (!rc) rc = dev_alloc(e, (void**)&p.mail_hdr, G * 2 * sizeof(rz_game), true);
    if (!rc) rc = dev_alloc(e, (void**)&p.mail_flag, G * 2, true);
    if (!rc) rc = dev_alloc(e, (void**)&p.status, sizeof(Status), true);
    if (!rc) rc = dev_alloc(e, (void**)&p.batch_count, 1024, true);
    if (!rc) rc = dev_alloc(e, (void**)&p.batch_own, (B + 2) * sizeof(u64), true);
    if (!rc) rc = dev_alloc(e, (void**)&p.batch_enemy, (B + 2) * sizeof(u64), true);
    if (!rc) rc = dev_alloc(e, (void**)&p.policy, (B + 2) * 64 * sizeof(float), true);
    if (!rc) rc = dev_alloc(e, (void**)&p.value, (B + 2) * sizeof(float), true);
    const bool solving = c.solver_turn > 0 || c.solver_sim_turn > 0;
    const size_t SR = G * (c.K + 1);
    if (!rc) rc = dev_alloc(e, (void**)&p.solve_count, 1024, true);
    p.sctx = nullptr; p.sactive = nullptr; p.keep_policy = nullptr; p.keep_value = nullptr; p.solver_tt = nullptr;
    e->solve_parity[0] = e->solve_parity[1] = 0;
    {
        const char* b = getenv("RZ_SOLVER_BUDGET_US");
        const long long us = b ? atoll(b) : 2000;
        e->solver_budget_ns = (us > 0 ? us : 2000) * 1000LL;
    }
    // one transposition table per lane of the solver step's grid (one CTA per SM), one set per slot group
    e->solver_tt_words_per_group = (size_t)num_sms() * solver::kBlockThreads * solver::kTtEntries * solver::kTtWordsPerEntry;
    if (solving) {
        if (!rc) rc = dev_alloc(e, (void**)&p.sctx, SR * sizeof(solver::SolveCtx), true);
        if (!rc) rc = dev_alloc(e, (void**)&p.sactive, 4 * SR * sizeof(uint32_t), true);
        if (!rc) rc = dev_alloc(e, (void**)&p.keep_policy, G * c.K * 64 * sizeof(float), true);
        if (!rc) rc = dev_alloc(e, (void**)&p.keep_value, G * c.K * sizeof(float), true);
        if (!rc) rc = dev_alloc(e, (void**)&p.solver_tt, e->solver_tt_words_per_group * 2 * sizeof(u64), true);
    }
    if (!rc && cudaMallocHost((void**)&e->h_status, sizeof(Status)) != cudaSuccess) { set_error("cudaMallocHost failed"); rc = RZ_ENOMEM; }
    if (!rc && cudaMallocHost((void**)&e->h_flags, G * 2) != cudaSuccess) { set_error("cudaMallocHost failed"); rc = RZ_ENOMEM; }
    for (int i = 0; i < 48 && !rc; ++i)
        if (cudaEventCreate(&e->ev[i]) != cudaSuccess) { set_error("cudaEventCreate failed"); rc = RZ_ECUDA; }
    for (int i = 0; i < 3 && !rc; ++i)
        if (cudaEventCreate(&e->ev_run[i]) != cudaSuccess) { set_error("cudaEventCreate failed"); rc = RZ_ECUDA; }
    if (rc) { rz_engine_destroy(e); return rc; }
    init_slots_kernel<<<(c.G + 127) / 128, 128, 0, e->stream>>>(c, p);
    ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) { set_error("rz_engine_create: %s", cudaGetErrorString(ce)); rz_engine_destroy(e); return RZ_ECUDA; }
    *out = e;
    return RZ_OK;
}

int rz_engine_destroy(rz_engine* e) {
    if (!e) return RZ_OK;
    cudaSetDevice(e->device);
    if (e->stream2) cudaStreamSynchronize(e->stream2);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (int i = 0; i < e->n_arena; ++i) cudaFree(e->arena[i]);
    if (e->h_status) cudaFreeHost(e->h_status);
    if (e->h_flags) cudaFreeHost(e->h_flags);
    for (int i = 0; i < 48; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    for (int i = 0; i < 3; ++i) if (e->ev_run[i]) cudaEventDestroy(e->ev_run[i]);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->stream2) cudaStreamDestroy(e->stream2);
    cudaGetLastError();
    delete e;
    return RZ_OK;
}

int rz_engine_run(rz_engine* e, uint64_t finished_target, uint64_t max_waves) {
    RZ_REQUIRE(e, "rz_engine_run: null engine");
    RZ_CUDA_TRY(cudaSetDevice(e->device));
    const uint64_t wave0 = e->waves;
    const int kCheck = 8;  // waves queued between host checks
    RZ_TRY(sync_all(e));
    RZ_CUDA_TRY(cudaEventRecord(e->ev_run[0], e->stream));
    if (e->n_groups > 1) RZ_CUDA_TRY(cudaStreamWaitEvent(e->stream2, e->ev_run[0], 0));
    int rc_loop = RZ_OK;
    while (true) {
        if (e->finished_total >= finished_target && finished_target > 0) break;
        if (max_waves && e->waves - wave0 >= max_waves) break;
        int burst = kCheck;
        if (max_waves && e->waves - wave0 + burst > max_waves) burst = (int)(max_waves - (e->waves - wave0));
        for (int i = 0; i < burst; ++i) RZ_TRY(launch_wave(e));
        RZ_TRY(read_status(e));
        if (e->h_status->games_finished > e->finished_total) RZ_TRY(drain_mailboxes(e));
        if (e->h_status->idle_slots >= (unsigned long long)e->dc.G) {  // every slot ran out of games
            RZ_TRY(drain_mailboxes(e));
            break;
        }
    }
    (void)rc_loop;
    if (e->n_groups > 1) {
        RZ_CUDA_TRY(cudaEventRecord(e->ev_run[2], e->stream2));
        RZ_CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_run[2], 0));
    }
    RZ_CUDA_TRY(cudaEventRecord(e->ev_run[1], e->stream));
    RZ_TRY(sync_all(e));
    float ms = 0.f;
    RZ_CUDA_TRY(cudaEventElapsedTime(&ms, e->ev_run[0], e->ev_run[1]));
    e->run_ms += ms;
    return RZ_OK;
}

int rz_engine_poll(rz_engine* e, rz_game* games, size_t game_cap, size_t* n_games, rz_ply* plies, size_t ply_cap, size_t* n_plies) {
    RZ_REQUIRE(e && n_games && n_plies, "rz_engine_poll: null pointer");
    size_t ng = 0, np = 0;
    while (!e->queue.empty() && ng < game_cap) {
        FinishedGame& fg = e->queue.front();
        if (np + fg.plies.size() > ply_cap) break;
        games[ng] = fg.hdr;
        games[ng].first_ply = (int32_t)np;
        if (!fg.plies.empty()) memcpy(plies + np, fg.plies.data(), fg.plies.size() * sizeof(rz_ply));
        np += fg.plies.size();
        ++ng;
        e->queue.pop_front();
    }
    *n_games = ng; *n_plies = np;
    return RZ_OK;
}

int rz_engine_stats(rz_engine* e, rz_stats* out) {
    RZ_REQUIRE(e && out, "rz_engine_stats: null pointer");
    RZ_CUDA_TRY(cudaSetDevice(e->device));
    RZ_TRY(sync_all(e));
    RZ_CUDA_TRY(cudaMemcpyAsync(e->h_status, e->dp.status, sizeof(Status), cudaMemcpyDeviceToHost, e->stream));
    RZ_CUDA_TRY(cudaStreamSynchronize(e->stream));
    const Status& s = *e->h_status;
    out->games_started = s.games_started; out->games_finished = s.games_finished; out->expansions = s.expansions;
    out->simulations = s.simulations; out->waves = e->waves; out->plies = s.plies; out->nn_launches = e->nn_launches;
    out->mcts_launches = e->mcts_launches; out->max_nodes_used = s.max_nodes; out->max_edges_used = s.max_edges;
    RZ_TRY(collect_timing(e));
    out->nn_ms = e->nn_ms; out->mcts_ms = e->mcts_ms; out->run_ms = e->run_ms;
    return RZ_OK;
}

int rz_engine_set_simulation_num(rz_engine* e, int32_t sims) {
    RZ_REQUIRE(e && sims >= 1, "rz_engine_set_simulation_num: bad argument");
    const uint64_t searches = e->cfg.max_searches_per_game > 0 ? (uint64_t)e->cfg.max_searches_per_game
                                                                  : (uint64_t)60 * (e->dc.thinking_loop > 2 ? 2 : e->dc.thinking_loop);
    uint64_t need = searches * sims * (uint64_t)e->dc.keep_games + 64;
    RZ_REQUIRE(need <= e->dc.nodes_cap, "simulation count %d exceeds the arenas sized at creation", sims);
    e->dc.S = sims;
    e->cfg.simulation_num_per_move = sims;
    return RZ_OK;
}

int rz_engine_set_second_net(rz_engine* e, rz_net* net_b, int enable) {
    RZ_REQUIRE(e, "rz_engine_set_second_net: null engine");
    RZ_REQUIRE(!enable || e->cfg.eval_mode == RZ_EVAL_FAKE || net_b, "rz_engine_set_second_net: a second network is required");
    RZ_REQUIRE(e->waves == 0, "rz_engine_set_second_net: must be called before the first wave");
    e->net_b = enable ? net_b : nullptr;
    e->dc.two_nets = enable ? 1 : 0;
    return RZ_OK;
}

int rz_engine_set_resign_threshold(rz_engine* e, int use_resign_threshold, float resign_threshold) {
    RZ_REQUIRE(e, "rz_engine_set_resign_threshold: null engine");
    e->dc.use_resign = use_resign_threshold ? 1 : 0;
    e->dc.resign_threshold = resign_threshold;
    e->cfg.use_resign_threshold = e->dc.use_resign;
    e->cfg.resign_threshold = resign_threshold;
    return RZ_OK;
}

int rz_engine_search_root(rz_engine* e, uint64_t own, uint64_t enemy, int player, int slot, int keep_tree, int32_t* n_visit,
                          float* w_sum) {
    RZ_REQUIRE(e && n_visit && w_sum && (player == 1 || player == 2) && slot >= 0 && slot < e->dc.G, "rz_engine_search_root: bad argument");
    RZ_CUDA_TRY(cudaSetDevice(e->device));
    RZ_TRY(sync_all(e));
    RZ_CUDA_TRY(cudaMemsetAsync(e->dp.status, 0, sizeof(Status), e->stream));
    setup_search_root_kernel<<<(e->dc.G + 127) / 128, 128, 0, e->stream>>>(e->dc, e->dp, own, enemy, player, keep_tree);
    RZ_LAUNCH_CHECK();
    RZ_CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int it = 0; it < 1000000; ++it) {
        for (int i = 0; i < 8; ++i) RZ_TRY(launch_wave(e));
        RZ_TRY(read_status(e));
        if (e->h_status->idle_slots >= (unsigned long long)e->dc.G) break;
    }
    int32_t* d_n; float* d_w;
    RZ_CUDA_TRY(cudaMalloc((void**)&d_n, 64 * 4));
    RZ_CUDA_TRY(cudaMalloc((void**)&d_w, 64 * 4));
    read_root_kernel<<<1, 1, 0, e->stream>>>(e->dc, e->dp, slot, d_n, d_w);
    cudaError_t ce = cudaMemcpyAsync(n_visit, d_n, 256, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(w_sum, d_w, 256, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    cudaFree(d_n); cudaFree(d_w);
    if (ce != cudaSuccess) { set_error("rz_engine_search_root: %s", cudaGetErrorString(ce)); return RZ_ECUDA; }
    return RZ_OK;
}

}  // extern "C"
