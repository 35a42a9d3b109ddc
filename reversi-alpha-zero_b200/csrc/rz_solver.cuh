// rz_solver.cuh -- endgame solver on the device (lib/alt/reversi_solver_cython.pyx:63-127, the variant agent/player.py:15
// imports): result (move, score) of a position for the side to move, identical to the reference's
//   * exactly = 1: best final disc difference and the FIRST move (ascending square) that reaches it (strict '<', :92-95);
//   * exactly = 0 (win/loss/draw): plain minimax in which every node stops at its first move with a positive score (:99);
//     the returned score is that early-stopped score, so only its sign and the chosen move are meaningful.
//
// Execution model: ONE LANE PER REQUEST.  A solve is an explicit stack machine whose whole state (SolveCtx: the frame
// stack) lives in global memory, so it can be advanced for a bounded number of cycles per launch and resumed in the next
// one: the engine never waits for its slowest request, a slot just keeps its simulations parked until their solves are
// done (the RESULT never depends on how the work was sliced).  The 32 lanes of a warp run 32 unrelated solves; they stay
// converged at the loop level (every iteration = one node step of each lane's own machine).
//   * WLD mode follows the reference literally (ascending move order, early stop) -- the reported move depends on that
//     order -- with the reference's cache as a per-lane transposition table in global memory;
//   * exact mode is alpha-beta (inner nodes only contribute their value): root moves in ascending order with the window
//     (best, 64), so a later root move replaces the best one only if it is strictly better, as in the reference; inner
//     nodes try the move that leaves the opponent the fewest replies first.
// Positions with more than kSolverMaxEmpties empty squares are refused (move = -1), the analogue of the reference's
// 30-second timeout (:78-79) after which the player falls back to the search.
#pragma once
#include "rz_bitboard.cuh"

namespace rz {
namespace solver {

constexpr int kSolverMaxEmpties = 12;
constexpr int kMaxDepth = 24;
constexpr int kOrderMinEmpties = 5;

struct Frame {  // 32 B
    u64 own, enemy, moves;
    int8_t best, alpha, beta, sign;  // sign: factor applied to this frame's value when it returns to its parent
    int8_t cur;                      // move being explored from this frame
    int8_t pad[3];
};

struct SolveCtx {
    u64 own, enemy;
    int8_t exactly;
    volatile int8_t done;  // written last by the solver, polled by the engine's tick kernel
    int8_t move, score;    // result: move -1 = none / refused
    int16_t depth;         // -1: not started yet
    int16_t pad;
    Frame f[kMaxDepth];
};

// Transposition table of the WLD mode (the reference's `cache` dict, reversi_solver_cython.pyx:83-90,102): direct-mapped,
// one table per solver lane in global memory.  An entry is three words (own ^ tag, enemy ^ tag, tag) with tag = upper 48
// bits of a hash of the position | marker | score byte, so a foreign entry fails validation.  A position's WLD value is
// a pure function of the position: entries stay valid across requests and never need clearing.
constexpr uint32_t kTtEntries = 1u << 10;  // per lane
constexpr int kTtWordsPerEntry = 4;        // 32 B
constexpr int kTtMinEmpties = 4;           // smaller subtrees are cheaper to recompute than to look up
struct TT {
    u64* base;  // nullptr: no table
    RZ_HD static u64 mix(u64 own, u64 enemy) {
        u64 h = own * 0x9E3779B97F4A7C15ULL ^ (enemy + 0x632BE59BD9B4E019ULL) * 0xC2B2AE3D27D4EB4FULL;
        h ^= h >> 31; h *= 0xD6E8FEB86659FD93ULL; h ^= h >> 29;
        return h;
    }
    RZ_HD bool probe(u64 own, u64 enemy, int& score) const {
        if (!base) return false;
        const u64 h = mix(own, enemy);
        const u64* e = base + (size_t)(h & (kTtEntries - 1)) * kTtWordsPerEntry;
        const u64 w0 = e[0], w1 = e[1], w2 = e[2];
        if ((w0 ^ w2) != own || (w1 ^ w2) != enemy || (w2 >> 16) != (h >> 16) || ((w2 >> 8) & 0xFF) != 0x5A) return false;
        score = (int)(int8_t)(w2 & 0xFF);
        return true;
    }
    RZ_HD void store(u64 own, u64 enemy, int score) const {
        if (!base) return;
        const u64 h = mix(own, enemy);
        u64* e = base + (size_t)(h & (kTtEntries - 1)) * kTtWordsPerEntry;
        const u64 w2 = (h & ~0xFFFFULL) | (0x5AULL << 8) | (u64)(uint8_t)(int8_t)score;
        e[0] = own ^ w2; e[1] = enemy ^ w2; e[2] = w2;
    }
};

RZ_HD void ctx_init(SolveCtx* c, u64 own, u64 enemy, int exactly) {
    c->own = own; c->enemy = enemy; c->exactly = (int8_t)exactly; c->move = -1; c->score = 0; c->depth = -1; c->pad = 0;
    c->done = 0;
}

// exact mode, inner nodes: the remaining move that leaves the opponent the fewest replies ("fastest first")
RZ_HD int pick_move(u64 own, u64 enemy, u64 moves, bool ordered) {
    if (!ordered || (moves & (moves - 1)) == 0) return ctz64(moves);
    int best_a = -1, best_mob = 99;
    for (u64 m = moves; m; m &= m - 1) {
        const int a = ctz64(m);
        const u64 fl = calc_flip(a, own, enemy);
        const int mob = popc64(find_correct_moves(enemy ^ fl, (own ^ fl) | (1ULL << a)));
        if (mob < best_mob) { best_mob = mob; best_a = a; }
    }
    return best_a;
}

RZ_HD long long global_ns() {
#ifdef __CUDA_ARCH__
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#else
    static long long fake_ns = 0;  // host build (tests/test_solver_host.py): a counter, so that deadlines can be exercised
    return ++fake_ns;
#endif
}
RZ_HD void publish_fence() {
#ifdef __CUDA_ARCH__
    __threadfence();
#endif
}

// Advance the request until it is finished or `deadline` (%globaltimer ns; 0 = no limit) has passed; at least 15 node steps
// are made per call, so a request always progresses.  Returns true when done.
RZ_HD bool solve_advance(SolveCtx* c, const TT& tt, long long deadline) {
    const bool exactly = c->exactly != 0;
    const bool cached = !exactly && tt.base != nullptr;
    int d = c->depth;
    Frame F;  // the frame being worked on lives in registers; c->f[0..d-1] hold its ancestors, c->f[d] is written on suspend
    if (d < 0) {  // first visit: set up the root frame
        const u64 legal = find_correct_moves(c->own, c->enemy);
        if (!legal || 64 - popc64(c->own | c->enemy) > kSolverMaxEmpties) { c->move = -1; c->score = 0; publish_fence(); c->done = 1; return true; }
        F.own = c->own; F.enemy = c->enemy; F.moves = legal; F.best = -100; F.alpha = -64; F.beta = 64; F.sign = 1; F.cur = -1;
        d = 0;
    } else {
        F = c->f[d];
    }
    int it = 0;
    while (true) {
        if (deadline && (++it & 15) == 0 && global_ns() > deadline) { c->f[d] = F; c->depth = (int16_t)d; return false; }
        if (F.moves == 0 || (!exactly && F.best > 0) || (exactly && F.best >= F.beta)) {
            if (cached && d > 0 && 64 - popc64(F.own | F.enemy) >= kTtMinEmpties) tt.store(F.own, F.enemy, F.best);
            if (d == 0) { c->score = F.best; publish_fence(); c->done = 1; c->depth = 0; return true; }
            const int v = F.best * F.sign;
            F = c->f[--d];
            if (F.best < v) { F.best = (int8_t)v; if (d == 0) c->move = F.cur; }
            continue;
        }
        // root: ascending order (the reference's tie-break); inner exact nodes: fastest first; WLD: ascending everywhere
        const int a = pick_move(F.own, F.enemy, F.moves, exactly && d > 0 && 64 - popc64(F.own | F.enemy) >= kOrderMinEmpties);
        F.moves &= ~(1ULL << a);
        F.cur = (int8_t)a;
        const u64 fl = calc_flip(a, F.own, F.enemy);
        const u64 own2 = (F.own ^ fl) | (1ULL << a), en2 = F.enemy ^ fl;
        const int lo = F.best > F.alpha ? F.best : F.alpha;  // alpha-beta lower bound at this node (exact mode)
        const u64 m = find_correct_moves(en2, own2);
        const u64 m_self = m ? 0 : find_correct_moves(own2, en2);
        if (!m && !m_self) {  // game over
            const int score = popc64(own2) - popc64(en2);
            if (F.best < score) { F.best = (int8_t)score; if (d == 0) c->move = (int8_t)a; }
            continue;
        }
        if (cached && 64 - popc64(own2 | en2) >= kTtMinEmpties) {  // child already solved?
            int sc;
            if (m ? tt.probe(en2, own2, sc) : tt.probe(own2, en2, sc)) {
                const int v = m ? -sc : sc;
                if (F.best < v) { F.best = (int8_t)v; if (d == 0) c->move = (int8_t)a; }
                continue;
            }
        }
        if (d + 1 >= kMaxDepth) { c->move = -1; c->score = 0; publish_fence(); c->done = 1; return true; }  // cannot happen within kSolverMaxEmpties
        c->f[d++] = F;  // descend: park this frame, the child becomes the working frame
        const int8_t p_beta = F.beta;
        if (m) {  // opponent to move
            F.own = en2; F.enemy = own2; F.moves = m; F.sign = -1;
            F.alpha = (int8_t)(-p_beta); F.beta = (int8_t)(-lo);
        } else {  // pass: same side again, no sign flip
            F.own = own2; F.enemy = en2; F.moves = m_self; F.sign = 1;
            F.alpha = (int8_t)lo; F.beta = p_beta;
        }
        F.best = -100; F.cur = -1;
    }
}

}  // namespace solver
}  // namespace rz
