// rz_solver.cuh -- endgame solver on the device (lib/alt/reversi_solver_cython.pyx:63-127, the variant agent/player.py:15
// imports): result (move, score) of a position for the side to move, identical to the reference's
//   * exactly = 1: best final disc difference, FIRST move (ascending square) that reaches it (strict '<' at :92-95);
//   * exactly = 0 (win/loss/draw): plain minimax in which every node stops at its first move with a positive score (:99);
//     the returned score is that early-stopped score, so only its sign and the chosen move are meaningful.
// One WARP solves one position: the two top plies are enumerated, every grand-child position becomes a task, the lanes
// solve the tasks independently (iterative negamax on a private stack; alpha-beta in exact mode -- inner nodes only need
// their value, which pruning does not change; the literal early-stop recursion in WLD mode), and the top two plies are
// then combined sequentially in the reference's move order, so ties and early stops resolve exactly as in the reference.
// Positions with more than kSolverMaxEmpties empty squares are refused (move = -1), the analogue of the reference's
// 30-second timeout (:78-79) after which the player falls back to the search.
#pragma once
#include "rz_bitboard.cuh"

namespace rz {
namespace solver {

constexpr int kSolverMaxEmpties = 12;
constexpr int kMaxTasks = 256;   // > 12 * 11 second-ply positions
constexpr int kMaxDepth = 24;

// Transposition table of the WLD mode (the reference's `cache` dict, reversi_solver_cython.pyx:83-90,102): one
// direct-mapped table per warp in global memory, shared by its 32 lanes without locks.  An entry is three words
// (own ^ tag, enemy ^ tag, tag) with tag = 48-bit hash of the position | score byte, so a torn or foreign entry fails
// validation instead of returning a wrong score.  A position's WLD value is a pure function of the position, so
// entries stay valid across requests and never need clearing.  Without it a 10-empties WLD solve walks millions of
// nodes (the reference is only fast because of its cache).
constexpr uint32_t kTtEntries = 1u << 13;           // per warp
constexpr int kTtWordsPerEntry = 4;                 // 32 B
constexpr int kTtMinEmpties = 4;                    // smaller subtrees are cheaper to recompute than to look up
struct TT {
    u64* base;  // nullptr: no table
    __device__ __forceinline__ static u64 mix(u64 own, u64 enemy) {
        u64 h = own * 0x9E3779B97F4A7C15ULL ^ (enemy + 0x632BE59BD9B4E019ULL) * 0xC2B2AE3D27D4EB4FULL;
        h ^= h >> 31; h *= 0xD6E8FEB86659FD93ULL; h ^= h >> 29;
        return h;
    }
    __device__ __forceinline__ bool probe(u64 own, u64 enemy, int& score) const {
        if (!base) return false;
        const u64 h = mix(own, enemy);
        const u64* e = base + (size_t)(h & (kTtEntries - 1)) * kTtWordsPerEntry;
        const u64 w0 = e[0], w1 = e[1], w2 = e[2];
        if ((w0 ^ w2) != own || (w1 ^ w2) != enemy || (w2 >> 16) != (h >> 16) || ((w2 >> 8) & 0xFF) != 0x5A) return false;
        score = (int)(int8_t)(w2 & 0xFF);
        return true;
    }
    __device__ __forceinline__ void store(u64 own, u64 enemy, int score) const {
        if (!base) return;
        const u64 h = mix(own, enemy);
        u64* e = base + (size_t)(h & (kTtEntries - 1)) * kTtWordsPerEntry;
        const u64 w2 = (h & ~0xFFFFULL) | (0x5AULL << 8) | (u64)(uint8_t)(int8_t)score;
        e[0] = own ^ w2; e[1] = enemy ^ w2; e[2] = w2;
    }
};

struct Frame {
    u64 own, enemy, moves;
    int8_t best, alpha, beta, sign;  // sign: factor applied to this frame's value when it returns to its parent
};
constexpr int kOrderMinEmpties = 5;

// exact mode only: next move to try = the remaining move that leaves the opponent the fewest replies ("fastest first").
// Inner nodes of the exact search only contribute their VALUE (the first-best-move rule of the reference is resolved at
// the two top plies, which keep the ascending order), and alpha-beta with any move order returns the same value.
__device__ __forceinline__ int pick_move(u64 own, u64 enemy, u64 moves, bool ordered) {
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

// value of the position for the side to move (`own`, which has at least one legal move)
__device__ inline int solve_subtree(u64 own, u64 enemy, bool exactly, const TT& tt) {
    const bool cached = !exactly && tt.base != nullptr;  // exact mode prunes (alpha-beta), its node values are bounds: not cached
    if (cached && 64 - popc64(own | enemy) >= kTtMinEmpties) {
        int sc;
        if (tt.probe(own, enemy, sc)) return sc;
    }
    Frame f[kMaxDepth];
    int d = 0;
    f[0].own = own; f[0].enemy = enemy; f[0].moves = find_correct_moves(own, enemy);
    f[0].best = -100; f[0].alpha = -64; f[0].beta = 64; f[0].sign = 1;
    while (true) {
        Frame& F = f[d];
        if (F.moves == 0 || (!exactly && F.best > 0) || (exactly && F.best >= F.beta)) {
            if (cached && 64 - popc64(F.own | F.enemy) >= kTtMinEmpties) tt.store(F.own, F.enemy, F.best);
            if (d == 0) return F.best;
            const int v = F.best * F.sign;
            --d;
            if (f[d].best < v) f[d].best = (int8_t)v;
            continue;
        }
        const int a = pick_move(F.own, F.enemy, F.moves, exactly && 64 - popc64(F.own | F.enemy) >= kOrderMinEmpties);
        F.moves &= ~(1ULL << a);
        const u64 fl = calc_flip(a, F.own, F.enemy);
        const u64 own2 = (F.own ^ fl) | (1ULL << a), en2 = F.enemy ^ fl;
        const int lo = F.best > F.alpha ? F.best : F.alpha;  // alpha-beta lower bound at this node (exact mode)
        u64 m = find_correct_moves(en2, own2);
        if (cached && 64 - popc64(own2 | en2) >= kTtMinEmpties) {  // child already solved?
            int sc;
            if (m ? tt.probe(en2, own2, sc) : (find_correct_moves(own2, en2) != 0 && tt.probe(own2, en2, sc))) {
                const int v = m ? -sc : sc;
                if (F.best < v) F.best = (int8_t)v;
                continue;
            }
        }
        if (m) {  // opponent to move
            if (d + 1 >= kMaxDepth) return F.best;  // cannot happen for <= kSolverMaxEmpties empties
            Frame& C = f[++d];
            C.own = en2; C.enemy = own2; C.moves = m; C.best = -100; C.sign = -1;
            C.alpha = (int8_t)(-F.beta); C.beta = (int8_t)(-lo);
        } else if ((m = find_correct_moves(own2, en2)) != 0) {  // pass: same side again, no sign flip
            if (d + 1 >= kMaxDepth) return F.best;
            Frame& C = f[++d];
            C.own = own2; C.enemy = en2; C.moves = m; C.best = -100; C.sign = 1;
            C.alpha = (int8_t)lo; C.beta = F.beta;
        } else {
            const int score = popc64(own2) - popc64(en2);
            if (F.best < score) F.best = (int8_t)score;
        }
    }
}

// position after `own` plays at a: *terminal -> score for the mover; else the next position in ITS mover's frame and the
// factor that converts its value back to the original mover's frame
__device__ __forceinline__ bool after_move(u64 own, u64 enemy, int a, u64& nown, u64& nenemy, int& sign, int& score) {
    const u64 fl = calc_flip(a, own, enemy);
    const u64 own2 = (own ^ fl) | (1ULL << a), en2 = enemy ^ fl;
    if (find_correct_moves(en2, own2)) { nown = en2; nenemy = own2; sign = -1; return false; }
    if (find_correct_moves(own2, en2)) { nown = own2; nenemy = en2; sign = 1; return false; }
    score = popc64(own2) - popc64(en2);
    return true;
}

// Called by a full warp with identical arguments; `vals` is a per-warp scratch array of kMaxTasks int8 in shared memory.
// Returns (move, score) in all lanes; move = -1: no legal move or position refused.
__device__ inline void solve_warp(u64 own, u64 enemy, bool exactly, int8_t* vals, int lane, int& move_out, int& score_out, const TT& tt) {
    move_out = -1; score_out = -100;
    const u64 legal = find_correct_moves(own, enemy);
    if (!legal || 64 - popc64(own | enemy) > kSolverMaxEmpties) return;
    if (!exactly) {
        // WLD mode: the reference's early stop makes the search lazy -- it never looks at the root moves behind the first
        // winning one, and with the transposition table the whole tree is a few thousand nodes -- so evaluating all
        // second-ply subtrees in parallel does far MORE work than the sequential walk.  One lane walks the tree in the
        // reference's order (the request-level parallelism comes from the thousands of warps in flight).
        int best = -100, best_move = -1;
        if (lane == 0) {
            for (u64 m1 = legal; m1; m1 &= m1 - 1) {
                if (best > 0) break;                              // reversi_solver_cython.pyx:99 at the root
                const int a1 = ctz64(m1);
                u64 o1, e1; int s1, sc1;
                const int v1 = after_move(own, enemy, a1, o1, e1, s1, sc1) ? sc1 : s1 * solve_subtree(o1, e1, false, tt);
                if (best < v1) { best = v1; best_move = a1; }
            }
        }
        move_out = __shfl_sync(0xffffffffu, best_move, 0);
        score_out = __shfl_sync(0xffffffffu, best, 0);
        return;
    }
    // exact mode -- pass 1: every lane enumerates the grand-child tasks identically and solves those with index == lane (mod 32)
    int t = 0;
    for (u64 m1 = legal; m1; m1 &= m1 - 1) {
        u64 o1, e1; int s1, sc1;
        if (after_move(own, enemy, ctz64(m1), o1, e1, s1, sc1)) continue;
        for (u64 m2 = find_correct_moves(o1, e1); m2; m2 &= m2 - 1) {
            u64 o2, e2; int s2, sc2;
            if (after_move(o1, e1, ctz64(m2), o2, e2, s2, sc2)) continue;
            if ((t & 31) == lane && t < kMaxTasks) vals[t] = (int8_t)solve_subtree(o2, e2, exactly, tt);
            ++t;
        }
    }
    __syncwarp();
    if (t > kMaxTasks) return;  // refused (cannot happen within kSolverMaxEmpties)
    // pass 2: the two top plies in the reference's order (all lanes compute the same thing)
    t = 0;
    int best = -100, best_move = -1;
    for (u64 m1 = legal; m1; m1 &= m1 - 1) {
        if (!exactly && best > 0) break;                      // :99 at the root
        const int a1 = ctz64(m1);
        u64 o1, e1; int s1, sc1;
        int v1;
        if (after_move(own, enemy, a1, o1, e1, s1, sc1)) {
            v1 = sc1;
        } else {
            int b1 = -100;
            bool stopped = false;
            for (u64 m2 = find_correct_moves(o1, e1); m2; m2 &= m2 - 1) {
                u64 o2, e2; int s2, sc2;
                const bool term = after_move(o1, e1, ctz64(m2), o2, e2, s2, sc2);
                const int v2 = term ? sc2 : s2 * (int)vals[t];
                if (!term) ++t;
                if (stopped) continue;                        // keep the task counter in step with pass 1
                if (b1 < v2) b1 = v2;
                if (!exactly && b1 > 0) stopped = true;       // :99 at the child
            }
            v1 = s1 * b1;
        }
        if (best < v1) { best = v1; best_move = a1; }
    }
    // tasks of root moves skipped by the early stop are simply not consumed
    move_out = best_move; score_out = best;
}

}  // namespace solver
}  // namespace rz
