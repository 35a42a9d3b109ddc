// rz_bitboard.cuh -- Reversi rules on 64-bit bitboards, __host__ __device__.
//
// Same functions as the reference's lib/bitboard.py (find_correct_moves :53-67, calc_flip :70-92,
// flip_vertical :119-125, rotate90 :154, rotate180 :158) and env/reversi_env.py (step :42-74,
// _game_over :76-85), written for the GPU: move generation uses a parallel-prefix (Kogge-Stone)
// fill per direction instead of six dependent shift steps, the 180-degree rotation is one bit
// reversal (BREV), popcount / find-first-set are single instructions.  Results are bit-identical to
// the reference on every input the reference accepts (tests/test_bitboard_gpu.py, oracle/).
//
// Bit numbering (lib/bitboard.py:11-17): bit i = square y*8+x, bit 0 = top-left, bit 63 = bottom-right.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RZ_HD __host__ __device__ __forceinline__
#else
#define RZ_HD static inline
#endif

namespace rz {

typedef uint64_t u64;

constexpr u64 kNotEdgeLR = 0x7e7e7e7e7e7e7e7eULL;  // columns 1..6
constexpr u64 kStartBlack = (0x10ULL << 24) | (0x08ULL << 32);  // env/reversi_env.py:135
constexpr u64 kStartWhite = (0x08ULL << 24) | (0x10ULL << 32);  // env/reversi_env.py:136

RZ_HD int popc64(u64 x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

RZ_HD u64 brev64(u64 x) {  // bit i -> bit 63-i  == rotate180 (lib/bitboard.py:158)
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0fULL) | ((x & 0x0f0f0f0f0f0f0f0fULL) << 4);
    return __builtin_bswap64(x);
#endif
}

RZ_HD int ctz64(u64 x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

// One direction pair (shift S toward higher and lower bit indices) of the legal-move search.
// `e` is the opponent mask already restricted so that a shift by S never wraps around a board edge.
// Kogge-Stone: after the three doubling steps `up` / `dn` hold every opponent disc that is connected to an
// own disc through an unbroken run of opponent discs (runs are at most 6 long).
template <int S>
RZ_HD u64 mobility_dir(u64 own, u64 e) {
    u64 up = e & (own << S), dn = e & (own >> S);
    u64 eu = e & (e << S), ed = e & (e >> S);
    up |= e & (up << S);          dn |= e & (dn >> S);
    up |= eu & (up << (2 * S));   dn |= ed & (dn >> (2 * S));
    up |= eu & (up << (2 * S));   dn |= ed & (dn >> (2 * S));
    return (up << S) | (dn >> S);
}

// lib/bitboard.py:53-67: squares where `own` may move.
// (Moving the 64-bit shifts to the FMA pipe as IMAD / IMAD.HI by powers of two was tried and measured slower,
// profiles/k1_shift_modes_r01.json.)
RZ_HD u64 find_correct_moves(u64 own, u64 enemy) {
    const u64 eh = enemy & kNotEdgeLR;  // horizontal / diagonal runs never include an edge column
    u64 m = mobility_dir<1>(own, eh) | mobility_dir<7>(own, eh) | mobility_dir<9>(own, eh) |
            mobility_dir<8>(own, enemy);  // vertical: bits shifted past row 0 / row 7 fall off the word
    return m & ~(own | enemy);
}

// lib/bitboard.py:84-92: the four rays that run toward higher bit indices from `pos`
// (down, right, down-left, down-right).  For each ray the carry of (e | ~ray) + 1 ripples through the
// contiguous opponent discs and lands on the first non-opponent square; if that square is ours the
// discs in between are flipped.
RZ_HD u64 flip_rays_up(int pos, u64 own, u64 enemy) {
    const u64 eh = enemy & kNotEdgeLR;
    u64 flipped = 0, ray, out;
    ray = 0x0101010101010100ULL << pos;  out = ray & ((enemy | ~ray) + 1) & own;  flipped |= (out - (out != 0)) & ray;
    ray = 0x00000000000000feULL << pos;  out = ray & ((eh | ~ray) + 1) & own;     flipped |= (out - (out != 0)) & ray;
    ray = 0x0002040810204080ULL << pos;  out = ray & ((eh | ~ray) + 1) & own;     flipped |= (out - (out != 0)) & ray;
    ray = 0x8040201008040200ULL << pos;  out = ray & ((eh | ~ray) + 1) & own;     flipped |= (out - (out != 0)) & ray;
    return flipped;
}

// lib/bitboard.py:70-81: discs flipped by a move at pos; the other four rays are obtained by
// rotating the board by 180 degrees (one BREV each way).  No legality / occupancy check, as in the reference.
RZ_HD u64 calc_flip(int pos, u64 own, u64 enemy) {
    return flip_rays_up(pos, own, enemy) | brev64(flip_rays_up(63 - pos, brev64(own), brev64(enemy)));
}

RZ_HD u64 flip_vertical(u64 x) {  // row y <-> row 7-y  == byte swap (lib/bitboard.py:119-125)
#if defined(__CUDA_ARCH__)
    return ((u64)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
#else
    return __builtin_bswap64(x);
#endif
}

RZ_HD u64 flip_diag_a1h8(u64 x) {  // lib/bitboard.py:141-151 (three delta swaps)
    u64 t;
    t = 0x0f0f0f0f00000000ULL & (x ^ (x << 28)); x ^= t ^ (t >> 28);
    t = 0x3333000033330000ULL & (x ^ (x << 14)); x ^= t ^ (t >> 14);
    t = 0x5500550055005500ULL & (x ^ (x << 7));  x ^= t ^ (t >> 7);
    return x;
}

RZ_HD u64 rotate90(u64 x) { return flip_diag_a1h8(flip_vertical(x)); }  // clockwise, lib/bitboard.py:154

// t in 0..7: flip_vertical if (t & 4), then (t & 3) x rotate90 (agent/player.py:166-179, :300-305).
RZ_HD u64 dihedral(u64 x, int t) {
    if (t & 4) x = flip_vertical(x);
    if (t & 2) x = brev64(x);
    if (t & 1) x = rotate90(x);
    return x;
}
// square s -> its location after dihedral(., t)
RZ_HD int dihedral_square(int s, int t) { return ctz64(dihedral(1ULL << s, t)); }

struct EnvState {
    u64 black, white;
    uint8_t next_player;  // 1 black, 2 white
    uint8_t turn, done, winner;
};

RZ_HD void env_reset(EnvState& s) {
    s.black = kStartBlack; s.white = kStartWhite; s.next_player = 1; s.turn = 0; s.done = 0; s.winner = 0;
}

// env/reversi_env.py:34-40 (+ Board.__init__ :133-140: a zero bitboard silently becomes the start stones)
RZ_HD void env_update(EnvState& s, u64 black, u64 white, int next_player) {
    s.black = black ? black : kStartBlack;
    s.white = white ? white : kStartWhite;
    s.next_player = (uint8_t)next_player;
    s.turn = (uint8_t)(popc64(s.black) + popc64(s.white) - 4);
    s.done = 0; s.winner = 0;
}

RZ_HD uint8_t winner_by_count(u64 black, u64 white) {  // env/reversi_env.py:76-85
    int b = popc64(black), w = popc64(white);
    return b > w ? 1 : (b < w ? 2 : 3);
}

// env/reversi_env.py:42-74.  action < 0 == None (resign).  Returns the legal-move mask of the side to
// move after the step (0 when the game is over).
RZ_HD u64 env_step(EnvState& s, int action) {
    const bool black_to_move = s.next_player == 1;
    if (action < 0) {  // :49-51, :95-104
        s.winner = black_to_move ? 2 : 1; s.done = 1; return 0;
    }
    u64 own = black_to_move ? s.black : s.white, enemy = black_to_move ? s.white : s.black;
    u64 fl = calc_flip(action, own, enemy);
    if (fl == 0) {  // illegal move loses, :56-59
        s.winner = black_to_move ? 2 : 1; s.done = 1; return 0;
    }
    own ^= fl; own |= 1ULL << action; enemy ^= fl;
    s.black = black_to_move ? own : enemy;
    s.white = black_to_move ? enemy : own;
    s.turn += 1;
    u64 m = find_correct_moves(enemy, own);
    if (m) { s.next_player = black_to_move ? 2 : 1; return m; }  // :67-68
    m = find_correct_moves(own, enemy);
    if (m) return m;                                              // :69-70 auto-pass
    s.done = 1; s.winner = winner_by_count(s.black, s.white);    // :71-72
    return 0;
}

}  // namespace rz
