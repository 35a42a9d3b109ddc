// rz_bitsliced.cuh -- bit-sliced formulation of the batched bitboard operators (K1), __host__ __device__.
//
// The scalar formulation (rz_bitboard.cuh) spends ~180 integer-pipe instructions per position on emulated 64-bit shifts
// and masks (Kogge-Stone fills per direction) and is bound by the integer issue rate at a third of the HBM roofline
// (profiles/k1_r01_ncu_summary.txt).  Here ONE THREAD owns 32 positions and holds them transposed: register S[k] carries
// square k of all 32 positions (bit i = position i).  A shift along a board direction is then a register rename, the edge
// masks disappear, and the fill of a whole ray is ONE 3-input logic instruction per square --
//     chain[s] = enemy[s] & (own[prev(s)] | chain[prev(s)])        (lib/bitboard.py:95-116: six shift-and-mask steps)
// -- for 32 positions at a time.  What remains is getting in and out of that layout: a 32x32 bit-matrix transpose in
// registers per 32-bit board half (two byte-permute stages + three delta-swap stages), 4 in + 2 out per 32 positions.
// Total: ~67 integer-pipe instructions per position instead of ~180.  Thread t of a warp handles positions
// {tile * 1024 + i * 32 + t}, so every load / store instruction of the warp is one contiguous 256-byte run.
// Results are bit-identical to the scalar functions (tests/test_host_mirror.py on the host twin, tests/test_bitboard_gpu.py).
#pragma once
#include "rz_bitboard.cuh"

namespace rz {
namespace bs {

RZ_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}

// 32x32 bit-matrix transpose in registers: afterwards bit i of w[k] is what bit k of w[i] was
RZ_HD void transpose32(uint32_t* w) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {  // 16-bit blocks: one byte permute per output word
        const uint32_t a = w[k], b = w[k + 16];
        w[k] = byte_perm(a, b, 0x5410);
        w[k + 16] = byte_perm(a, b, 0x7632);
    }
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 16)
#pragma unroll
        for (int k = k0; k < k0 + 8; ++k) {  // 8-bit blocks
            const uint32_t a = w[k], b = w[k + 8];
            w[k] = byte_perm(a, b, 0x6240);
            w[k + 8] = byte_perm(a, b, 0x7351);
        }
#pragma unroll
    for (int j = 4; j >= 1; j >>= 1) {  // 4-, 2-, 1-bit blocks: delta swaps
        const uint32_t m = j == 4 ? 0x0f0f0f0fu : (j == 2 ? 0x33333333u : 0x55555555u);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (k & j) continue;
            const uint32_t t = ((w[k] >> j) ^ w[k + j]) & m;
            w[k + j] ^= t;
            w[k] ^= t << j;
        }
    }
}

// One direction pair (+-(DX, DY)) of find_correct_moves for 32 positions: for every line of the board along the direction,
// cF[i] = "square i of the line holds an opponent disc and an unbroken run of opponent discs behind it ends at one of
// ours" walking forward, cB[i] the same walking backward; the square after (before) such a run is a candidate move.
template <int DX, int DY>
RZ_HD void moves_dir(const uint32_t* O, const uint32_t* E, uint32_t* M) {
#pragma unroll
    for (int s = 0; s < 64; ++s) {
        const int x0 = s & 7, y0 = s >> 3;
        const int px = x0 - DX, py = y0 - DY;
        if (px >= 0 && px < 8 && py >= 0 && py < 8) continue;  // not the first square of its line
        int n = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int x = x0 + i * DX, y = y0 + i * DY;
            if (x >= 0 && x < 8 && y >= 0 && y < 8) n = i + 1;
        }
        if (n < 3) continue;  // own + opponent + empty need three squares
        uint32_t cF[8], cB[8];
        cF[0] = 0; cB[n - 1] = 0;
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            if (i >= n) break;
            const int cur = (x0 + i * DX) + 8 * (y0 + i * DY), prv = (x0 + (i - 1) * DX) + 8 * (y0 + (i - 1) * DY);
            cF[i] = E[cur] & (O[prv] | cF[i - 1]);
        }
#pragma unroll
        for (int i = 6; i >= 0; --i) {
            if (i > n - 2) continue;
            const int cur = (x0 + i * DX) + 8 * (y0 + i * DY), nxt = (x0 + (i + 1) * DX) + 8 * (y0 + (i + 1) * DY);
            cB[i] = E[cur] & (O[nxt] | cB[i + 1]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i >= n) break;
            const int cur = (x0 + i * DX) + 8 * (y0 + i * DY);
            const uint32_t f = i >= 2 ? cF[i - 1] : 0u, b = i <= n - 3 ? cB[i + 1] : 0u;
            M[cur] |= f | b;
        }
    }
}

// lib/bitboard.py:53-67 for 32 positions held by one thread: own[i], enemy[i] in (i = 0..31), legal-move masks out
RZ_HD void find_correct_moves32(const u64* own, const u64* enemy, u64* out) {
    uint32_t O[64], E[64], M[64];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        O[i] = (uint32_t)own[i]; O[32 + i] = (uint32_t)(own[i] >> 32);
        E[i] = (uint32_t)enemy[i]; E[32 + i] = (uint32_t)(enemy[i] >> 32);
    }
    transpose32(O); transpose32(O + 32);   // O[k]: square k of the 32 positions
    transpose32(E); transpose32(E + 32);
#pragma unroll
    for (int k = 0; k < 64; ++k) M[k] = 0;
    moves_dir<1, 0>(O, E, M);
    moves_dir<0, 1>(O, E, M);
    moves_dir<1, 1>(O, E, M);
    moves_dir<-1, 1>(O, E, M);
#pragma unroll
    for (int k = 0; k < 64; ++k) M[k] &= ~(O[k] | E[k]);  // only empty squares (bitboard.py:66)
    transpose32(M); transpose32(M + 32);
#pragma unroll
    for (int i = 0; i < 32; ++i) out[i] = (u64)M[i] | ((u64)M[32 + i] << 32);
}

// One direction (DX, DY) of calc_flip for 32 positions: the run of opponent discs that starts right after the move square
// P (one-hot per position) and is closed by one of ours.  f[i]: opponent discs reached from P walking forward; g[i]: those
// of them from which the walk goes on to an own disc.  Accumulates into F only the squares of board half HALF (0: squares
// 0-31, 1: 32-63) -- the flip mask of a position is assembled in two passes to stay within the register file.
template <int DX, int DY, int HALF>
RZ_HD void flips_dir(const uint32_t* O, const uint32_t* E, const uint32_t* P, uint32_t* F) {
#pragma unroll
    for (int s = 0; s < 64; ++s) {
        const int x0 = s & 7, y0 = s >> 3;
        const int px = x0 - DX, py = y0 - DY;
        if (px >= 0 && px < 8 && py >= 0 && py < 8) continue;  // not the first square of its line
        int n = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int x = x0 + i * DX, y = y0 + i * DY;
            if (x >= 0 && x < 8 && y >= 0 && y < 8) n = i + 1;
        }
        if (n < 3) continue;  // move square + opponent + own need three squares
        uint32_t f[8], g[8];
        f[0] = 0; g[n - 1] = 0;
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            if (i >= n) break;
            const int cur = (x0 + i * DX) + 8 * (y0 + i * DY), prv = (x0 + (i - 1) * DX) + 8 * (y0 + (i - 1) * DY);
            f[i] = E[cur] & (P[prv] | f[i - 1]);
        }
#pragma unroll
        for (int i = 6; i >= 1; --i) {
            if (i > n - 2) continue;
            const int nxt = (x0 + (i + 1) * DX) + 8 * (y0 + (i + 1) * DY);
            g[i] = f[i] & (O[nxt] | g[i + 1]);
        }
#pragma unroll
        for (int i = 1; i < 7; ++i) {
            if (i > n - 2) break;
            const int cur = (x0 + i * DX) + 8 * (y0 + i * DY);
            if ((cur >> 5) == HALF) F[cur & 31] |= g[i];
        }
    }
}

template <int HALF>
RZ_HD void flips_half(const uint32_t* O, const uint32_t* E, const uint32_t* P, uint32_t* F) {
#pragma unroll
    for (int k = 0; k < 32; ++k) F[k] = 0;
    flips_dir<1, 0, HALF>(O, E, P, F);  flips_dir<-1, 0, HALF>(O, E, P, F);
    flips_dir<0, 1, HALF>(O, E, P, F);  flips_dir<0, -1, HALF>(O, E, P, F);
    flips_dir<1, 1, HALF>(O, E, P, F);  flips_dir<-1, -1, HALF>(O, E, P, F);
    flips_dir<-1, 1, HALF>(O, E, P, F); flips_dir<1, -1, HALF>(O, E, P, F);
    transpose32(F);
}

// lib/bitboard.py:70-92 for 32 positions held by one thread; like the reference it ignores what stands on pos[i] itself.
// own and enemy sharing a square is not a board, but the reference's carry-trick arithmetic gives such input a definite answer
// (a run ends at the first NON-opponent square) that the ray walk does not reproduce: a group of 32 that contains such a
// position goes through the scalar code -- here when CHECK_OVERLAP (host twin), in the kernels before they call this
// (compact loops over the staged inputs, so that the hot path carries no 32-fold unrolled scalar code).
RZ_HD bool any_overlap32(const u64* own, const u64* enemy) {
    u64 overlap = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) overlap |= own[i] & enemy[i];
    return overlap != 0;
}
template <bool CHECK_OVERLAP = true>
RZ_HD void calc_flip32(const uint8_t* pos, const u64* own, const u64* enemy, u64* out) {
    if (CHECK_OVERLAP && any_overlap32(own, enemy)) {
        for (int i = 0; i < 32; ++i) out[i] = calc_flip(pos[i] & 63, own[i], enemy[i]);
        return;
    }
    uint32_t O[64], E[64], P[64], F[32];
    {
        uint32_t pb[32];  // after the transpose pb[b] holds bit b of the 32 move squares
#pragma unroll
        for (int i = 0; i < 32; ++i) pb[i] = pos[i] & 63u;
        transpose32(pb);
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            lo[c] = ((c & 1) ? pb[0] : ~pb[0]) & ((c & 2) ? pb[1] : ~pb[1]) & ((c & 4) ? pb[2] : ~pb[2]);
            hi[c] = ((c & 1) ? pb[3] : ~pb[3]) & ((c & 2) ? pb[4] : ~pb[4]) & ((c & 4) ? pb[5] : ~pb[5]);
        }
#pragma unroll
        for (int k = 0; k < 64; ++k) P[k] = lo[k & 7] & hi[k >> 3];
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        O[i] = (uint32_t)own[i]; O[32 + i] = (uint32_t)(own[i] >> 32);
        E[i] = (uint32_t)enemy[i]; E[32 + i] = (uint32_t)(enemy[i] >> 32);
    }
    transpose32(O); transpose32(O + 32);
    transpose32(E); transpose32(E + 32);
    flips_half<0>(O, E, P, F);
#pragma unroll
    for (int i = 0; i < 32; ++i) out[i] = F[i];
    flips_half<1>(O, E, P, F);
#pragma unroll
    for (int i = 0; i < 32; ++i) out[i] |= (u64)F[i] << 32;
}

}  // namespace bs
}  // namespace rz
