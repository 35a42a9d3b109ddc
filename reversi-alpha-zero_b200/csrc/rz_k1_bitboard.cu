// rz_k1_bitboard.cu -- K1: stateless batched bitboard operators + their C ABI.
//
// HBM-bound integer work: 24 B / position for the legal-move mask, 25 B for the flip
// mask, 42+ B for the fused step.  Layout is structure-of-arrays so every warp access is a fully
// coalesced run; each thread handles two consecutive positions through 128-bit loads/stores, inputs
// are streamed with ld.global.nc.L1::no_allocate, outputs with st.global.cs.  Grids are a multiple of
// the SM count (persistent grid-stride loop).
#include <stdlib.h>
#include "rz_bitboard.cuh"
#include "rz_common.cuh"

namespace rz {

__device__ __forceinline__ ulonglong2 ldg_stream_u64x2(const u64* p) {
    ulonglong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ u64 ldg_stream_u64(const u64* p) {
    u64 r;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_u64x2(u64* p, u64 a, u64 b) {
    asm volatile("st.global.cs.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k1_find_correct_moves(const u64* __restrict__ own, const u64* __restrict__ enemy,
                                                                  u64* __restrict__ out, size_t n, int vec_ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {
        const size_t pairs = n >> 1;
        for (size_t i = tid; i < pairs; i += stride) {
            ulonglong2 o = ldg_stream_u64x2(own + 2 * i), e = ldg_stream_u64x2(enemy + 2 * i);
            stg_stream_u64x2(out + 2 * i, find_correct_moves(o.x, e.x), find_correct_moves(o.y, e.y));
        }
        if ((n & 1) && tid == 0) out[n - 1] = find_correct_moves(own[n - 1], enemy[n - 1]);
    } else {
        for (size_t i = tid; i < n; i += stride) out[i] = find_correct_moves(ldg_stream_u64(own + i), ldg_stream_u64(enemy + i));
    }
}

__global__ void __launch_bounds__(kThreads) k1_calc_flip(const uint8_t* __restrict__ pos, const u64* __restrict__ own,
                                                         const u64* __restrict__ enemy, u64* __restrict__ out, size_t n,
                                                         int vec_ok) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {
        const size_t pairs = n >> 1;
        for (size_t i = tid; i < pairs; i += stride) {
            ulonglong2 o = ldg_stream_u64x2(own + 2 * i), e = ldg_stream_u64x2(enemy + 2 * i);
            const uchar2 p = *reinterpret_cast<const uchar2*>(pos + 2 * i);
            stg_stream_u64x2(out + 2 * i, calc_flip(p.x & 63, o.x, e.x), calc_flip(p.y & 63, o.y, e.y));
        }
        if ((n & 1) && tid == 0) out[n - 1] = calc_flip(pos[n - 1] & 63, own[n - 1], enemy[n - 1]);
    } else {
        for (size_t i = tid; i < n; i += stride) out[i] = calc_flip(pos[i] & 63, own[i], enemy[i]);
    }
}

__global__ void __launch_bounds__(kThreads) k1_step(u64* __restrict__ black, u64* __restrict__ white, uint8_t* __restrict__ next_player,
                                                    uint8_t* __restrict__ turn, uint8_t* __restrict__ done, uint8_t* __restrict__ winner,
                                                    const int8_t* __restrict__ action, u64* __restrict__ legal_out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        EnvState s{black[i], white[i], next_player[i], turn[i], done[i], winner[i]};
        u64 legal = env_step(s, action[i]);
        black[i] = s.black; white[i] = s.white; next_player[i] = s.next_player;
        turn[i] = s.turn; done[i] = s.done; winner[i] = s.winner;
        if (legal_out) legal_out[i] = legal;
    }
}

__global__ void __launch_bounds__(kThreads) k1_dihedral(const u64* __restrict__ x, const uint8_t* __restrict__ t, u64* __restrict__ out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = dihedral(x[i], t[i] & 7);
}

static int grid_for(size_t work_items) {
    const int sms = num_sms();
    size_t blocks = (work_items + kThreads - 1) / kThreads;
    size_t cap = (size_t)sms * 8;  // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Scratch device buffers for the host-pointer variants (grown on demand, per thread).
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return RZ_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { set_error("cudaMalloc(%zu) failed", bytes); cudaGetLastError(); return RZ_ENOMEM; }
        cap = bytes;
        return RZ_OK;
    }
};
static thread_local Scratch g_scratch;

}  // namespace rz

using namespace rz;

extern "C" {

int rz_find_correct_moves_dev(const uint64_t* own, const uint64_t* enemy, uint64_t* out, size_t n, void* stream) {
    RZ_REQUIRE(n == 0 || (own && enemy && out), "rz_find_correct_moves_dev: null pointer");
    if (n == 0) return RZ_OK;
    const int vec = aligned16(own) && aligned16(enemy) && aligned16(out);
    k1_find_correct_moves<<<grid_for(vec ? (n + 1) / 2 : n), kThreads, 0, (cudaStream_t)stream>>>(own, enemy, out, n, vec);
    RZ_LAUNCH_CHECK();
    return RZ_OK;
}

int rz_calc_flip_dev(const uint8_t* pos, const uint64_t* own, const uint64_t* enemy, uint64_t* out, size_t n, void* stream) {
    RZ_REQUIRE(n == 0 || (pos && own && enemy && out), "rz_calc_flip_dev: null pointer");
    if (n == 0) return RZ_OK;
    const int vec = aligned16(own) && aligned16(enemy) && aligned16(out) && ((reinterpret_cast<uintptr_t>(pos) & 1) == 0);
    k1_calc_flip<<<grid_for(vec ? (n + 1) / 2 : n), kThreads, 0, (cudaStream_t)stream>>>(pos, own, enemy, out, n, vec);
    RZ_LAUNCH_CHECK();
    return RZ_OK;
}

int rz_step_dev(uint64_t* black, uint64_t* white, uint8_t* next_player, uint8_t* turn, uint8_t* done, uint8_t* winner,
                const int8_t* action, uint64_t* legal_out, size_t n, void* stream) {
    RZ_REQUIRE(n == 0 || (black && white && next_player && turn && done && winner && action), "rz_step_dev: null pointer");
    if (n == 0) return RZ_OK;
    k1_step<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(black, white, next_player, turn, done, winner, action, legal_out, n);
    RZ_LAUNCH_CHECK();
    return RZ_OK;
}

int rz_dihedral_dev(const uint64_t* x, const uint8_t* t, uint64_t* out, size_t n, void* stream) {
    RZ_REQUIRE(n == 0 || (x && t && out), "rz_dihedral_dev: null pointer");
    if (n == 0) return RZ_OK;
    k1_dihedral<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(x, t, out, n);
    RZ_LAUNCH_CHECK();
    return RZ_OK;
}

// ---- host-buffer variants: H2D, kernel, D2H on the default stream --------------------------------
int rz_find_correct_moves(const uint64_t* own, const uint64_t* enemy, uint64_t* out, size_t n) {
    RZ_REQUIRE(n == 0 || (own && enemy && out), "rz_find_correct_moves: null pointer");
    if (n == 0) return RZ_OK;
    const size_t b = ((n * 8 + 255) / 256) * 256;
    RZ_TRY(g_scratch.ensure(3 * b));
    char* d = (char*)g_scratch.p;
    RZ_CUDA_TRY(cudaMemcpyAsync(d, own, n * 8, cudaMemcpyHostToDevice, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(d + b, enemy, n * 8, cudaMemcpyHostToDevice, 0));
    RZ_TRY(rz_find_correct_moves_dev((u64*)d, (u64*)(d + b), (u64*)(d + 2 * b), n, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(out, d + 2 * b, n * 8, cudaMemcpyDeviceToHost, 0));
    RZ_CUDA_TRY(cudaStreamSynchronize(0));
    return RZ_OK;
}

int rz_calc_flip(const uint8_t* pos, const uint64_t* own, const uint64_t* enemy, uint64_t* out, size_t n) {
    RZ_REQUIRE(n == 0 || (pos && own && enemy && out), "rz_calc_flip: null pointer");
    if (n == 0) return RZ_OK;
    const size_t b = ((n * 8 + 255) / 256) * 256;
    RZ_TRY(g_scratch.ensure(4 * b));
    char* d = (char*)g_scratch.p;
    RZ_CUDA_TRY(cudaMemcpyAsync(d, own, n * 8, cudaMemcpyHostToDevice, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(d + b, enemy, n * 8, cudaMemcpyHostToDevice, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(d + 3 * b, pos, n, cudaMemcpyHostToDevice, 0));
    RZ_TRY(rz_calc_flip_dev((uint8_t*)(d + 3 * b), (u64*)d, (u64*)(d + b), (u64*)(d + 2 * b), n, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(out, d + 2 * b, n * 8, cudaMemcpyDeviceToHost, 0));
    RZ_CUDA_TRY(cudaStreamSynchronize(0));
    return RZ_OK;
}

int rz_step(uint64_t* black, uint64_t* white, uint8_t* next_player, uint8_t* turn, uint8_t* done, uint8_t* winner,
            const int8_t* action, uint64_t* legal_out, size_t n) {
    RZ_REQUIRE(n == 0 || (black && white && next_player && turn && done && winner && action), "rz_step: null pointer");
    if (n == 0) return RZ_OK;
    const size_t b = ((n * 8 + 255) / 256) * 256;
    RZ_TRY(g_scratch.ensure(8 * b));
    char* d = (char*)g_scratch.p;
    u64 *dB = (u64*)d, *dW = (u64*)(d + b), *dL = (u64*)(d + 2 * b);
    uint8_t *dP = (uint8_t*)(d + 3 * b), *dT = (uint8_t*)(d + 4 * b), *dD = (uint8_t*)(d + 5 * b), *dWn = (uint8_t*)(d + 6 * b);
    int8_t* dA = (int8_t*)(d + 7 * b);
    RZ_CUDA_TRY(cudaMemcpyAsync(dB, black, n * 8, cudaMemcpyHostToDevice, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(dW, white, n * 8, cudaMemcpyHostToDevice, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(dP, next_player, n, cudaMemcpyHostToDevice, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(dT, turn, n, cudaMemcpyHostToDevice, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(dD, done, n, cudaMemcpyHostToDevice, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(dWn, winner, n, cudaMemcpyHostToDevice, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(dA, action, n, cudaMemcpyHostToDevice, 0));
    RZ_TRY(rz_step_dev(dB, dW, dP, dT, dD, dWn, dA, legal_out ? dL : nullptr, n, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(black, dB, n * 8, cudaMemcpyDeviceToHost, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(white, dW, n * 8, cudaMemcpyDeviceToHost, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(next_player, dP, n, cudaMemcpyDeviceToHost, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(turn, dT, n, cudaMemcpyDeviceToHost, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(done, dD, n, cudaMemcpyDeviceToHost, 0));
    RZ_CUDA_TRY(cudaMemcpyAsync(winner, dWn, n, cudaMemcpyDeviceToHost, 0));
    if (legal_out) RZ_CUDA_TRY(cudaMemcpyAsync(legal_out, dL, n * 8, cudaMemcpyDeviceToHost, 0));
    RZ_CUDA_TRY(cudaStreamSynchronize(0));
    return RZ_OK;
}

// ---- scalar host twins (single-environment Python objects) ---------------------------------------
uint64_t rz_find_correct_moves_host(uint64_t own, uint64_t enemy) { return find_correct_moves(own, enemy); }
uint64_t rz_calc_flip_host(int pos, uint64_t own, uint64_t enemy) { return calc_flip(pos & 63, own, enemy); }
uint64_t rz_dihedral_host(uint64_t x, int t) { return dihedral(x, t & 7); }
void rz_env_reset_host(rz_env_state* s) {
    EnvState e; env_reset(e);
    s->black = e.black; s->white = e.white; s->next_player = e.next_player; s->turn = e.turn; s->done = e.done; s->winner = e.winner;
}
void rz_env_update_host(rz_env_state* s, uint64_t black, uint64_t white, int next_player) {
    EnvState e; env_update(e, black, white, next_player);
    s->black = e.black; s->white = e.white; s->next_player = e.next_player; s->turn = e.turn; s->done = e.done; s->winner = e.winner;
}
void rz_env_step_host(rz_env_state* s, int action) {
    EnvState e{s->black, s->white, s->next_player, s->turn, s->done, s->winner};
    env_step(e, action);
    s->black = e.black; s->white = e.white; s->next_player = e.next_player; s->turn = e.turn; s->done = e.done; s->winner = e.winner;
}

}  // extern "C"
