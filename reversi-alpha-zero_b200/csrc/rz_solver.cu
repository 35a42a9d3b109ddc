// rz_solver.cu -- batched endgame solver operator of the C ABI (replaces ReversiSolver.solve,
// lib/alt/reversi_solver_cython.pyx:40-61, for batches of positions).
#include "rz_common.cuh"
#include "rz_solver.cuh"

namespace rz {
namespace solver {

constexpr int kWarpsPerBlock = 4;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) solve_kernel(const u64* __restrict__ own, const u64* __restrict__ enemy,
                                                                    const uint8_t* __restrict__ exactly, int8_t* __restrict__ move,
                                                                    int8_t* __restrict__ score, size_t n, u64* tt_base) {
    __shared__ int8_t vals[kWarpsPerBlock][kMaxTasks];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const TT tt{tt_base ? tt_base + ((size_t)blockIdx.x * kWarpsPerBlock + w) * kTtEntries * kTtWordsPerEntry : nullptr};
    for (size_t i = (size_t)blockIdx.x * kWarpsPerBlock + w; i < n; i += (size_t)gridDim.x * kWarpsPerBlock) {
        int mv, sc;
        solve_warp(own[i], enemy[i], exactly[i] != 0, vals[w], lane, mv, sc, tt);
        if (lane == 0) { move[i] = (int8_t)mv; score[i] = (int8_t)(mv < 0 ? 0 : sc); }
        __syncwarp();
    }
}

}  // namespace solver
}  // namespace rz

using namespace rz;

extern "C" {

int rz_solve_dev(const uint64_t* own, const uint64_t* enemy, const uint8_t* exactly, int8_t* move, int8_t* score, size_t n, void* stream) {
    RZ_REQUIRE(n == 0 || (own && enemy && exactly && move && score), "rz_solve_dev: null pointer");
    if (n == 0) return RZ_OK;
    size_t blocks = (n + solver::kWarpsPerBlock - 1) / solver::kWarpsPerBlock;
    const size_t cap = (size_t)num_sms() * 4;
    if (blocks > cap) blocks = cap;
    // per-warp transposition tables (kept for the life of the process; entries are position facts and never go stale)
    static u64* tt_base = nullptr;
    static size_t tt_warps = 0;
    if (tt_warps < cap * solver::kWarpsPerBlock) {
        if (tt_base) cudaFree(tt_base);
        tt_base = nullptr; tt_warps = 0;
        const size_t bytes = cap * solver::kWarpsPerBlock * (size_t)solver::kTtEntries * solver::kTtWordsPerEntry * sizeof(u64);
        RZ_CUDA_TRY(cudaMalloc((void**)&tt_base, bytes));
        RZ_CUDA_TRY(cudaMemsetAsync(tt_base, 0, bytes, (cudaStream_t)stream));
        tt_warps = cap * solver::kWarpsPerBlock;
    }
    solver::solve_kernel<<<(unsigned)blocks, solver::kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(own, enemy, exactly, move, score, n, tt_base);
    RZ_LAUNCH_CHECK();
    return RZ_OK;
}

int rz_solve(const uint64_t* own, const uint64_t* enemy, const uint8_t* exactly, int8_t* move, int8_t* score, size_t n) {
    RZ_REQUIRE(n == 0 || (own && enemy && exactly && move && score), "rz_solve: null pointer");
    if (n == 0) return RZ_OK;
    char* d = nullptr;
    const size_t b = ((n * 8 + 255) / 256) * 256;
    RZ_CUDA_TRY(cudaMalloc((void**)&d, 2 * b + 3 * ((n + 255) / 256) * 256));
    u64 *d_own = (u64*)d, *d_en = (u64*)(d + b);
    uint8_t* d_ex = (uint8_t*)(d + 2 * b);
    int8_t* d_mv = (int8_t*)(d_ex + ((n + 255) / 256) * 256);
    int8_t* d_sc = d_mv + ((n + 255) / 256) * 256;
    cudaError_t ce = cudaMemcpyAsync(d_own, own, n * 8, cudaMemcpyHostToDevice, 0);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_en, enemy, n * 8, cudaMemcpyHostToDevice, 0);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_ex, exactly, n, cudaMemcpyHostToDevice, 0);
    int rc = RZ_OK;
    if (ce == cudaSuccess) rc = rz_solve_dev(d_own, d_en, d_ex, d_mv, d_sc, n, 0);
    if (ce == cudaSuccess && rc == RZ_OK) ce = cudaMemcpyAsync(move, d_mv, n, cudaMemcpyDeviceToHost, 0);
    if (ce == cudaSuccess && rc == RZ_OK) ce = cudaMemcpyAsync(score, d_sc, n, cudaMemcpyDeviceToHost, 0);
    if (ce == cudaSuccess && rc == RZ_OK) ce = cudaStreamSynchronize(0);
    cudaFree(d);
    if (rc != RZ_OK) return rc;
    if (ce != cudaSuccess) { set_error("rz_solve: %s", cudaGetErrorString(ce)); return RZ_ECUDA; }
    return RZ_OK;
}

}  // extern "C"
