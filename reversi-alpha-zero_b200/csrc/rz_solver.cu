// rz_solver.cu -- batched endgame solver operator of the C ABI (replaces ReversiSolver.solve,
// lib/alt/reversi_solver_cython.pyx:40-61, for batches of positions).
#include "rz_common.cuh"
#include "rz_solver.cuh"

namespace rz {
namespace solver {

constexpr int kBlockThreads = 128;

// request r of a launch is served by lane (r / n_warps) of warp (r % n_warps): a batch smaller than the grid is spread over
// all warps, so each warp carries as few divergent lanes as possible
__device__ __forceinline__ uint32_t spread_index(uint32_t tid, uint32_t total) {
    const uint32_t n_warps = total >> 5;
    return (tid & 31u) * n_warps + (tid >> 5);
}

__global__ void __launch_bounds__(kBlockThreads) solve_kernel(const u64* __restrict__ own, const u64* __restrict__ enemy,
                                                              const uint8_t* __restrict__ exactly, int8_t* __restrict__ move,
                                                              int8_t* __restrict__ score, size_t n, SolveCtx* scratch, u64* tt_base) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, total = gridDim.x * blockDim.x;
    const TT tt{tt_base + (size_t)tid * kTtEntries * kTtWordsPerEntry};
    SolveCtx* c = scratch + tid;
    for (size_t i = spread_index(tid, total); i < n; i += total) {
        ctx_init(c, own[i], enemy[i], exactly[i] != 0);
        solve_advance(c, tt, 0);
        move[i] = c->move;
        score[i] = (int8_t)(c->move < 0 ? 0 : c->score);
    }
}

}  // namespace solver
}  // namespace rz

using namespace rz;

extern "C" {

int rz_solve_dev(const uint64_t* own, const uint64_t* enemy, const uint8_t* exactly, int8_t* move, int8_t* score, size_t n, void* stream) {
    RZ_REQUIRE(n == 0 || (own && enemy && exactly && move && score), "rz_solve_dev: null pointer");
    if (n == 0) return RZ_OK;
    const size_t blocks = (size_t)num_sms() * 2;
    const size_t lanes = blocks * solver::kBlockThreads;
    // per-lane request contexts and transposition tables, one set per device (kept for the life of the process; table
    // entries are position facts and never go stale).  Concurrent calls on one device must use one stream at a time.
    constexpr int kMaxDevices = 64;
    static solver::SolveCtx* scratch_of[kMaxDevices] = {};
    static u64* tt_of[kMaxDevices] = {};
    int dev = 0;
    RZ_CUDA_TRY(cudaGetDevice(&dev));
    RZ_REQUIRE(dev >= 0 && dev < kMaxDevices, "rz_solve_dev: device index out of range");
    if (!scratch_of[dev]) {
        const size_t tt_bytes = lanes * (size_t)solver::kTtEntries * solver::kTtWordsPerEntry * sizeof(u64);
        u64* t = nullptr;
        RZ_CUDA_TRY(cudaMalloc((void**)&t, tt_bytes));
        RZ_CUDA_TRY(cudaMemsetAsync(t, 0, tt_bytes, (cudaStream_t)stream));
        solver::SolveCtx* c = nullptr;
        RZ_CUDA_TRY(cudaMalloc((void**)&c, lanes * sizeof(solver::SolveCtx)));
        tt_of[dev] = t; scratch_of[dev] = c;
    }
    solver::SolveCtx* scratch = scratch_of[dev];
    u64* tt_base = tt_of[dev];
    solver::solve_kernel<<<(unsigned)blocks, solver::kBlockThreads, 0, (cudaStream_t)stream>>>(own, enemy, exactly, move, score, n, scratch, tt_base);
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
