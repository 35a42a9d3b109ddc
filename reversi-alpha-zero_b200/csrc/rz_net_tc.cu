// rz_net_tc.cu -- K4/K5: fused persistent tcgen05 residual tower + heads for the 256-filter network.
//
// One CTA per SM, each CTA owns a tile of TWO boards (M = 128 pixel rows) and carries it through the
// WHOLE network without touching HBM in between:
//   * activations live in shared memory as fp16 in the UMMA "K-major, no-swizzle" canonical layout,
//     with a one-pixel zero border, so that each of the nine 3x3 taps is just a different start
//     address of the same buffer (implicit GEMM, no im2col copy):
//         chunk(cg, slot, xp) at cg*2896 + slot*144 + xp*16 bytes   (8 channels = 16 B per chunk)
//         slot = 2*(y+1) + board (the two boards' rows are interleaved so that a dy shift is a
//         uniform 2-slot offset), xp = x+1; xp = 9 of one slot aliases xp = 0 of the next (shared
//         zero chunk).  MMA row m = (2y+board)*8 + x  ->  8-row core matrices are board rows.
//   * weights (fp16, 1.18 MB per conv, L2-resident) are streamed by a producer thread with
//     cp.async.bulk in pre-packed 32 KB stages (one tap x 64 input channels x 256 output channels)
//     through a 3-deep mbarrier ring;
//   * one thread issues tcgen05.mma (M=128, N=256, K=16, fp16 in / fp32 accumulate): 144 MMAs per conv
//     layer into a 256-column TMEM accumulator;
//   * eight epilogue warps read the accumulator with tcgen05.ld, apply the folded BatchNorm
//     (scale, shift), the residual (kept in fp32 in the other 256 TMEM columns) and ReLU, and write the
//     next layer's fp16 activations straight back into the shared-memory operand buffer;
//   * the first conv (2 -> 256 channels, K = 18 padded to 32) is a 2-MMA GEMM on an im2col tile built
//     from the two bitboards; the policy / value heads run on the epilogue warps from the fp32 tower
//     output.
// HBM traffic per position: 16 B in, 260 B out.  Algorithmic work: 2 * 755,343,616 flop.
#include <stdlib.h>
#include "rz_bitboard.cuh"
#include "rz_net.cuh"
#include "rz_tc_common.cuh"

namespace rz {
namespace tc {

constexpr int kThreads = 320;  // warp 0 producer, warp 1 MMA issuer + TMEM owner, warps 2..9 epilogue
constexpr int kEpiThreads = 256;
constexpr uint32_t kActCg = 2896, kActSlot = 144;
constexpr uint32_t kActBytes = 32 * kActCg;  // 92,672
constexpr uint32_t kStageBytes = 32768, kStages = 3;
constexpr uint32_t kA0Bytes = 8192, kW0Bytes = 16384;
constexpr uint32_t kOffAct = 0;
constexpr uint32_t kOffW = kOffAct + kActBytes;
constexpr uint32_t kOffA0 = kOffW + kStages * kStageBytes;
constexpr uint32_t kOffW0 = kOffA0 + kA0Bytes;
constexpr uint32_t kOffSS = kOffW0 + kW0Bytes;           // 2 x [scale 256][shift 256] fp32
constexpr uint32_t kOffPart = kOffSS + 2 * 2048;         // [2 halves][128 rows][4] fp32 head partial sums
constexpr uint32_t kOffHp = kOffPart + 2 * 128 * 4 * 4;  // [2 boards][128]
constexpr uint32_t kOffHv = kOffHp + 2 * 128 * 4;        // [2][64]
constexpr uint32_t kOffLogit = kOffHv + 2 * 64 * 4;      // [2][64]
constexpr uint32_t kMaxV = 512;
constexpr uint32_t kOffFc1 = kOffLogit + 2 * 64 * 4;     // [2][kMaxV]
constexpr uint32_t kOffBar = kOffFc1 + 2 * kMaxV * 4;    // mbarriers
constexpr uint32_t kNumBars = 2 * kStages + 3;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemPtr + 16;
constexpr uint32_t kSmemAlloc = kSmemBytes + 128;  // slack for manual 128 B alignment
static_assert(kSmemAlloc <= 232448, "shared memory budget exceeded");

// instruction descriptor, kind::f16: D = f32 (bits 4-5 = 1), A = B = f16 (0), K-major both,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28
constexpr uint32_t kIdesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

// CL = thread-block-cluster size (1 or 2).  With CL = 2 the two CTAs of a cluster each fetch half of every weight
// stage from L2 and multicast it into both CTAs' shared memory (L2 -> SM weight traffic halves); MMAs, TMEM and the
// epilogue stay per-CTA (cta_group::1).  A stage may be refilled only after BOTH CTAs' MMAs have read it, so the
// `empty` barriers count CL commits (each MMA thread commits to every CTA of the cluster).
// EXP != 0 are MEASUREMENT variants (RZ_TOWER_EXPERIMENT, tools/nn_bench.py; results are garbage): 1 = the epilogue only
// keeps the barrier protocol (no accumulator read-out, no BN / ReLU, no operand stores): the time of the MMA stream alone,
// i.e. what a perfect overlap of epilogue and MMA could reach; 2 = the MMA thread issues no MMAs (barriers only): the time
// of the epilogues alone.
template <int CL, int EXP = 0>
__global__ void __launch_bounds__(kThreads, 1) net_tower_kernel(const Params pp) {
    Params p = pp;
    if (p.n_dev) p.n = *p.n_dev;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = base + kOffBar;
    auto bar_full = [&](uint32_t s) { return bar0 + s * 8; };
    auto bar_empty = [&](uint32_t s) { return bar0 + (kStages + s) * 8; };
    const uint32_t bar_w0 = bar0 + 2 * kStages * 8, bar_a = bar_w0 + 8, bar_acc = bar_w0 + 16;
    const uint32_t ntiles = (p.n + 1) >> 1;
    const int L = p.n_layers;
    // every CTA of a cluster runs the same number of tile iterations (the producers / MMA threads are coupled through
    // the shared weight ring); a CTA whose last tile index is past the batch processes an all-empty dummy tile
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const uint32_t cbase = blockIdx.x - crank;
    const uint32_t iters = cbase < ntiles ? (ntiles - cbase + gridDim.x - 1) / gridDim.x : 0u;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

    // ---- one-time setup -----------------------------------------------------------------------------
    for (uint32_t i = threadIdx.x * 16; i < kActBytes; i += kThreads * 16) *reinterpret_cast<uint4*>(sm + kOffAct + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), CL); }
        mbar_init(bar_w0, 1);
        mbar_init(bar_a, kEpiThreads);
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (this kernel is the only resident CTA on its SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(base + kOffTmemPtr) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kOffTmemPtr);
    const uint32_t tm_acc = tmem, tm_res = tmem + 256;

    if (warp == 0) {
        // ===== weight producer =====================================================================
        if (lane == 0) {
            mbar_expect_tx(bar_w0, kW0Bytes);
            bulk_g2s(base + kOffW0, p.w0, kW0Bytes, bar_w0);
            uint32_t stage = 0, phase = 0;
            for (uint32_t it = 0; it < iters; ++it) {
                for (int l = 1; l < L; ++l) {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w) + (size_t)(l - 1) * 36 * kStageBytes;
                    for (int s = 0; s < 36; ++s) {
                        mbar_wait(bar_empty(stage), phase ^ 1);
                        mbar_expect_tx(bar_full(stage), kStageBytes);
                        if (CL == 1) {
                            bulk_g2s(base + kOffW + stage * kStageBytes, src + (size_t)s * kStageBytes, kStageBytes, bar_full(stage));
                        } else {  // this CTA's slice of the stage, delivered to every CTA of the cluster
                            constexpr uint32_t kSlice = kStageBytes / CL;
                            bulk_g2s_mc(base + kOffW + stage * kStageBytes + crank * kSlice, src + (size_t)s * kStageBytes + crank * kSlice,
                                        kSlice, bar_full(stage), kMask);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer ==========================================================================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, a_par = 0;
            bool first = true;
            for (uint32_t it = 0; it < iters; ++it) {
                for (int l = 0; l < L; ++l) {
                    mbar_wait(bar_a, a_par);
                    a_par ^= 1;
                    tc_fence_after();
                    if (l == 0) {
                        if (first) { mbar_wait(bar_w0, 0); first = false; }
#pragma unroll
                        for (uint32_t j = 0; j < 2; ++j)
                            umma_f16(tm_acc, smem_desc(base + kOffA0 + j * 2 * 2048, 2048, 128),
                                     smem_desc(base + kOffW0 + j * 2 * 4096, 4096, 128), kIdesc, j);
                    } else {
                        for (uint32_t tap = 0; tap < 9; ++tap) {
                            // tap (kh, kw) reads input pixel (y + kh - 1, x + kw - 1): slot offset 2*kh, chunk offset kw
                            const uint32_t a_tap = base + kOffAct + (2 * (tap / 3)) * kActSlot + (tap % 3) * 16;
                            for (uint32_t kb = 0; kb < 4; ++kb) {
                                mbar_wait(bar_full(stage), phase);
                                tc_fence_after();
                                const uint32_t b_st = base + kOffW + stage * kStageBytes;
                                if (EXP != 2) {
#pragma unroll
                                    for (uint32_t j = 0; j < 4; ++j)
                                        umma_f16(tm_acc, smem_desc(a_tap + (kb * 8 + 2 * j) * kActCg, kActCg, kActSlot),
                                                 smem_desc(b_st + 2 * j * 4096, 4096, 128), kIdesc, (tap | kb | j) != 0);
                                }
                                if (CL == 1) umma_commit(bar_empty(stage)); else umma_commit_mc(bar_empty(stage), kMask);
                                if (++stage == kStages) { stage = 0; phase ^= 1; }
                            }
                        }
                    }
                    umma_commit(bar_acc);
                }
            }
        }
    } else {
        // ===== epilogue warps (8) ==================================================================
        const int et = threadIdx.x - 64;        // 0..255
        const int q = warp & 3;                 // TMEM sub-partition this warp may access
        const int h = (warp - 2) >> 2;          // column half handled by this warp
        const int m = q * 32 + lane;            // accumulator row == TMEM lane
        const int g = m >> 3, x = m & 7, brd = g & 1, y = g >> 1;
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        float* ss_s = reinterpret_cast<float*>(sm + kOffSS);
        float* part = reinterpret_cast<float*>(sm + kOffPart);
        float* hp = reinterpret_cast<float*>(sm + kOffHp);
        float* hv = reinterpret_cast<float*>(sm + kOffHv);
        float* logit = reinterpret_cast<float*>(sm + kOffLogit);
        float* fc1 = reinterpret_cast<float*>(sm + kOffFc1);
        const uint32_t act_row = base + kOffAct + (g + 2) * kActSlot + (x + 1) * 16;  // + cg * kActCg
        uint32_t acc_par = 0;
        uint32_t ss_buf = 0;

        for (uint32_t it = 0; it < iters; ++it) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;  // may be >= ntiles: dummy tile (no valid board)
            const uint32_t pos0 = tile * 2;
            const bool valid = pos0 + brd < p.n;
            // ---- layer-0 operand: im2col of the two bit planes, K index = tap*2 + plane, padded to 32 ----
            build_layer0_operand(sm + kOffA0, valid ? p.own[pos0 + brd] : 0, valid ? p.enemy[pos0 + brd] : 0, 2 * h, g, x, y);
            fence_proxy_async();
            mbar_arrive(bar_a);

            float hp0 = 0.f, hp1 = 0.f, hvv = 0.f;
            for (int l = 0; l < L; ++l) {
                // stage this layer's folded BN parameters (double-buffered across layers)
                float* sc = ss_s + ss_buf * 512;
                sc[et] = __ldg(p.ss + (size_t)l * 512 + et);
                sc[256 + et] = __ldg(p.ss + (size_t)l * 512 + 256 + et);
                ss_buf ^= 1;
                epi_bar();
                mbar_wait(bar_acc, acc_par);
                acc_par ^= 1;
                tc_fence_after();
                if (EXP == 1) {
                    if (l != L - 1) { tc_fence_before(); mbar_arrive(bar_a); }
                    continue;
                }
                const bool is_conv2 = l > 0 && (l & 1) == 0;   // second conv of a block: add the skip connection
                const bool keep_res = l == 0 || is_conv2;      // block output: keep fp32 copy in TMEM
                const bool last = l == L - 1;
                // 4 chunks of 32 accumulator columns per thread; the TMEM load of chunk c+1 is in flight while chunk c
                // is processed (tcgen05.wait::ld only after the math and the stores of chunk c).
                uint32_t va[32], vb[32], rr[32];
                auto prefetch = [&](int c4, uint32_t (&v)[32], uint32_t (&r)[32]) {
                    const int c0 = h * 128 + c4 * 32;
                    tmem_ld32(tm_acc + lane_sel + c0, v);
                    if (is_conv2) tmem_ld32(tm_res + lane_sel + c0, r);
                };
                auto math = [&](int c4, uint32_t (&v)[32], uint32_t (&r)[32]) {
                    epi_math(v, r, sc, h * 128 + c4 * 32, is_conv2, keep_res || last);
                };
                auto store = [&](int c4, uint32_t (&v)[32]) {
                    const int c0 = h * 128 + c4 * 32;
                    if (keep_res && !last) tmem_st32(tm_res + lane_sel + c0, v);
                    if (!last) epi_store_operand(v, act_row, c0 >> 3, keep_res);
                    else epi_head_partial(v, c0, p, hp0, hp1, hvv,
                                          (p.dbg_tower && valid) ? p.dbg_tower + ((size_t)(pos0 + brd) * 64 + y * 8 + x) * 256 : nullptr);
                };
                // the residual buffer rr is consumed by math(c) before prefetch(c+1) refills it
                prefetch(0, va, rr);
                tmem_wait_ld_dep(va); if (is_conv2) tmem_dep(rr);
                math(0, va, rr); prefetch(1, vb, rr); store(0, va);
                tmem_wait_ld_dep(vb); if (is_conv2) tmem_dep(rr);
                math(1, vb, rr); prefetch(2, va, rr); store(1, vb);
                tmem_wait_ld_dep(va); if (is_conv2) tmem_dep(rr);
                math(2, va, rr); prefetch(3, vb, rr); store(2, va);
                tmem_wait_ld_dep(vb); if (is_conv2) tmem_dep(rr);
                math(3, vb, rr); store(3, vb);
                if (!last) {
                    if (keep_res) tmem_wait_st();
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(bar_a);
                }
            }
            // ---- heads (agent/model.py:43-56) on the 256 epilogue threads --------------------------------
            heads_phase(p, hp0, hp1, hvv, h, m, brd, y, x, et, warp - 2, lane, pos0, part, hp, hv, logit, fc1);
            // the next tile's layer-0 operand build only touches the A0 region, whose last reader (this tile's
            // layer-0 MMAs) completed before the first bar_acc of this tile
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it / arrive on its barriers
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- weight packing -----------------------------------------------------------------------------------
__global__ void pack_w0_kernel(const float* __restrict__ k0, __half* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over [4 kc][256 n][8 j]
    if (i >= 4 * 256 * 8) return;
    const int j = i & 7, n = (i >> 3) & 255, kc = i >> 11, k = kc * 8 + j;
    // conv0.kernel[kh][kw][c][n], K index = (kh*3+kw)*2 + c
    out[i] = __float2half_rn(k < 18 ? k0[(size_t)k * 256 + n] : 0.f);
}
__global__ void pack_w_kernel(const float* __restrict__ blob, size_t off_res0, size_t stride, int n_layers, __half* __restrict__ out) {
    const size_t total = (size_t)n_layers * 36 * 8 * 256 * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = i & 7, n = (i >> 3) & 255, kc = (i >> 11) & 7;
        const size_t ls = i >> 14;
        const int s = (int)(ls % 36), l = (int)(ls / 36);
        const int tap = s >> 2, kb = s & 3, ci = kb * 64 + kc * 8 + j;
        out[i] = __float2half_rn(blob[off_res0 + (size_t)l * stride + ((size_t)tap * 256 + ci) * 256 + n]);
    }
}

}  // namespace tc

int net_pack_tc(rz_net* net, cudaStream_t stream) {
    tc::pack_w0_kernel<<<(4 * 256 * 8 + 255) / 256, 256, 0, stream>>>(net->blob + net->off_conv0, net->tc_w0);
    if (net->cfg.res_blocks > 0)
        tc::pack_w_kernel<<<num_sms() * 8, 256, 0, stream>>>(net->blob, net->off_res0, net->res_stride_conv, 2 * net->cfg.res_blocks, net->tc_w);
    RZ_LAUNCH_CHECK();
    return RZ_OK;
}

int net_forward_tc(rz_net* net, const uint64_t* own, const uint64_t* enemy, float* policy, float* value, size_t n, cudaStream_t stream,
                   float* dbg_tower, const uint32_t* n_dev, float* dbg_logits, float* dbg_vlogit) {
    RZ_REQUIRE(net->cfg.filters == 256, "tcgen05 tower requires 256 filters");
    RZ_REQUIRE(net->cfg.value_fc <= (int)tc::kMaxV, "tcgen05 tower supports value_fc_size <= %u", tc::kMaxV);
    RZ_REQUIRE(n < (1ull << 31), "batch too large");
    static int cluster = -1, experiment = 0;
    if (cluster < 0) {
        const char* cs = getenv("RZ_TOWER_CLUSTER");
        cluster = (cs && atoi(cs) == 1) ? 1 : 2;
        const char* ex = getenv("RZ_TOWER_EXPERIMENT");
        experiment = ex ? atoi(ex) : 0;
        RZ_CUDA_TRY(cudaFuncSetAttribute(tc::net_tower_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemAlloc));
        RZ_CUDA_TRY(cudaFuncSetAttribute(tc::net_tower_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemAlloc));
        RZ_CUDA_TRY(cudaFuncSetAttribute(tc::net_tower_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemAlloc));
        RZ_CUDA_TRY(cudaFuncSetAttribute(tc::net_tower_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemAlloc));
    }
    tc::Params p;
    p.w0 = net->tc_w0; p.w = net->tc_w; p.ss = net->scale_shift; p.blob = net->blob;
    p.off_policy_conv = net->off_policy_conv; p.off_policy_fc_k = net->off_policy_fc_k; p.off_policy_fc_b = net->off_policy_fc_b;
    p.off_value_conv = net->off_value_conv; p.off_value_fc1_k = net->off_value_fc1_k; p.off_value_fc1_b = net->off_value_fc1_b;
    p.off_value_fc2_k = net->off_value_fc2_k; p.off_value_fc2_b = net->off_value_fc2_b;
    p.own = own; p.enemy = enemy; p.policy = policy; p.value = value; p.dbg_tower = dbg_tower;
    p.dbg_logits = dbg_logits; p.dbg_vlogit = dbg_vlogit;
    p.n = (uint32_t)n; p.n_dev = n_dev; p.n_layers = 1 + 2 * net->cfg.res_blocks; p.V = net->cfg.value_fc;
    const uint32_t ntiles = (uint32_t)((n + 1) / 2);
    uint32_t grid = ntiles < (uint32_t)num_sms() ? ntiles : (uint32_t)num_sms();
    if (cluster == 1) {
        tc::net_tower_kernel<1><<<grid, tc::kThreads, tc::kSmemAlloc, stream>>>(p);
    } else {
        grid = (grid + 1) & ~1u;  // whole clusters; a surplus CTA runs dummy tiles
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc::kThreads); cfg.dynamicSmemBytes = tc::kSmemAlloc; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (experiment == 1) RZ_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::net_tower_kernel<2, 1>, p));
        else if (experiment == 2) RZ_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::net_tower_kernel<2, 2>, p));
        else RZ_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc::net_tower_kernel<2>, p));
    }
    RZ_LAUNCH_CHECK();
    return RZ_OK;
}

}  // namespace rz
