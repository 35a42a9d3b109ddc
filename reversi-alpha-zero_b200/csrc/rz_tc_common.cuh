// rz_tc_common.cuh -- PTX wrappers (mbarrier, bulk copy, tcgen05 MMA / TMEM load-store / commit, cluster) and the
// parameter block shared by the two tcgen05 tower kernels (rz_net_tc.cu: one CTA per tile, cta_group::1;
// rz_net_tc2.cu: CTA pairs, cta_group::2, epilogue overlapped with the MMA stream).
#pragma once
#include <cuda_fp16.h>
#include "rz_bitboard.cuh"
#include "rz_net.cuh"

namespace rz {
namespace tc {

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait (~4 s of SM clocks): a protocol bug traps and is reported to the host instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && (spin & 1023) == 1023) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000LL) __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// multicast variants (thread-block cluster): the copy lands at the same CTA-relative offset in every CTA of
// `mask` and signals the mbarrier at the same offset there; the commit arrives on every CTA's barrier
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE; core matrix = 8 rows x 16 B (rows 16 B apart);
// LBO = byte distance between the two K-halves of one MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ULL << 46);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait::ld that also carries a register dependency on the loaded values, so the compiler cannot schedule a use of
// v[] above the wait (tcgen05.ld completes asynchronously)
__device__ __forceinline__ void tmem_wait_ld_dep(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_dep(uint32_t (&v)[32]) {
    asm volatile(""
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// two fp32 -> packed fp16x2 (a in the low half), saturating at +-65504; RELU folds max(x, 0) into the convert
template <bool RELU>
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t d;
    if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    else      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}

struct Params {
    const __half* w0;   // layer-0 weight image (16 KB)
    const __half* w;    // tower weight stages
    const float* ss;    // folded BN [L][2][256], then heads
    const float* blob;  // fp32 blob for the head weights
    size_t off_policy_conv, off_policy_fc_k, off_policy_fc_b, off_value_conv, off_value_fc1_k, off_value_fc1_b, off_value_fc2_k,
        off_value_fc2_b;
    const u64* own;
    const u64* enemy;
    float* policy;
    float* value;
    float* dbg_tower;  // nullable
    float* dbg_logits; // nullable: [n][64] policy logits (before the softmax)
    float* dbg_vlogit; // nullable: [n] value before the tanh
    uint32_t n;
    const uint32_t* n_dev;  // nullable: batch size produced on the device (engine waves)
    int n_layers;  // 1 + 2R
    int V;
};

// ---- pieces of the epilogue shared by the two tower kernels (rz_net_tc.cu, rz_net_tc2.cu) ----------------------------
constexpr uint32_t kTcActCg = 2896;   // byte distance between channel groups of 8 in the operand layout
constexpr uint32_t kTcMaxV = 512;

// layer-0 operand (agent/model.py:30-33 first convolution as a GEMM): im2col of the two bit planes of one board row m =
// (g, x), K index = tap * 2 + plane padded to 32, two of the four 8-wide K chunks (kc0, kc0 + 1) per calling thread;
// a0 = the [4 kc][16 g][8 x][8] fp16 tile in shared memory
__device__ __forceinline__ void build_layer0_operand(uint8_t* a0, u64 o, u64 e, int kc0, int g, int x, int y) {
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const int kc = kc0 + kk;
        uint32_t w[4];
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            uint32_t packed = 0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int k = kc * 8 + jp * 2 + half;
                uint32_t bit = 0;
                if (k < 18) {
                    const int tap = k >> 1, yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
                    if (yy >= 0 && yy < 8 && xx >= 0 && xx < 8) bit = (uint32_t)((((k & 1) ? e : o) >> (yy * 8 + xx)) & 1ULL);
                }
                packed |= (bit ? 0x3C00u : 0u) << (16 * half);  // fp16 1.0
            }
            w[jp] = packed;
        }
        *reinterpret_cast<uint4*>(a0 + kc * 2048 + g * 128 + x * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// folded BatchNorm (+ skip connection + ReLU) on 32 accumulator columns c0 .. c0 + 31 of one row; sc = [scale 256][shift 256].
// conv2: second convolution of a block (adds the fp32 residual r, ReLU); relu_now: block outputs / the last layer (ReLU here);
// otherwise the first convolution of a block, whose ReLU is folded into the fp16 convert of epi_store_operand
__device__ __forceinline__ void epi_math(uint32_t (&v)[32], const uint32_t (&r)[32], const float* sc, int c0, bool conv2, bool relu_now) {
    if (conv2) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            v[j] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[j]), sc[c0 + j], sc[256 + c0 + j]) + __uint_as_float(r[j]), 0.f));
    } else if (relu_now) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(fmaf(__uint_as_float(v[j]), sc[c0 + j], sc[256 + c0 + j]), 0.f));
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), sc[c0 + j], sc[256 + c0 + j]));
    }
}

// 32 fp32 values of one row -> fp16, four 16-byte chunks (8 channels each) at row_addr + (cg0 + jj) * kTcActCg
__device__ __forceinline__ void epi_store_operand(const uint32_t (&v)[32], uint32_t row_addr, int cg0, bool already_relu) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        uint4 pk;
        if (already_relu) {
            pk.x = pack_h2<false>(__uint_as_float(v[jj * 8 + 0]), __uint_as_float(v[jj * 8 + 1]));
            pk.y = pack_h2<false>(__uint_as_float(v[jj * 8 + 2]), __uint_as_float(v[jj * 8 + 3]));
            pk.z = pack_h2<false>(__uint_as_float(v[jj * 8 + 4]), __uint_as_float(v[jj * 8 + 5]));
            pk.w = pack_h2<false>(__uint_as_float(v[jj * 8 + 6]), __uint_as_float(v[jj * 8 + 7]));
        } else {
            pk.x = pack_h2<true>(__uint_as_float(v[jj * 8 + 0]), __uint_as_float(v[jj * 8 + 1]));
            pk.y = pack_h2<true>(__uint_as_float(v[jj * 8 + 2]), __uint_as_float(v[jj * 8 + 3]));
            pk.z = pack_h2<true>(__uint_as_float(v[jj * 8 + 4]), __uint_as_float(v[jj * 8 + 5]));
            pk.w = pack_h2<true>(__uint_as_float(v[jj * 8 + 6]), __uint_as_float(v[jj * 8 + 7]));
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + (cg0 + jj) * kTcActCg), "r"(pk.x), "r"(pk.y), "r"(pk.z),
                     "r"(pk.w)
                     : "memory");
    }
}

// tower output columns c0 .. c0 + 31 of one row feed the 1x1 head convolutions from registers (policy: 2 filters, value: 1)
__device__ __forceinline__ void epi_head_partial(const uint32_t (&v)[32], int c0, const Params& p, float& hp0, float& hp1, float& hvv,
                                                 float* dbg_row /* nullable: this row's 256 tower outputs */) {
    const float* wp = p.blob + p.off_policy_conv;
    const float* wv = p.blob + p.off_value_conv;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float a = __uint_as_float(v[j]);
        const float2 w2 = __ldg(reinterpret_cast<const float2*>(wp) + c0 + j);
        hp0 = fmaf(a, w2.x, hp0);
        hp1 = fmaf(a, w2.y, hp1);
        hvv = fmaf(a, __ldg(wv + c0 + j), hvv);
    }
    if (dbg_row) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dbg_row[c0 + j] = __uint_as_float(v[j]);
    }
}

// heads (agent/model.py:43-56) on the 256 epilogue threads of a CTA for its two boards: the per-thread partial sums of the
// two column halves (colhalf 0 / 1 of row m) -> BN + ReLU -> Dense(128 -> 64) + softmax, Dense(64 -> V) + ReLU -> Dense(V -> 1)
// + tanh.  part [2][128][4], hp [2][128], hv [2][64], logit [2][64], fc1 [2][kTcMaxV]: shared-memory scratch.
__device__ __forceinline__ void heads_phase(const Params& p, float hp0, float hp1, float hvv, int colhalf, int m, int brd, int y, int x,
                                            int et, int ew, int lane, uint32_t pos0, float* part, float* hp, float* hv, float* logit,
                                            float* fc1) {
    const float* ssh = p.ss + (size_t)p.n_layers * 512;
    part[(colhalf * 128 + m) * 4 + 0] = hp0;
    part[(colhalf * 128 + m) * 4 + 1] = hp1;
    part[(colhalf * 128 + m) * 4 + 2] = hvv;
    epi_bar();
    if (colhalf == 0) {
        const float a0 = part[m * 4 + 0] + part[(128 + m) * 4 + 0];
        const float a1 = part[m * 4 + 1] + part[(128 + m) * 4 + 1];
        const float av = part[m * 4 + 2] + part[(128 + m) * 4 + 2];
        const int pix = y * 8 + x;
        hp[brd * 128 + pix] = fmaxf(fmaf(a0, ssh[0], ssh[2]), 0.f);        // Flatten is (C,H,W): index c*64 + pix
        hp[brd * 128 + 64 + pix] = fmaxf(fmaf(a1, ssh[1], ssh[3]), 0.f);
        hv[brd * 64 + pix] = fmaxf(fmaf(av, ssh[4], ssh[5]), 0.f);
    }
    epi_bar();
    if (et < 128) {  // policy logits: Dense(128 -> 64)
        const int b = et >> 6, j = et & 63;
        const float* k = p.blob + p.off_policy_fc_k;
        float acc = __ldg(p.blob + p.off_policy_fc_b + j);
#pragma unroll 8
        for (int i = 0; i < 128; ++i) acc = fmaf(hp[b * 128 + i], __ldg(k + i * 64 + j), acc);
        logit[b * 64 + j] = acc;
    }
    for (int idx = et; idx < 2 * p.V; idx += 256) {  // value Dense(64 -> V) + ReLU
        const int b = idx / p.V, j = idx - b * p.V;
        const float* k = p.blob + p.off_value_fc1_k;
        float acc = __ldg(p.blob + p.off_value_fc1_b + j);
#pragma unroll 8
        for (int i = 0; i < 64; ++i) acc = fmaf(hv[b * 64 + i], __ldg(k + (size_t)i * p.V + j), acc);
        fc1[b * kTcMaxV + j] = fmaxf(acc, 0.f);
    }
    epi_bar();
    if (ew < 2) {  // softmax over 64 logits, one warp per board
        const int b = ew;
        const float l0 = logit[b * 64 + lane], l1 = logit[b * 64 + 32 + lane];
        float mx = fmaxf(l0, l1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
        float s = e0 + e1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (pos0 + b < p.n) {
            p.policy[(size_t)(pos0 + b) * 64 + lane] = e0 / s;
            p.policy[(size_t)(pos0 + b) * 64 + 32 + lane] = e1 / s;
            if (p.dbg_logits) {
                p.dbg_logits[(size_t)(pos0 + b) * 64 + lane] = l0;
                p.dbg_logits[(size_t)(pos0 + b) * 64 + 32 + lane] = l1;
            }
        }
    } else if (ew < 4) {  // value Dense(V -> 1) + tanh, one warp per board
        const int b = ew - 2;
        float acc = 0.f;
        for (int j = lane; j < p.V; j += 32) acc = fmaf(fc1[b * kTcMaxV + j], __ldg(p.blob + p.off_value_fc2_k + j), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && pos0 + b < p.n) {
            const float pre = acc + __ldg(p.blob + p.off_value_fc2_b);
            p.value[pos0 + b] = tanhf(pre);
            if (p.dbg_vlogit) p.dbg_vlogit[pos0 + b] = pre;
        }
    }
}

}  // namespace tc
}  // namespace rz
