// rz_ingest.cu -- trainer-side ingest: compact play rows -> training arrays on the device, and the
// binary row file written next to play_*.json.  Replaces the per-record Python loop of
// OptimizeWorker.convert_to_training_data (worker/optimize.py:215-231) over json.load (lib/data_helper.py:28-30).
//
// HBM-bound byte work: per row the kernel reads 280 B and writes 8 x (128 + 256 + 4) = 3104 B.  One WARP expands one row
// (no block barriers in the loop): phase A loads the 64 visit counts (two per lane, 2 x 128 B coalesced), reduces sum /
// first arg-max with shuffles and leaves the stored policy and the 16 transformed boards in shared memory; phase B writes
// the row's 8 records as 16-byte vector stores (64 for the planes, 128 for the policy through a 512-byte inverse-
// permutation table); the 8 records of a row are contiguous in every output array, so each warp store instruction covers
// 512 consecutive bytes and the loops are free of divergence.
#include <stdio.h>
#include <string.h>
#include <string>
#include "rz_bitboard.cuh"
#include "rz_common.cuh"

namespace rz {

constexpr int kIngestThreads = 256;
constexpr int kIngestWarps = kIngestThreads / 32;

__device__ __forceinline__ uint32_t bits4_to_bytes(uint32_t b) { return (b * 0x00204081u) & 0x01010101u; }  // bit i -> byte i (b < 16)

// float32(float64(n) / float64(total)) -- what the trainer feeds the model from the JSON's float64 policy.  For
// total < 2^24 both integers are exact in float32 and the correctly rounded float32 quotient is the same number: a
// double-rounding difference needs the exact quotient q within 2^-54 (relative) of a float32 midpoint m = M * 2^E (M odd,
// 25 bits) without being one; but q - m = (n * 2^-E - M * total) * 2^E / total is a non-zero integer multiple of
// 2^E / total, i.e. at least 2^-25 / total > 2^-49 relative.
__device__ __forceinline__ float visit_fraction(int n, long long total) {
    if (n == 0) return 0.f;
    if (total < (1LL << 24)) return __fdiv_rn((float)n, (float)total);
    return (float)((double)n / (double)total);
}

__global__ void __launch_bounds__(kIngestThreads, 4) ingest_kernel(const rz_play_row* __restrict__ rows, size_t n_rows, int save_tau1,
                                                                int change_tau_turn, uint8_t* __restrict__ planes,
                                                                float* __restrict__ policy, float* __restrict__ z) {
    __shared__ __align__(16) float pol_s[kIngestWarps][64];
    __shared__ u64 brd_s[kIngestWarps][8][2];
    __shared__ __align__(4) uint8_t src_s[8][64];  // src_s[t][a] = the square whose policy entry lands on square a under symmetry t
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 512; i += kIngestThreads) {
        const int t = i >> 6, a = i & 63;
        // the policy moves with the stones: out[dihedral_square(s, t)] = pol[s]  <=>  out[a] = pol[dihedral_square(a, t^-1)];
        // rotations invert to the opposite rotation, the four reflections (flip + rotation) are involutions
        src_s[t][a] = (uint8_t)dihedral_square(a, t < 4 ? (4 - t) & 3 : t);
    }
    __syncthreads();
    const size_t n_warps = (size_t)gridDim.x * kIngestWarps;
    for (size_t r = (size_t)blockIdx.x * kIngestWarps + w; r < n_rows; r += n_warps) {
        const rz_play_row* row = rows + r;
        // ---- phase A: sum and first arg-max of the 64 visit counts (np.sum / np.argmax, agent/player.py:377-385) ----
        const int n0 = row->n_visit[lane], n1 = row->n_visit[lane + 32];
        long long s = (long long)n0 + (long long)n1;
        int bn = n1 > n0 ? n1 : n0, ba = n1 > n0 ? lane + 32 : lane;
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            const int on = __shfl_xor_sync(0xffffffffu, bn, o), oa = __shfl_xor_sync(0xffffffffu, ba, o);
            if (on > bn || (on == bn && oa < ba)) { bn = on; ba = oa; }
        }
        u64 brd = 0;
        if (lane < 16) {  // the 8 symmetries of both boards, one per lane
            brd = dihedral((lane & 1) ? row->enemy : row->own, lane >> 1);
            brd_s[w][lane >> 1][lane & 1] = brd;
        }
        const u64 own = __shfl_sync(0xffffffffu, brd, 0), enemy = __shfl_sync(0xffffffffu, brd, 1);  // identity symmetry
        const int turn = popc64(own) + popc64(enemy) - 4;
        if (save_tau1 || turn < change_tau_turn) {
            pol_s[w][lane] = visit_fraction(n0, s);
            pol_s[w][lane + 32] = visit_fraction(n1, s);
        } else {
            pol_s[w][lane] = lane == ba ? 1.f : 0.f;
            pol_s[w][lane + 32] = lane + 32 == ba ? 1.f : 0.f;
        }
        __syncwarp();
        // ---- phase B: the row's 8 records are contiguous in each output array: planes 8 x 128 B = 64 sixteen-byte pieces
        //      (2 per lane), policy 8 x 256 B = 128 pieces (4 per lane); a warp store covers 512 consecutive bytes ----
#pragma unroll
        for (int j = 0; j < 2; ++j) {  // 16 squares of one plane: bit -> byte (bit_to_array, lib/bitboard.py:136-138)
            const int piece = lane + 32 * j, t = piece >> 3, part = piece & 7;
            const uint32_t bits = (uint32_t)(brd_s[w][t][part >> 2] >> (16 * (part & 3))) & 0xFFFFu;
            const uint4 v = make_uint4(bits4_to_bytes(bits & 15u), bits4_to_bytes((bits >> 4) & 15u), bits4_to_bytes((bits >> 8) & 15u),
                                       bits4_to_bytes(bits >> 12));
            __stcs(reinterpret_cast<uint4*>(planes + r * 1024) + piece, v);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int piece = lane + 32 * j, t = piece >> 4, a0 = (piece & 15) * 4;
            const uchar4 src = *reinterpret_cast<const uchar4*>(&src_s[t][a0]);
            const float4 v = make_float4(pol_s[w][src.x], pol_s[w][src.y], pol_s[w][src.z], pol_s[w][src.w]);
            __stcs(reinterpret_cast<float4*>(policy + r * 512) + piece, v);
        }
        if (lane < 2) {
            const float zf = (float)row->z;
            __stcs(reinterpret_cast<float4*>(z + r * 8 + lane * 4), make_float4(zf, zf, zf, zf));
        }
        __syncwarp();
    }
}

static const char kMagic[8] = {'R', 'Z', 'R', 'O', 'W', 'S', 0, 1};
struct RowFileHeader { char magic[8]; int32_t save_tau1, change_tau_turn; uint64_t n_rows; uint64_t zero; };
static_assert(sizeof(RowFileHeader) == 32, "row file header");
static_assert(sizeof(rz_play_row) == 280, "play row");

}  // namespace rz

using namespace rz;

extern "C" {

int rz_ingest_dev(const rz_play_row* rows, size_t n_rows, int save_policy_of_tau_1, int change_tau_turn, uint8_t* planes, float* policy,
                  float* z, void* stream) {
    RZ_REQUIRE(n_rows == 0 || (rows && planes && policy && z), "rz_ingest_dev: null pointer");
    if (n_rows == 0) return RZ_OK;
    size_t blocks = (n_rows + kIngestWarps - 1) / kIngestWarps;
    const size_t cap = (size_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    ingest_kernel<<<(unsigned)blocks, kIngestThreads, 0, (cudaStream_t)stream>>>(rows, n_rows, save_policy_of_tau_1, change_tau_turn, planes,
                                                                               policy, z);
    RZ_LAUNCH_CHECK();
    return RZ_OK;
}

int rz_ingest(const rz_play_row* rows, size_t n_rows, int save_policy_of_tau_1, int change_tau_turn, uint8_t* planes, float* policy, float* z) {
    RZ_REQUIRE(n_rows == 0 || (rows && planes && policy && z), "rz_ingest: null pointer");
    if (n_rows == 0) return RZ_OK;
    const size_t nrec = n_rows * 8;
    const size_t b_rows = ((n_rows * sizeof(rz_play_row) + 255) / 256) * 256, b_pl = nrec * 128, b_po = nrec * 64 * sizeof(float);
    char* d = nullptr;
    RZ_CUDA_TRY(cudaMalloc((void**)&d, b_rows + b_pl + b_po + nrec * sizeof(float)));
    rz_play_row* d_rows = (rz_play_row*)d;
    uint8_t* d_pl = (uint8_t*)(d + b_rows);
    float* d_po = (float*)(d + b_rows + b_pl);
    float* d_z = (float*)(d + b_rows + b_pl + b_po);
    cudaError_t ce = cudaMemcpyAsync(d_rows, rows, n_rows * sizeof(rz_play_row), cudaMemcpyHostToDevice, 0);
    int rc = RZ_OK;
    if (ce == cudaSuccess) rc = rz_ingest_dev(d_rows, n_rows, save_policy_of_tau_1, change_tau_turn, d_pl, d_po, d_z, 0);
    if (ce == cudaSuccess && rc == RZ_OK) ce = cudaMemcpyAsync(planes, d_pl, b_pl, cudaMemcpyDeviceToHost, 0);
    if (ce == cudaSuccess && rc == RZ_OK) ce = cudaMemcpyAsync(policy, d_po, b_po, cudaMemcpyDeviceToHost, 0);
    if (ce == cudaSuccess && rc == RZ_OK) ce = cudaMemcpyAsync(z, d_z, nrec * sizeof(float), cudaMemcpyDeviceToHost, 0);
    if (ce == cudaSuccess && rc == RZ_OK) ce = cudaStreamSynchronize(0);
    cudaFree(d);
    if (rc != RZ_OK) return rc;
    if (ce != cudaSuccess) { set_error("rz_ingest: %s", cudaGetErrorString(ce)); return RZ_ECUDA; }
    return RZ_OK;
}

int rz_write_play_rows(const char* path, const rz_game* games, size_t n_games, const rz_ply* plies, int save_policy_of_tau_1,
                       int change_tau_turn, size_t* n_rows) {
    RZ_REQUIRE(path && (n_games == 0 || (games && plies)), "rz_write_play_rows: null pointer");
    std::string out(sizeof(RowFileHeader), '\0');
    size_t n = 0;
    for (size_t gi = 0; gi < n_games; ++gi) {
        const rz_game& g = games[gi];
        for (int pid = 1; pid <= 2; ++pid) {  // the order of rz_write_play_data: black's plies, then white's (self_play.py:183)
            for (int i = 0; i < g.n_plies; ++i) {
                const rz_ply& pl = plies[g.first_ply + i];
                if (pl.player != pid || !pl.recorded) continue;
                rz_play_row row;
                memset(&row, 0, sizeof(row));
                row.own = pl.own; row.enemy = pl.enemy;
                memcpy(row.n_visit, pl.n_visit, sizeof(row.n_visit));
                row.z = pid == 1 ? g.black_z : -g.black_z;
                out.append((const char*)&row, sizeof(row));
                ++n;
            }
        }
    }
    RowFileHeader hd;
    memcpy(hd.magic, kMagic, 8);
    hd.save_tau1 = save_policy_of_tau_1; hd.change_tau_turn = change_tau_turn; hd.n_rows = n; hd.zero = 0;
    memcpy(&out[0], &hd, sizeof(hd));
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { set_error("rz_write_play_rows: cannot open %s", tmp.c_str()); return RZ_EIO; }
    const size_t w = fwrite(out.data(), 1, out.size(), f);
    const int cl = fclose(f);
    if (w != out.size() || cl != 0) { set_error("rz_write_play_rows: short write to %s", tmp.c_str()); remove(tmp.c_str()); return RZ_EIO; }
    if (rename(tmp.c_str(), path) != 0) { set_error("rz_write_play_rows: rename to %s failed", path); remove(tmp.c_str()); return RZ_EIO; }
    if (n_rows) *n_rows = n;
    return RZ_OK;
}

int rz_read_play_rows(const char* path, rz_play_row* rows, size_t capacity, size_t* n_rows, int* save_policy_of_tau_1, int* change_tau_turn) {
    RZ_REQUIRE(path && n_rows, "rz_read_play_rows: null pointer");
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("rz_read_play_rows: cannot open %s", path); return RZ_EIO; }
    RowFileHeader hd;
    if (fread(&hd, 1, sizeof(hd), f) != sizeof(hd) || memcmp(hd.magic, kMagic, 8) != 0) {
        fclose(f);
        set_error("rz_read_play_rows: %s is not a play-row file", path);
        return RZ_EINVAL;
    }
    *n_rows = (size_t)hd.n_rows;
    if (save_policy_of_tau_1) *save_policy_of_tau_1 = hd.save_tau1;
    if (change_tau_turn) *change_tau_turn = hd.change_tau_turn;
    int rc = RZ_OK;
    if (rows) {
        if (capacity < hd.n_rows) { set_error("rz_read_play_rows: capacity %zu < %llu rows", capacity, (unsigned long long)hd.n_rows); rc = RZ_ECAPACITY; }
        else if (fread(rows, sizeof(rz_play_row), (size_t)hd.n_rows, f) != (size_t)hd.n_rows) { set_error("rz_read_play_rows: %s is truncated", path); rc = RZ_EIO; }
    }
    fclose(f);
    return rc;
}

}  // extern "C"
