// k1_dense.cu -- K1: dense masked reduction  out[f, m] (+)= sum_k tile[f, k] * masks[m, k]
//
// Replaces ApplyMasksEngine.process_flat (reference udf/masks.py:31-83) for dense masks; the
// CoM moments (udf/com.py:577-582) and SumSigUDF (udf/sumsigudf.py:28-38) ride along as extra
// mask rows, so one pass over the frames serves all of them.
//
// The contraction is a tall-skinny GEMM (n_masks <= 24 per launch) that is HBM-bound: each
// frame element is read from HBM exactly once.  Design (sm_100a):
//   * persistent CTAs (one per SM), work item = (64-frame block, K-split)
//   * warps 0..3 = producer warpgroup (setmaxnreg.dec; one elected lane): 2D tiled cp.async.bulk.tensor loads of a [64 frames x 128 px]
//     fp32 data box (evict_first) and the matching [n_masks x 128 px] mask box (evict_last,
//     L2-resident) into a multi-stage shared-memory ring, completion on mbarriers
//   * warps 4..11 = FMA consumers (setmaxnreg.inc 232).  A lane is (fl = lane/8, q = lane%8): it owns 4 frames
//     (rows fg*16 + j*4 + fl) and 4 consecutive pixels per step.  The 8 lanes of a quarter warp
//     read one contiguous 128 B row segment (conflict-free LDS.128); mask rows are read by
//     broadcast (1 wavefront per mask per step), so shared memory traffic stays ~1x the data.
//     acc[4][NM] float2 registers (even/odd pixel sums) updated with packed fma.rn.f32x2
//     (FFMA2): exact fp32 FMA at half the issue slots and without the register-bank parity
//     conflicts scalar FFMA has on float4 components (ncu: dispatch stalls, profiles/).
//     TF32 tensor cores cannot hold the 1e-5 parity tolerance without 3x splitting.
//   * blocked accumulation: every 256 terms the per-lane partial sums are transpose-reduced
//     over the 8 pixel lanes (shuffles) into NM "total" registers, which bounds rounding error
//     growth (sequential chains <= 256 + few hundred block adds) to ~1e-6 relative.
//   * n_masks in 13..24 -> two mask groups handled by different warps on the same smem tile.
//   * OOB rows/columns (ragged F, K) are zero-filled by TMA -> no edge code in the hot loop.
// Anything the TMA path cannot take (K % 4 != 0, unaligned or strided tiles, integer or
// float64 inputs) goes through the generic kernel below (same arithmetic, no staging).
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace ltb {

constexpr int K1_FB = 64;          // frames per CTA tile
constexpr int K1_KT = 128;         // pixels per pipeline chunk (512 B per frame row)
constexpr int K1_CWARPS = 8;       // consumer warps
constexpr int K1_PWARPS = 4;       // producer warpgroup (one elected lane issues TMA)
constexpr int K1_THREADS = (K1_CWARPS + K1_PWARPS) * 32;
constexpr int K1_MAX_STAGES = 8;
constexpr int K1_CHAIN = 256;      // max sequential FMA chain per accumulator between flushes

struct K1Params {
    int64_t n_frames;
    int64_t sig_size;
    int n_masks;
    int ksplit;
    int64_t k_per_split;   // multiple of K1_KT
    int64_t n_items;
    float* out;            // (n_frames, ld_out) when ksplit == 1
    int64_t ld_out;
    float* part;           // (ksplit, n_frames, n_masks) when ksplit > 1
    int accumulate;
    int n_stages;
    float* sig_part;       // pair kernel only: (gridDim.x, sig_size) per-CTA frame sums or NULL
    uint32_t zero;         // 0, unknown to the compiler (orders the stage release after the loads)
};

__host__ __device__ constexpr size_t k1_stage_bytes(int nrows) {
    return (size_t)K1_FB * K1_KT * 4 + (size_t)nrows * K1_KT * 4;
}

// transpose-reduce step: N values on each of two partner lanes -> N/2 values each
template <int N>
__device__ __forceinline__ void xreduce_half(const float (&v)[N], float (&r)[N / 2], bool upper,
                                             int lane_xor) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        float mine = upper ? v[i + N / 2] : v[i];
        float theirs = upper ? v[i] : v[i + N / 2];
        r[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, lane_xor);
    }
}

}  // namespace ltb
#include "k1_pair.cuh"
namespace ltb {

template <int NM, int MG>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_dense_tma_kernel(const __grid_constant__ CUtensorMap tm_data,
                    const __grid_constant__ CUtensorMap tm_mask, const K1Params p) {
    constexpr int NROWS = NM * MG;                 // mask rows in a stage
    constexpr int KS = 2 / MG;                     // k-split across warps
    constexpr int KW = K1_KT / KS;                 // pixels per warp per chunk
    constexpr int SPC = KW / 32;                   // steps per chunk
    constexpr int FR = 4;                          // frames per lane
    constexpr int FLUSH_EVERY = K1_CHAIN / (2 * SPC);   // 2 terms per accumulator per step
    constexpr size_t DATA_BYTES = (size_t)K1_FB * K1_KT * 4;
    constexpr size_t STAGE_BYTES = k1_stage_bytes(NROWS);

    extern __shared__ __align__(128) uint8_t smem[];
    const int S = p.n_stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)S * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + K1_MAX_STAGES;
    float* red = reinterpret_cast<float*>(empty_bar + K1_MAX_STAGES);   // [8][NM][32]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], K1_CWARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp < K1_PWARPS) {
        // ===== producer warpgroup: give registers back, one elected lane issues TMA =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0 && lane == 0) {
            prefetch_tmap(&tm_data);
            prefetch_tmap(&tm_mask);
            const uint64_t pol_stream = l2_policy_evict_first();
            const uint64_t pol_keep = l2_policy_evict_last();
            uint32_t it = 0;
            for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int64_t fb = item / p.ksplit;
                const int ksi = (int)(item % p.ksplit);
                const int64_t k0 = (int64_t)ksi * p.k_per_split;
                int64_t k1 = k0 + p.k_per_split;
                if (k1 > p.sig_size) k1 = p.sig_size;
                const int nchunks = (int)((k1 - k0 + K1_KT - 1) / K1_KT);
                const int32_t f0 = (int32_t)(fb * K1_FB);
                for (int c = 0; c < nchunks; c++, it++) {
                    const int stage = it % S;
                    mbar_wait(&empty_bar[stage], ((it / S) & 1) ^ 1);
                    uint8_t* dst = smem + (size_t)stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)STAGE_BYTES);
                    const int32_t kc = (int32_t)(k0 + (int64_t)c * K1_KT);
                    tma_load_2d(dst, &tm_data, kc, f0, &full_bar[stage], pol_stream);
                    tma_load_2d(dst + DATA_BYTES, &tm_mask, kc, 0, &full_bar[stage], pol_keep);
                }
            }
        }
        return;
    }

    // ===== FFMA2 consumers (two warpgroups, 232 registers per thread) =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // cw = (fg, mg, ks): 16-frame group, mask group, pixel split.  lane = (fl, q).
    const int cw = warp - K1_PWARPS;
    const int fg = cw / (MG * KS);
    const int mg = (cw / KS) % MG;
    const int ks = cw % KS;
    const int fl = lane >> 3;
    const int q = lane & 7;
    const int row_base = fg * 16 + fl;             // + j*4
    const int kk_base = ks * KW + q * 4;           // + s*32

    // acc[j][m] = (sum over even pixels, sum over odd pixels): packed fma.rn.f32x2 keeps both
    // register banks busy without the parity conflicts scalar FFMA on float4 components has.
    float2 acc[FR][NM];
    float tot[NM];
    uint32_t it = 0;

    for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int64_t fb = item / p.ksplit;
        const int ksi = (int)(item % p.ksplit);
        const int64_t k0 = (int64_t)ksi * p.k_per_split;
        int64_t k1 = k0 + p.k_per_split;
        if (k1 > p.sig_size) k1 = p.sig_size;
        const int nchunks = (int)((k1 - k0 + K1_KT - 1) / K1_KT);

#pragma unroll
        for (int m = 0; m < NM; m++) {
            tot[m] = 0.f;
#pragma unroll
            for (int j = 0; j < FR; j++) acc[j][m] = make_float2(0.f, 0.f);
        }
        int since_flush = 0;

        // blocked accumulation: fold the lane partials of the 8 pixel lanes into tot[]
        // (transpose-reduce: 4*NM -> 2*NM -> NM values, then a 2-lane butterfly)
        auto flush = [&]() {
            float v[FR * NM];
#pragma unroll
            for (int j = 0; j < FR; j++)
#pragma unroll
                for (int m = 0; m < NM; m++) {
                    v[j * NM + m] = acc[j][m].x + acc[j][m].y;
                    acc[j][m] = make_float2(0.f, 0.f);
                }
            float r1[2 * NM], r2[NM];
            xreduce_half<4 * NM>(v, r1, (q & 4) != 0, 4);
            xreduce_half<2 * NM>(r1, r2, (q & 2) != 0, 2);
#pragma unroll
            for (int m = 0; m < NM; m++)
                tot[m] += r2[m] + __shfl_xor_sync(0xffffffffu, r2[m], 1);
        };

        for (int c = 0; c < nchunks; c++, it++) {
            const int stage = it % S;
            mbar_wait(&full_bar[stage], (it / S) & 1);
            const float* d = reinterpret_cast<const float*>(smem + (size_t)stage * STAGE_BYTES);
            const float* mk = d + K1_FB * K1_KT + mg * NM * K1_KT;
#pragma unroll
            for (int s = 0; s < SPC; s++) {
                const int kk = kk_base + s * 32;
                float4 dv[FR];
#pragma unroll
                for (int j = 0; j < FR; j++) dv[j] = lds128(d + (row_base + j * 4) * K1_KT + kk);
#pragma unroll
                for (int m = 0; m < NM; m++) {
                    const float4 mv = lds128(mk + m * K1_KT + kk);
                    const float2 mlo = make_float2(mv.x, mv.y), mhi = make_float2(mv.z, mv.w);
#pragma unroll
                    for (int j = 0; j < FR; j++) {
                        float2 a = acc[j][m];
                        a = __ffma2_rn(make_float2(dv[j].x, dv[j].y), mlo, a);
                        a = __ffma2_rn(make_float2(dv[j].z, dv[j].w), mhi, a);
                        acc[j][m] = a;
                    }
                }
            }
            // the stage goes back to the producer only when its loads have RETURNED: the barrier
            // address depends on the accumulator fed by the last data and mask loads (p.zero is
            // 0 at run time only) -- mbarrier.arrive does not wait for LDS in flight by itself
            // (found with K10, csrc/k10_walk.cu)
            const uint32_t dep = __float_as_uint(acc[FR - 1][NM - 1].y) & p.zero;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage] + dep);
            if (++since_flush == FLUSH_EVERY) {
                flush();
                since_flush = 0;
            }
        }
        if (since_flush) flush();

        // lanes (fl, q) and (fl, q^1) now hold frame row fg*16 + (q>>1)*4 + fl, masks
        // mg*NM .. +NM, partial over this warp's pixel share.  Combine the KS pixel-split
        // warps through shared memory.
#pragma unroll
        for (int m = 0; m < NM; m++) red[(cw * NM + m) * 32 + lane] = tot[m];
        named_bar_sync(1, K1_CWARPS * 32);
        if (ks == 0 && (q & 1) == 0) {
            const int64_t f = fb * K1_FB + fg * 16 + (q >> 1) * 4 + fl;
            if (f < p.n_frames) {
#pragma unroll
                for (int m = 0; m < NM; m++) {
                    float sum = red[(cw * NM + m) * 32 + lane];
#pragma unroll
                    for (int s2 = 1; s2 < KS; s2++)
                        sum += red[((cw + s2) * NM + m) * 32 + lane];
                    const int col = mg * NM + m;
                    if (col < p.n_masks) {
                        if (p.ksplit == 1) {
                            float* o = p.out + f * p.ld_out + col;
                            *o = p.accumulate ? (*o + sum) : sum;
                        } else {
                            p.part[((int64_t)ksi * p.n_frames + f) * p.n_masks + col] = sum;
                        }
                    }
                }
            }
        }
        named_bar_sync(1, K1_CWARPS * 32);
    }
}

// out[f, m] (+)= sum_s part[s, f, m]   (fixed order -> deterministic)
__global__ void k1_finalize_kernel(const float* __restrict__ part, int ksplit, int64_t n_frames,
                                   int n_masks, float* __restrict__ out, int64_t ld_out,
                                   int accumulate) {
    const int64_t total = n_frames * n_masks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < ksplit; k++) s += part[(int64_t)k * total + i];
        const int64_t f = i / n_masks;
        const int m = (int)(i % n_masks);
        float* o = out + f * ld_out + m;
        *o = accumulate ? (*o + s) : s;
    }
}

// ---------------------------------------------------------------------------------------
// Generic kernel: any dtype / stride / alignment.  One warp per frame, lanes along pixels
// (coalesced), NC mask columns per pass, blocked accumulation (chains of 64).
// ---------------------------------------------------------------------------------------
template <typename T, typename A, int NC>
__global__ void __launch_bounds__(256)
k1_dense_generic_kernel(const T* __restrict__ tile, int64_t n_frames, int64_t sig_size,
                        int64_t ld_tile, const A* __restrict__ masks, int n_masks,
                        int64_t ld_masks, A* __restrict__ out, int64_t ld_out, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t f = warp_global; f < n_frames; f += n_warps) {
        const T* row = tile + f * ld_tile;
        for (int m0 = 0; m0 < n_masks; m0 += NC) {
            A tot[NC], acc[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) tot[c] = acc[c] = A(0);
            int since = 0;
            for (int64_t k = lane; k < sig_size; k += 32) {
                const A d = static_cast<A>(row[k]);
#pragma unroll
                for (int c = 0; c < NC; c++)
                    if (m0 + c < n_masks) acc[c] += d * masks[(int64_t)(m0 + c) * ld_masks + k];
                if (++since == 64) {
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        tot[c] += acc[c];
                        acc[c] = A(0);
                    }
                    since = 0;
                }
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                A v = tot[c] + acc[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && m0 + c < n_masks) {
                    A* dst = out + f * ld_out + m0 + c;
                    *dst = accumulate ? (*dst + v) : v;
                }
            }
        }
    }
}

// sig_sum[k] += sum_f tile[f, k]  -- deterministic two-stage column sum (SumUDF, udf/sum.py:44-49)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ tile, int64_t n_frames, int64_t sig_size,
                      int64_t ld_tile, int frames_per_split, float* __restrict__ partial) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= sig_size) return;
    const int64_t f0 = (int64_t)blockIdx.y * frames_per_split;
    int64_t f1 = f0 + frames_per_split;
    if (f1 > n_frames) f1 = n_frames;
    float s = 0.f;
    for (int64_t f = f0; f < f1; f++) s += static_cast<float>(tile[f * ld_tile + k]);
    partial[(int64_t)blockIdx.y * sig_size + k] = s;
}

__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, int n_splits, int64_t sig_size,
                    float* __restrict__ sig_sum) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= sig_size) return;
    float s = 0.f;
    for (int i = 0; i < n_splits; i++) s += partial[(int64_t)i * sig_size + k];
    sig_sum[k] += s;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
constexpr int COLSUM_FRAMES = 512;

static int choose_ksplit(int64_t n_fb, int64_t sig_size, int sms) {
    int best = 1;
    double best_eff = 0.0;
    for (int ks = 1; ks <= 64; ks *= 2) {
        if (ks > 1 && sig_size / ks < 8 * K1_KT) break;
        const int64_t items = n_fb * ks;
        const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = ks;
        }
        if (eff >= 0.9) break;
    }
    return best;
}

static size_t k1_part_bytes(int64_t n_frames, int64_t sig_size, int n_masks) {
    const int64_t n_fb = (n_frames + K1_FB - 1) / K1_FB;
    const int ks = choose_ksplit(n_fb, sig_size, sm_count());
    return ks > 1 ? (size_t)ks * n_frames * n_masks * sizeof(float) : 0;
}

static size_t colsum_bytes(int64_t n_frames, int64_t sig_size) {
    const int64_t splits = (n_frames + COLSUM_FRAMES - 1) / COLSUM_FRAMES;
    return (size_t)splits * sig_size * sizeof(float);
}

constexpr size_t K1_SIG_SMEM_MAX = 96 * 1024;   // fused SumUDF needs sig_size*4 bytes of smem

struct WsLayout {
    size_t part_off, pack_off, sig_off, total;
};

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static WsLayout ws_layout(int64_t n_frames, int64_t sig_size, int n_masks, int with_sig) {
    WsLayout w;
    const int nm = n_masks > 24 ? 24 : n_masks;
    w.part_off = 0;
    size_t part = nm > 0 ? k1_part_bytes(n_frames, sig_size, nm) : 0;
    w.pack_off = align256(part);
    size_t pack = nm > 0 ? (size_t)24 * (((sig_size + 31) / 32) * 32) * sizeof(float) : 0;
    w.sig_off = w.pack_off + align256(pack);
    size_t sig = 0;
    if (with_sig) {
        sig = colsum_bytes(n_frames, sig_size);
        const size_t fused = (size_t)sm_count() * sig_size * sizeof(float);
        if (fused > sig) sig = fused;
    }
    w.total = w.sig_off + align256(sig);
    return w;
}

template <typename TIN, int NP, int FR, int MG>
static int launch_k1_pair(const CUtensorMap& tmd, const CUtensorMap& tmm, const K1Params& p0,
                          int grid, cudaStream_t st) {
    using C = K1PairCfg<TIN, NP, FR, MG>;
    K1Params p = p0;
    int dev = 0, smem_max = 0;
    LTB_CUDA_CHECK(cudaGetDevice(&dev));
    LTB_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t sig_bytes = p.sig_part ? (size_t)p.sig_size * 4 : 0;
    const size_t fixed = C::FIXED_BYTES + sig_bytes + 128;
    int stages = (int)(((size_t)smem_max - fixed) / C::STAGE_BYTES);
    if (stages > 6) stages = 6;
    if (const char* e = getenv("LTB200_MAX_STAGES")) {      // tuning knob
        const int cap = atoi(e);
        if (cap >= 2 && cap < stages) stages = cap;
    }
    if (stages < 2) {
        set_error("k1 pair: not enough shared memory for 2 stages");
        return LTB_ERR_UNSUPPORTED;
    }
    p.n_stages = stages;
    const size_t smem = (size_t)stages * C::STAGE_BYTES + C::FIXED_BYTES + sig_bytes;
    auto kern = k1_pair_kernel<TIN, NP, FR, MG>;
    static thread_local int configured_dev = -1;
    if (configured_dev != dev) {
        LTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            smem_max));
        configured_dev = dev;
    }
    kern<<<grid, K1_THREADS, smem, st>>>(tmd, tmm, p);
    count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

typedef int (*K1PairLauncher)(const CUtensorMap&, const CUtensorMap&, const K1Params&, int,
                              cudaStream_t);

struct PairChoice {
    K1PairLauncher fn;
    int np, fr, mg;
    int kt() const { return fr == 16 ? 256 : 128; }
};

// n_masks -> (pairs per lane, frames per lane, mask groups)
template <typename TIN>
static PairChoice pair_choice(int n_masks, bool want_fr16) {
    const int npt = (n_masks + 1) / 2;
    if (want_fr16 && npt <= 3) {
        switch (npt) {
            case 1: return {launch_k1_pair<TIN, 1, 16, 1>, 1, 16, 1};
            case 2: return {launch_k1_pair<TIN, 2, 16, 1>, 2, 16, 1};
            default: return {launch_k1_pair<TIN, 3, 16, 1>, 3, 16, 1};
        }
    }
    if (npt <= 6) {
        switch (npt) {
            case 1: return {launch_k1_pair<TIN, 1, 8, 1>, 1, 8, 1};
            case 2: return {launch_k1_pair<TIN, 2, 8, 1>, 2, 8, 1};
            case 3: return {launch_k1_pair<TIN, 3, 8, 1>, 3, 8, 1};
            case 4: return {launch_k1_pair<TIN, 4, 8, 1>, 4, 8, 1};
            case 5: return {launch_k1_pair<TIN, 5, 8, 1>, 5, 8, 1};
            default: return {launch_k1_pair<TIN, 6, 8, 1>, 6, 8, 1};
        }
    }
    const int np = (npt + 1) / 2;   // two mask groups
    switch (np) {
        case 4: return {launch_k1_pair<TIN, 4, 8, 2>, 4, 8, 2};
        case 5: return {launch_k1_pair<TIN, 5, 8, 2>, 5, 8, 2};
        default: return {launch_k1_pair<TIN, 6, 8, 2>, 6, 8, 2};
    }
}

static int g_k1_variant = -1;

// smallest column count routed to the tensor-core kernel in auto mode (LTB200_K6_MIN to tune)
static int k6_min_columns() {
    static int v = -1;
    if (v < 0) {
        v = 1;
        if (const char* e = getenv("LTB200_K6_MIN")) v = atoi(e) > 0 ? atoi(e) : 1;
    }
    return v;
}

static bool k6_u16_enabled() {
    static int v = -1;
    if (v < 0) {
        v = 1;
        if (const char* e = getenv("LTB200_K6U")) v = atoi(e) != 0;
    }
    return v != 0;
}

static int k1_variant() {
    // which FFMA2 register tile the dense path uses: 0 auto, 1 even/odd-pixel pairs ("eo"),
    // 2 mask pairs ("pair"); LTB200_K1=eo|pair|auto or ltb200_set_k1_variant()
    if (g_k1_variant < 0) {
        const char* e = getenv("LTB200_K1");
        g_k1_variant = 0;
        if (e && !strcmp(e, "eo")) g_k1_variant = 1;
        if (e && !strcmp(e, "pair")) g_k1_variant = 2;
        if (e && !strcmp(e, "tc")) g_k1_variant = 3;
    }
    return g_k1_variant;
}

template <int NM, int MG>
static int launch_k1_tma(const CUtensorMap& tmd, const CUtensorMap& tmm, const K1Params& p0,
                         int grid, cudaStream_t st) {
    K1Params p = p0;
    constexpr size_t stage = k1_stage_bytes(NM * MG);
    constexpr size_t fixed = 2 * K1_MAX_STAGES * sizeof(uint64_t) + (size_t)8 * NM * 32 * 4;
    int dev = 0, smem_max = 0;
    LTB_CUDA_CHECK(cudaGetDevice(&dev));
    LTB_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int stages = (int)(((size_t)smem_max - fixed - 128) / stage);
    if (stages > 6) stages = 6;
    if (stages < 2) {
        set_error("k1: not enough shared memory (%d B) for 2 stages of %zu B", smem_max, stage);
        return LTB_ERR_UNSUPPORTED;
    }
    p.n_stages = stages;
    const size_t smem = (size_t)stages * stage + fixed;
    auto kern = k1_dense_tma_kernel<NM, MG>;
    static thread_local int configured_dev = -1;
    if (configured_dev != dev) {
        LTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            smem_max));
        configured_dev = dev;
    }
    kern<<<grid, K1_THREADS, smem, st>>>(tmd, tmm, p);
    count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

typedef int (*K1Launcher)(const CUtensorMap&, const CUtensorMap&, const K1Params&, int,
                          cudaStream_t);

static K1Launcher k1_launcher_for(int n_masks) {
    switch (n_masks) {
        case 1: return launch_k1_tma<1, 1>;
        case 2: return launch_k1_tma<2, 1>;
        case 3: return launch_k1_tma<3, 1>;
        case 4: return launch_k1_tma<4, 1>;
        case 5: return launch_k1_tma<5, 1>;
        case 6: return launch_k1_tma<6, 1>;
        case 7: return launch_k1_tma<7, 1>;
        case 8: return launch_k1_tma<8, 1>;
        case 9: return launch_k1_tma<9, 1>;
        case 10: return launch_k1_tma<10, 1>;
        case 11: return launch_k1_tma<11, 1>;
        case 12: return launch_k1_tma<12, 1>;
        case 13: case 14: return launch_k1_tma<7, 2>;
        case 15: case 16: return launch_k1_tma<8, 2>;
        case 17: case 18: return launch_k1_tma<9, 2>;
        case 19: case 20: return launch_k1_tma<10, 2>;
        case 21: case 22: return launch_k1_tma<11, 2>;
        case 23: case 24: return launch_k1_tma<12, 2>;
        default: return nullptr;
    }
}

template <typename T, typename A>
static int launch_generic(const void* tile, int64_t F, int64_t K, int64_t ld, const A* masks,
                          int n_masks, int64_t ldm, A* out, int64_t ldo, int accumulate,
                          cudaStream_t st) {
    int64_t blocks = (F + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k1_dense_generic_kernel<T, A, 8><<<(int)blocks, 256, 0, st>>>(
        (const T*)tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate);
    count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

template <typename A>
static int dispatch_generic(const void* tile, int dtype, int64_t F, int64_t K, int64_t ld,
                            const A* masks, int n_masks, int64_t ldm, A* out, int64_t ldo,
                            int accumulate, cudaStream_t st) {
    switch (dtype) {
        case LTB_F32: return launch_generic<float, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        case LTB_U16: return launch_generic<uint16_t, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        case LTB_U8: return launch_generic<uint8_t, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        case LTB_I8: return launch_generic<int8_t, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        case LTB_I16: return launch_generic<int16_t, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        case LTB_F64: return launch_generic<double, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        case LTB_I32: return launch_generic<int32_t, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        case LTB_U32: return launch_generic<uint32_t, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        case LTB_I64: return launch_generic<int64_t, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        case LTB_U64: return launch_generic<uint64_t, A>(tile, F, K, ld, masks, n_masks, ldm, out, ldo, accumulate, st);
        default:
            set_error("masks_dense: unknown tile dtype %d", dtype);
            return LTB_ERR_ARG;
    }
}

template <typename T>
static int launch_colsum(const void* tile, int64_t F, int64_t K, int64_t ld, float* sig_sum,
                         float* ws, cudaStream_t st) {
    const int splits = (int)((F + COLSUM_FRAMES - 1) / COLSUM_FRAMES);
    dim3 grid((unsigned)((K + 255) / 256), (unsigned)splits);
    colsum_partial_kernel<T><<<grid, 256, 0, st>>>((const T*)tile, F, K, ld, COLSUM_FRAMES, ws);
    colsum_final_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(ws, splits, K, sig_sum);
    count_launch(2);
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

static size_t dtype_size(int dtype) {
    switch (dtype) {
        case LTB_U8: case LTB_I8: return 1;
        case LTB_U16: case LTB_I16: return 2;
        case LTB_F32: case LTB_I32: case LTB_U32: return 4;
        case LTB_F64: case LTB_I64: case LTB_U64: return 8;
        default: return 0;
    }
}

}  // namespace ltb

using namespace ltb;

extern "C" int ltb200_set_k1_variant(int variant) {
    LTB_REQUIRE(variant >= 0 && variant <= 3, "set_k1_variant: 0 auto, 1 eo, 2 pair, 3 tc");
    ltb::g_k1_variant = variant;
    return LTB_OK;
}

extern "C" size_t ltb200_masks_dense_workspace(int64_t n_frames, int64_t sig_size, int n_masks,
                                               int with_sig_sum) {
    if (n_frames <= 0 || sig_size <= 0 || n_masks < 0) return 0;
    size_t need = ws_layout(n_frames, sig_size, n_masks, with_sig_sum).total;
    if (n_masks > 0) {
        const size_t k6 = k6_workspace(n_frames, sig_size, n_masks);
        if (k6 > need) need = k6;
        if (n_masks <= 16) {
            const size_t k6u = k6_u16_workspace(n_frames, sig_size, n_masks, with_sig_sum);
            if (k6u > need) need = k6u;
        }
    }
    return need;
}

namespace ltb {

// one launch of a TMA-staged kernel over <= 24 mask columns; returns whether the frame sum
// (SumUDF) was fused into it
template <typename TIN>
static int run_tma_group(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
                         const float* mk, int nm, int64_t ld_masks, float* o, int64_t ld_out,
                         int accumulate, float* sig_sum, uint8_t* ws, const WsLayout& wl,
                         bool allow_eo, cudaStream_t st, bool* sig_fused) {
    const int sms = sm_count();
    const int64_t n_fb = (n_frames + K1_FB - 1) / K1_FB;
    K1Params p;
    p.zero = 0u;
    p.n_frames = n_frames;
    p.sig_size = sig_size;
    p.n_masks = nm;
    p.ksplit = choose_ksplit(n_fb, sig_size, sms);
    const int64_t chunks = (sig_size + K1_KT - 1) / K1_KT;
    p.k_per_split = ((chunks + p.ksplit - 1) / p.ksplit) * K1_KT;
    p.n_items = n_fb * p.ksplit;
    p.out = o;
    p.ld_out = ld_out;
    p.part = (float*)(ws + wl.part_off);
    p.accumulate = accumulate;
    p.n_stages = 0;
    p.sig_part = nullptr;
    const int grid = (int)(p.n_items < sms ? p.n_items : sms);
    *sig_fused = false;

    const bool want_sig = sig_sum != nullptr && (size_t)sig_size * 4 <= K1_SIG_SMEM_MAX &&
                          nm <= 6;
    const int variant = k1_variant();
    // auto: the mask-pair tile everywhere (one third less LSU wavefronts per element than the
    // even/odd tile, so it stays HBM-bound when the SM clock drops under sustained load)
    bool use_pair = !allow_eo || variant == 2 || variant == 0;
    if (variant == 1 && allow_eo) use_pair = false;

    CUtensorMap tmd, tmm;
    int rc;
    if (use_pair) {
        PairChoice pc = pair_choice<TIN>(nm, want_sig);
        const int kt = pc.kt();
        const int64_t chunks_kt = (sig_size + kt - 1) / kt;
        p.k_per_split = ((chunks_kt + p.ksplit - 1) / p.ksplit) * kt;
        rc = encode_tmap_2d(&tmd, tile, K1In<TIN>::TMAP, sizeof(TIN), (uint64_t)sig_size,
                            (uint64_t)n_frames, (uint64_t)ld_tile * sizeof(TIN), (uint32_t)kt,
                            K1_FB);
        if (rc != LTB_OK) return rc;
        const int n_pairs = pc.np * pc.mg;
        float* packed = (float*)(ws + wl.pack_off);
        const int64_t sig_pad = ((sig_size + 31) / 32) * 32;
        {
            const int64_t total = (int64_t)n_pairs * sig_pad;
            int blocks = (int)((total + 255) / 256);
            if (blocks > sms * 8) blocks = sms * 8;
            k1_pack_masks_kernel<<<blocks, 256, 0, st>>>(mk, nm, ld_masks, sig_size, sig_pad,
                                                         n_pairs, packed);
            count_launch();
        }
        rc = encode_tmap_2d(&tmm, packed, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                            (uint64_t)sig_pad * 2, (uint64_t)n_pairs, (uint64_t)sig_pad * 8,
                            256, (uint32_t)n_pairs);
        if (rc != LTB_OK) return rc;
        if (want_sig && pc.fr == 16) {
            p.sig_part = (float*)(ws + wl.sig_off);
            *sig_fused = true;
        }
        rc = pc.fn(tmd, tmm, p, grid, st);
        if (rc != LTB_OK) return rc;
        set_last_kernel(3);
    } else {
        rc = encode_tmap_2d(&tmd, tile, K1In<TIN>::TMAP, sizeof(TIN), (uint64_t)sig_size,
                            (uint64_t)n_frames, (uint64_t)ld_tile * sizeof(TIN), K1_KT, K1_FB);
        if (rc != LTB_OK) return rc;
        const int nrows = nm <= 12 ? nm : 2 * ((nm + 1) / 2);
        rc = encode_tmap_2d(&tmm, mk, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)sig_size,
                            (uint64_t)nm, (uint64_t)ld_masks * 4, K1_KT, (uint32_t)nrows);
        if (rc != LTB_OK) return rc;
        rc = k1_launcher_for(nm)(tmd, tmm, p, grid, st);
        if (rc != LTB_OK) return rc;
        set_last_kernel(1);
    }
    if (p.ksplit > 1) {
        const int64_t total = n_frames * nm;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sms * 8) blocks = sms * 8;
        k1_finalize_kernel<<<blocks, 256, 0, st>>>(p.part, p.ksplit, n_frames, nm, o, ld_out,
                                                   accumulate);
        count_launch();
    }
    if (*sig_fused) {
        colsum_final_kernel<<<(unsigned)((sig_size + 255) / 256), 256, 0, st>>>(
            p.sig_part, grid, sig_size, sig_sum);
        count_launch();
    }
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

}  // namespace ltb

extern "C" int ltb200_masks_dense(const void* tile, int tile_dtype, int64_t n_frames,
                                  int64_t sig_size, int64_t ld_tile, const float* masks,
                                  int n_masks, int64_t ld_masks, float* out, int64_t ld_out,
                                  int accumulate, float* sig_sum, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_size >= 0 && n_masks >= 0, "masks_dense: negative size");
    LTB_REQUIRE(dtype_size(tile_dtype) != 0, "masks_dense: unknown tile dtype %d", tile_dtype);
    if (n_frames == 0) return LTB_OK;
    LTB_REQUIRE(tile != nullptr || sig_size == 0, "masks_dense: tile is NULL");
    LTB_REQUIRE(ld_tile >= sig_size, "masks_dense: ld_tile %lld < sig_size %lld",
                (long long)ld_tile, (long long)sig_size);
    LTB_REQUIRE(n_masks == 0 || (masks != nullptr && out != nullptr) || sig_size == 0,
                "masks_dense: masks/out is NULL");
    LTB_REQUIRE(n_masks == 0 || (ld_masks >= sig_size && ld_out >= n_masks),
                "masks_dense: ld_masks/ld_out too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t need = ltb200_masks_dense_workspace(n_frames, sig_size, n_masks, sig_sum != nullptr);
    if (need > workspace_bytes || (need > 0 && workspace == nullptr)) {
        set_error("masks_dense: workspace of %zu B required, %zu B given", need, workspace_bytes);
        return LTB_ERR_WORKSPACE;
    }
    if (sig_size == 0) {
        // empty signal: result is all zeros
        if (!accumulate && n_masks > 0)
            LTB_CUDA_CHECK(cudaMemset2DAsync(out, ld_out * sizeof(float), 0,
                                             n_masks * sizeof(float), n_frames, st));
        return LTB_OK;
    }
    const WsLayout wl = ws_layout(n_frames, sig_size, n_masks, sig_sum != nullptr);
    uint8_t* ws = (uint8_t*)workspace;

    const size_t esz = dtype_size(tile_dtype);
    const bool tma_shape_ok = (sig_size * esz) % 16 == 0 && (ld_tile * esz) % 16 == 0 &&
                              (uintptr_t)tile % 16 == 0 && sig_size >= K1_KT && n_frames >= 8 &&
                              sig_size < (1ll << 30) && n_frames < (1ll << 31) &&
                              sig_size % 4 == 0;
    const bool tma_f32 = tile_dtype == LTB_F32 && tma_shape_ok && (ld_masks % 4 == 0) &&
                         ((uintptr_t)masks % 16 == 0);
    const bool tma_u16 = tile_dtype == LTB_U16 && tma_shape_ok;

    bool sig_done = sig_sum == nullptr;
    // float32 tiles of >= 1024 frames go to the tensor cores (K6): HBM-bound up to 32 columns
    // per pass, where the FFMA2 kernel is bound by the FP32 pipe from ~13 columns;
    // LTB200_K6_MIN raises the column threshold, variant 3 forces K6 for every shape it takes
    const bool k6_ok = tile_dtype == LTB_F32 && n_masks > 0 &&
                       k6_shape_ok(tile, n_frames, sig_size, ld_tile);
    const int variant_all = k1_variant();
    // (a frame sum the FFMA2 kernel can fuse into its single pass stays there)
    const bool k1_fuses_sig = sig_sum != nullptr && (size_t)sig_size * 4 <= K1_SIG_SMEM_MAX &&
                              n_masks <= 6;
    if (k6_ok && (variant_all == 3 || (variant_all == 0 && n_masks >= k6_min_columns() &&
                                       n_frames >= 1024 && !k1_fuses_sig))) {
        int rc = k6_run(tile, n_frames, sig_size, ld_tile, masks, n_masks, ld_masks, out, ld_out,
                        accumulate, 0, workspace, st);
        if (rc != LTB_OK) return rc;
        n_masks = 0;   // columns done; a requested frame sum still runs below
    }
    // uint16 tiles of >= 1024 frames and 7..16 columns: the same tensor-core kernel with the
    // conversion in registers and the frame sum (SumUDF) fused in (LTB200_K6U=0: FFMA2 kernel)
    // (<= 6 columns stay on the FFMA2 kernel, which is faster there and fuses the frame sum
    // too -- measured 0.68 vs 0.60 of the HBM roofline at 5 columns + sum; variant 3 forces K6)
    if (tma_u16 && n_masks > 0 && n_frames >= 1024 && (variant_all == 0 || variant_all == 3) &&
        (variant_all == 3 || n_masks > 6 || (sig_sum != nullptr && !k1_fuses_sig)) &&
        k6_u16_enabled() && k6_u16_shape_ok(tile, n_frames, sig_size, ld_tile, n_masks)) {
        int rc = k6_run_u16(tile, n_frames, sig_size, ld_tile, masks, n_masks, ld_masks, out,
                            ld_out, accumulate, 0, sig_sum, workspace, st);
        if (rc != LTB_OK) return rc;
        return LTB_OK;
    }
    // mask columns are processed in groups of <= 24 (one pass over the frames per group)
    for (int m0 = 0; m0 < n_masks; m0 += 24) {
        const int nm = (n_masks - m0) > 24 ? 24 : (n_masks - m0);
        const float* mk = masks + (int64_t)m0 * ld_masks;
        float* o = out + m0;
        bool fused = false;
        int rc;
        if (tma_f32) {
            rc = run_tma_group<float>(tile, n_frames, sig_size, ld_tile, mk, nm, ld_masks, o,
                                      ld_out, accumulate, sig_done ? nullptr : sig_sum, ws, wl,
                                      true, st, &fused);
        } else if (tma_u16) {
            rc = run_tma_group<uint16_t>(tile, n_frames, sig_size, ld_tile, mk, nm, ld_masks, o,
                                         ld_out, accumulate, sig_done ? nullptr : sig_sum, ws,
                                         wl, false, st, &fused);
        } else {
            rc = dispatch_generic<float>(tile, tile_dtype, n_frames, sig_size, ld_tile, mk, nm,
                                         ld_masks, o, ld_out, accumulate, st);
            set_last_kernel(2);
        }
        if (rc != LTB_OK) return rc;
        if (fused) sig_done = true;
    }

    if (!sig_done) {
        float* wsf = (float*)(ws + wl.sig_off);
        switch (tile_dtype) {
            case LTB_F32: return launch_colsum<float>(tile, n_frames, sig_size, ld_tile, sig_sum, wsf, st);
            case LTB_U16: return launch_colsum<uint16_t>(tile, n_frames, sig_size, ld_tile, sig_sum, wsf, st);
            case LTB_U8: return launch_colsum<uint8_t>(tile, n_frames, sig_size, ld_tile, sig_sum, wsf, st);
            case LTB_I8: return launch_colsum<int8_t>(tile, n_frames, sig_size, ld_tile, sig_sum, wsf, st);
            case LTB_I16: return launch_colsum<int16_t>(tile, n_frames, sig_size, ld_tile, sig_sum, wsf, st);
            default:
                set_error("masks_dense: sig_sum not supported for tile dtype %d (float32 path)",
                          tile_dtype);
                return LTB_ERR_UNSUPPORTED;
        }
    }
    return LTB_OK;
}

extern "C" int ltb200_masks_dense_f64(const void* tile, int tile_dtype, int64_t n_frames,
                                      int64_t sig_size, int64_t ld_tile, const double* masks,
                                      int n_masks, int64_t ld_masks, double* out, int64_t ld_out,
                                      int accumulate, void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_size >= 0 && n_masks >= 0, "masks_dense_f64: negative size");
    LTB_REQUIRE(dtype_size(tile_dtype) != 0, "masks_dense_f64: unknown tile dtype %d", tile_dtype);
    if (n_frames == 0 || n_masks == 0) return LTB_OK;
    LTB_REQUIRE(masks != nullptr && out != nullptr && (tile != nullptr || sig_size == 0),
                "masks_dense_f64: NULL pointer");
    LTB_REQUIRE(ld_tile >= sig_size && ld_masks >= sig_size && ld_out >= n_masks,
                "masks_dense_f64: leading dimension too small");
    set_last_kernel(2);
    return dispatch_generic<double>(tile, tile_dtype, n_frames, sig_size, ld_tile, masks, n_masks,
                                    ld_masks, out, ld_out, accumulate, (cudaStream_t)stream);
}
