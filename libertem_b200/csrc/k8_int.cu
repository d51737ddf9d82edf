// k8_int.cu -- K8: integer masked reduction of uint16 tiles on the int8 tensor cores
//
//     out[f, m] (+)= sum_k tile[f, k] * masks[m, k]     tile uint16, masks int8 (binary / small
//     sig_sum[k]  += sum_f tile[f, k]                    integer weights), exact integer sums
//
// Same seam as K1/K6 (ApplyMasksEngine.process_flat, reference udf/masks.py:31-83, with the
// frame sum of SumUDF, udf/sum.py:44-49, fused in) for the most common detector case: integer
// counts (uint16) against binary virtual-detector masks (ring / disk / all-ones for SumSigUDF).
// There the float32 arithmetic of the reference is exact, so integer arithmetic reproduces it
// bit for bit -- and it needs NO per-pixel instruction at all:
//
//   * a uint16 pixel is two bytes.  The TMA'd stage [256 frames x 64 px] is, byte-wise, a
//     K-major [256 x 128] uint8 matrix -- a valid A operand of tcgen05.mma.kind::i8 as it
//     lands in shared memory.  The masks are packed once per call into int8 rows over the BYTE
//     index: row c holds m[c, p] at byte 2p (weights of the low bytes), row NC + c holds it at
//     byte 2p + 1 (weights of the high bytes).  The accumulators (int32 in TMEM) then hold
//     L[f, c] = sum lowbyte * m and H[f, c] = sum highbyte * m, and out = L + 256 H.
//   * the frame sum is a second MMA on the same stage: D2[., n] = sum_f ones[., f] * tile[f, n]
//     with the stage read as an MN-major [K = 256 frames x N = 128 bytes] B operand and an
//     all-ones A operand held in TMEM; every row of D2 is the per-byte column sum.
//
// uint8 tiles (BPP = 1) are the same kernel with one byte per pixel: 128 pixels per stage, one
// int8 mask row per column, out = L.
//
// So the SM only moves data: TMA -> shared memory -> tensor core.  Warps: 0 frame-stream TMA,
// 1 mask-tile TMA, 2 MMA issuer (warp-uniform loop, one elected lane), 3 TMEM allocation,
// 4..7 drain (per stage: the 128 byte sums -> 64 pixel sums -> 64-bit RED into the frame-sum
// accumulator; per item: the (256 frames x columns) block).  int32 accumulation is exact:
// |L|, |H| <= 255 * 127 * 65536 < 2^31 per K split of <= 65536 pixels (larger signals are split on
// the host side of the call and recombined as int64).
#include "common.cuh"
#include <cstdlib>

namespace ltb {

constexpr int K8_FB = 256;             // frames per item (2 groups of 128 TMEM lanes)
constexpr int K8_DS = 5;               // data ring depth
constexpr int K8_MS = 4;               // mask ring depth
constexpr int K8_THREADS = 256;
constexpr int K8_DRAIN_WARPS = 4;
constexpr uint32_t K8_STAGE_BYTES = K8_FB * 128;          // 32 KiB
constexpr int K8_TMEM_COLS = 512;
constexpr int K8_ONES_COL = 128;       // TMEM columns [128, 136): all-ones A operand (32 int8 / row)
constexpr int K8_SUM_COL = 256;        // TMEM columns [256, 512): 2 x 128 byte-sum accumulators
constexpr int K8_MAX_COLUMNS = 32;     // rows 17..32: N = 64 (uint16) / N = 32 (uint8) with one
                                       // accumulator buffer per item (tests/test_k8_gpu.py)

struct K8Params {
    int64_t n_frames;
    int64_t sig_size;
    int n_masks;
    int ksplit;
    int64_t k_per_split;   // pixels, multiple of the stage width (128 / BPP)
    int64_t n_items;
    float* out;
    int64_t ld_out;
    long long* part;       // (ksplit, n_frames, n_masks) exact partial sums when ksplit > 1
    int accumulate;
    unsigned long long* sig_acc;   // (sig_size) exact frame sums, or NULL
    int debug;             // LTB200_K8_DEBUG: 1 no mask MMAs, 2 no frame-sum MMAs
};

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ void k8_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void k8_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void k8_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 operands, int32 accumulate
__device__ __forceinline__ void k8_mma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void k8_mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void k8_st8(uint32_t taddr, uint32_t v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(
            taddr),
        "r"(v)
        : "memory");
}
__device__ __forceinline__ void k8_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, "
        "%11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void k8_ld_fence16(uint32_t (&r)[16]) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]),
                   "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]),
                   "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void k8_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void k8_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool k8_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// shared-memory matrix descriptors, 128-byte swizzle, 8-row groups 1024 bytes apart
// (bits: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=2 [61,64)).
// K-major: a row is 128 contiguous bytes of K, rows are M/N.  MN-major: a row is 128 contiguous
// bytes of N, rows are K (8 K rows per swizzle atom, atoms SBO apart).
__device__ __forceinline__ uint64_t k8_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major operand (LBO field = 1 as for every swizzled K-major layout)
__device__ __forceinline__ uint64_t k8_desc_k(uint32_t smem_addr) {
    return k8_desc_sw128(smem_addr, 16);
}
// MN-major operand of one 128-byte atom along N (LBO = distance to the next atom, unused here)
__device__ __forceinline__ uint64_t k8_desc_mn(uint32_t smem_addr) {
    return k8_desc_sw128(smem_addr, 1024);
}

// instruction descriptor of kind::i8: D = S32 (2 << 4), A format (0 = u8, 1 = s8) << 7,
// B format << 10, B MN-major << 16, N >> 3 << 17, M >> 4 << 24
__host__ __device__ constexpr uint32_t k8_idesc(int a_signed, int b_signed, int b_mn_major, int n) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}

// uint16 tiles: packed[r, 2p + b] (int8, (2 NC) x (2 sig_pad)): r < NC -> mask r weights the low
// bytes, r >= NC -> mask r - NC weights the high bytes; zero padded.
__global__ void k8_pack_masks_kernel(const int8_t* __restrict__ masks, int n_masks,
                                     int64_t ld_masks, int64_t sig_size, int64_t sig_pad, int nc,
                                     uint16_t* __restrict__ packed) {
    const int64_t total = (int64_t)nc * sig_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / sig_pad);
        const int64_t k = i % sig_pad;
        uint16_t m = 0;
        if (r < n_masks && k < sig_size) m = (uint8_t)masks[(int64_t)r * ld_masks + k];
        packed[(int64_t)r * sig_pad + k] = m;
        packed[(int64_t)(r + nc) * sig_pad + k] = (uint16_t)(m << 8);
    }
}

// uint8 tiles: packed[r, p] (int8, N x sig_pad) = mask r, zero padded
__global__ void k8_pack_masks_u8_kernel(const int8_t* __restrict__ masks, int n_masks,
                                        int64_t ld_masks, int64_t sig_size, int64_t sig_pad,
                                        int n, int8_t* __restrict__ packed) {
    const int64_t total = (int64_t)n * sig_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / sig_pad);
        const int64_t k = i % sig_pad;
        packed[i] = (r < n_masks && k < sig_size) ? masks[(int64_t)r * ld_masks + k] : (int8_t)0;
    }
}

__global__ void k8_finalize_kernel(const long long* __restrict__ part, int ksplit,
                                   int64_t n_frames,
                                   int n_masks, float* __restrict__ out, int64_t ld_out,
                                   int accumulate) {
    const int64_t total = n_frames * n_masks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        long long t = 0;
        for (int k = 0; k < ksplit; k++) t += part[(int64_t)k * total + i];
        const float s = (float)t;
        float* o = out + (i / n_masks) * ld_out + (i % n_masks);
        *o = accumulate ? (*o + s) : s;
    }
}

__global__ void k8_sig_finalize_kernel(const unsigned long long* __restrict__ acc,
                                       int64_t sig_size, float* __restrict__ sig_sum) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < sig_size) sig_sum[k] += (float)acc[k];
}

struct K8Smem {
    static constexpr uint32_t data_off(int s) { return (uint32_t)s * K8_STAGE_BYTES; }
    static constexpr uint32_t mask_off(int s, int n) {
        return K8_DS * K8_STAGE_BYTES + (uint32_t)s * (uint32_t)n * 128u;
    }
    static constexpr uint32_t bar_off(int n) { return mask_off(K8_MS, n); }
    static constexpr uint32_t total(int n) { return bar_off(n) + 256 + 1024; }
};

// N accumulator columns per frame group: BPP = 2 (uint16): N = 2 NC, NC = 8 or 16 mask columns
// (low-byte and high-byte sums); BPP = 1 (uint8): N = 16 mask columns
template <int N, int BPP>
__global__ void __launch_bounds__(K8_THREADS, 1)
k8_int_kernel(const __grid_constant__ CUtensorMap tm_data,
              const __grid_constant__ CUtensorMap tm_mask, const K8Params p) {
    constexpr int NC = N / BPP;
    constexpr int K8_PX = 128 / BPP;                            // pixels per stage
    constexpr uint32_t MASK_BYTES = (uint32_t)N * 128u;
    constexpr uint32_t IDESC_MASK = k8_idesc(0, 1, 0, N);       // u8 data x s8 masks
    constexpr uint32_t IDESC_SUM = k8_idesc(0, 0, 1, 128);      // u8 ones x u8 data (MN-major)
    static_assert(N == 16 || N == 32 || N == 64, "K8: N in {16, 32, 64}");
    // item accumulators: 2 frame groups x N columns, double-buffered across items while they fit
    // in TMEM columns [0, 128); N = 64 keeps ONE buffer (the MMA warp waits for the drain at the
    // item boundary, once per >= 64 stages)
    constexpr bool ONE_ACC = N == 64;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K8Smem::bar_off(N));
    uint64_t* data_full = bars;                       // [DS]  TMA landed
    uint64_t* data_free = data_full + K8_DS;          // [DS]  MMAs that read the stage completed
    uint64_t* mask_full = data_free + K8_DS;          // [MS]
    uint64_t* mask_empty = mask_full + K8_MS;         // [MS]
    uint64_t* sum_full = mask_empty + K8_MS;          // [2]   byte sums of a stage complete
    uint64_t* sum_free = sum_full + 2;                // [2]   drained
    uint64_t* acc_full = sum_free + 2;                // [2]   item accumulators complete
    uint64_t* acc_free = acc_full + 2;                // [2]   drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const bool with_sum = p.sig_acc != nullptr;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K8_DS; s++) {
            mbar_init(&data_full[s], 1);
            mbar_init(&data_free[s], 1);
        }
        for (int s = 0; s < K8_MS; s++) {
            mbar_init(&mask_full[s], 1);
            mbar_init(&mask_empty[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(&sum_full[s], 1);
            mbar_init(&sum_free[s], K8_DRAIN_WARPS);
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_free[s], K8_DRAIN_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "n"(K8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    k8_fence_before();
    __syncthreads();
    k8_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 4) {
        // the all-ones A operand of the frame-sum MMAs: 32 int8 ones per TMEM lane
        const uint32_t lane_sel = (uint32_t)((warp - 4) * 32) << 16;
        k8_st8(tmem_base + lane_sel + K8_ONES_COL, 0x01010101u);
        k8_wait_st();
    }
    k8_fence_before();
    __syncthreads();
    k8_fence_after();

    auto item_range = [&](int64_t item, int64_t& fb, int& ksi, int64_t& k0, int& n_sub) {
        fb = item / p.ksplit;
        ksi = (int)(item % p.ksplit);
        k0 = (int64_t)ksi * p.k_per_split;
        int64_t k1 = k0 + p.k_per_split;
        if (k1 > p.sig_size) k1 = p.sig_size;
        n_sub = (int)((k1 - k0 + K8_PX - 1) / K8_PX);
    };

    if (warp == 0) {
        // ===== frame stream producer =====
        if (lane == 0) {
            prefetch_tmap(&tm_data);
            const uint64_t pol = l2_policy_evict_first();
            uint32_t it = 0;
            for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int64_t fb, k0;
                int ksi, n_sub;
                item_range(item, fb, ksi, k0, n_sub);
                const int32_t f0 = (int32_t)(fb * K8_FB);
                for (int i = 0; i < n_sub; i++, it++) {
                    const int ds = it % K8_DS;
                    mbar_wait(&data_free[ds], ((it / K8_DS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&data_full[ds], K8_STAGE_BYTES);
                    tma_load_2d(smem + K8Smem::data_off(ds), &tm_data, (int32_t)(k0 + i * K8_PX),
                                f0, &data_full[ds], pol);
                }
            }
        }
    } else if (warp == 1) {
        // ===== mask tile producer (byte-expanded int8 rows, L2-resident) =====
        if (lane == 0) {
            prefetch_tmap(&tm_mask);
            const uint64_t pol = l2_policy_evict_last();
            uint32_t it = 0;
            for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                int64_t fb, k0;
                int ksi, n_sub;
                item_range(item, fb, ksi, k0, n_sub);
                for (int i = 0; i < n_sub; i++, it++) {
                    const int ms = it % K8_MS;
                    mbar_wait(&mask_empty[ms], ((it / K8_MS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&mask_full[ms], MASK_BYTES);
                    tma_load_2d(smem + K8Smem::mask_off(ms, N), &tm_mask,
                                (int32_t)(BPP * (k0 + i * K8_PX)), 0, &mask_full[ms], pol);
                }
            }
        }
    } else if (warp == 2) {
        // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
        uint32_t it = 0, item_n = 0;
        for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x, item_n++) {
            int64_t fb, k0;
            int ksi, n_sub;
            item_range(item, fb, ksi, k0, n_sub);
            const uint32_t ab = ONE_ACC ? 0u : (item_n & 1);
            mbar_wait(&acc_free[ab], ONE_ACC ? ((item_n & 1) ^ 1) : (((item_n >> 1) & 1) ^ 1));
            for (int i = 0; i < n_sub; i++, it++) {
                const int ds = it % K8_DS;
                const int ms = it % K8_MS;
                const uint32_t sb = it & 1;
                mbar_wait(&mask_full[ms], (it / K8_MS) & 1);
                mbar_wait(&data_full[ds], (it / K8_DS) & 1);
                if (with_sum) mbar_wait(&sum_free[sb], ((it >> 1) & 1) ^ 1);
                k8_fence_after();
                const uint32_t data_addr = smem_u32(smem + K8Smem::data_off(ds));
                const uint64_t mdesc0 = k8_desc_k(smem_u32(smem + K8Smem::mask_off(ms, N)));
                const uint32_t d0 = tmem_base + ab * 2 * N;
                if (k8_elect_one()) {
                    if (!(p.debug & 1)) {
#pragma unroll
                        for (int g = 0; g < 2; g++) {
                            const uint64_t adesc0 = k8_desc_k(data_addr + g * 128 * 128);
#pragma unroll
                            for (int kk = 0; kk < 4; kk++)     // 32 bytes of K per MMA
                                k8_mma_i8_ss(d0 + g * N, adesc0 + (uint64_t)(kk * 2),
                                             mdesc0 + (uint64_t)(kk * 2), IDESC_MASK,
                                             (i | kk) != 0 ? 1u : 0u);
                        }
                    }
                    if (with_sum && !(p.debug & 2)) {
#pragma unroll
                        for (int ks = 0; ks < K8_FB / 32; ks++)   // 32 frames of K per MMA
                            k8_mma_i8_ts(tmem_base + K8_SUM_COL + sb * 128,
                                         tmem_base + K8_ONES_COL,
                                         k8_desc_mn(data_addr + ks * 32 * 128), IDESC_SUM,
                                         ks != 0 ? 1u : 0u);
                    }
                    k8_commit(&data_free[ds]);
                    k8_commit(&mask_empty[ms]);
                    if (with_sum) k8_commit(&sum_full[sb]);
                    if (i == n_sub - 1) k8_commit(&acc_full[ab]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== drain: per stage the byte sums, per item the accumulators =====
        const int w = warp - 4;                      // TMEM lane quarter (== warp % 4)
        const uint32_t lane_sel = (uint32_t)(w * 32) << 16;
        uint32_t it = 0, item_n = 0;
        for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x, item_n++) {
            int64_t fb, k0;
            int ksi, n_sub;
            item_range(item, fb, ksi, k0, n_sub);
            if (with_sum) {
                for (int i = 0; i < n_sub; i++, it++) {
                    const uint32_t sb = it & 1;
                    mbar_wait(&sum_full[sb], (it >> 1) & 1);
                    k8_fence_after();
                    // every row of D2 holds the 128 byte sums; this warp takes bytes
                    // [32 w, 32 w + 32) = pixels [16 w, 16 w + 16) of the stage
                    uint32_t r[2][16];
                    const uint32_t a = tmem_base + lane_sel + K8_SUM_COL + sb * 128 + w * 32;
                    k8_ld16(a, r[0]);
                    k8_ld16(a + 16, r[1]);
                    k8_wait_ld();
                    k8_ld_fence16(r[0]);
                    k8_ld_fence16(r[1]);
                    k8_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&sum_free[sb]);
                        if constexpr (BPP == 2) {
                            const int64_t px0 = k0 + (int64_t)i * K8_PX + w * 16;
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                const uint32_t lo = r[j >> 3][(2 * j) & 15];
                                const uint32_t hi = r[j >> 3][(2 * j + 1) & 15];
                                if (px0 + j < p.sig_size)
                                    atomicAdd(p.sig_acc + px0 + j,
                                              (unsigned long long)(lo + (hi << 8)));
                            }
                        } else {
                            const int64_t px0 = k0 + (int64_t)i * K8_PX + w * 32;
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (px0 + j < p.sig_size)
                                    atomicAdd(p.sig_acc + px0 + j,
                                              (unsigned long long)r[j >> 4][j & 15]);
                        }
                    }
                }
            }
            // item accumulators
            if constexpr (!ONE_ACC) {
                const uint32_t ab = item_n & 1;
                mbar_wait(&acc_full[ab], (item_n >> 1) & 1);
                k8_fence_after();
                uint32_t v[2][N / 16][16];
    #pragma unroll
                for (int g = 0; g < 2; g++)
    #pragma unroll
                    for (int q = 0; q < N / 16; q++)
                        k8_ld16(tmem_base + lane_sel + ab * 2 * N + g * N + q * 16, v[g][q]);
                k8_wait_ld();
    #pragma unroll
                for (int g = 0; g < 2; g++)
    #pragma unroll
                    for (int q = 0; q < N / 16; q++) k8_ld_fence16(v[g][q]);
                k8_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_free[ab]);
    #pragma unroll
                for (int g = 0; g < 2; g++) {
                    const int64_t f = fb * K8_FB + g * 128 + w * 32 + lane;
                    if (f >= p.n_frames) continue;
                    float* o = p.out + f * p.ld_out;
                    long long* po = p.part + ((int64_t)ksi * p.n_frames + f) * p.n_masks;
    #pragma unroll
                    for (int c = 0; c < NC; c++) {
                        if (c >= p.n_masks) break;
                        long long tot = (long long)(int32_t)v[g][c / 16][c % 16];
                        if constexpr (BPP == 2)
                            tot += 256ll * (int32_t)v[g][(NC + c) / 16][(NC + c) % 16];
                        if (p.ksplit == 1) {
                            const float val = (float)tot;
                            o[c] = p.accumulate ? (o[c] + val) : val;
                        } else {
                            po[c] = tot;
                        }
                    }
                }
            } else {
                // 64 columns per group, one accumulator buffer: one group at a time keeps the
                // drain at 64 registers; the buffer is handed back after the second load
                mbar_wait(&acc_full[0], item_n & 1);
                k8_fence_after();
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    uint32_t v[N / 16][16];
#pragma unroll
                    for (int q = 0; q < N / 16; q++)
                        k8_ld16(tmem_base + lane_sel + g * N + q * 16, v[q]);
                    k8_wait_ld();
#pragma unroll
                    for (int q = 0; q < N / 16; q++) k8_ld_fence16(v[q]);
                    if (g == 1) {
                        k8_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_free[0]);
                    }
                    const int64_t f = fb * K8_FB + g * 128 + w * 32 + lane;
                    if (f >= p.n_frames) continue;
                    float* o = p.out + f * p.ld_out;
                    long long* po = p.part + ((int64_t)ksi * p.n_frames + f) * p.n_masks;
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        if (c >= p.n_masks) break;
                        long long tot = (long long)(int32_t)v[c / 16][c % 16];
                        if constexpr (BPP == 2)
                            tot += 256ll * (int32_t)v[(NC + c) / 16][(NC + c) % 16];
                        if (p.ksplit == 1) {
                            const float val = (float)tot;
                            o[c] = p.accumulate ? (o[c] + val) : val;
                        } else {
                            po[c] = tot;
                        }
                    }
                }
            }
        }
    }

    k8_fence_before();
    __syncthreads();
    if (warp == 3) {
        k8_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(K8_TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// K split: enough items for the SMs, and never more than K8_MAX_SPLIT_PX pixels per split so
// that the int32 accumulators stay exact (255 * 127 * 65536 < 2^31); the splits are recombined
// as int64 (k8_finalize_kernel)
constexpr int64_t K8_MAX_SPLIT_PX = 65536;
constexpr int64_t K8_MAX_SIG = K8_MAX_SPLIT_PX * 64;

static int k8_choose_ksplit(int64_t n_fb, int64_t sig_size, int sms, int px) {
    int ks0 = 1;
    while ((int64_t)ks0 * K8_MAX_SPLIT_PX < sig_size) ks0 *= 2;      // exactness: forced splits
    int best = ks0;
    double best_eff = 0.0;
    for (int ks = ks0; ks <= 64; ks *= 2) {
        if (ks > ks0 && sig_size / ks < 16 * px) break;
        const int64_t items = n_fb * ks;
        const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = ks;
        }
        if (eff >= 0.95) break;
    }
    // every split must own at least one stage: with `per` stages per split only
    // ceil(stages / per) splits are non-empty (an empty split would never commit its
    // accumulator and the drain warps would wait for it forever)
    const int64_t subs = (sig_size + px - 1) / px;
    const int64_t per = (subs + best - 1) / best;
    return (int)((subs + per - 1) / per);
}

static size_t k8_align256(size_t v) { return (v + 255) & ~(size_t)255; }

// accumulator columns N of the kernel instantiation
static int k8_n(int n_masks, int bpp) {
    if (bpp == 2) return n_masks <= 8 ? 16 : n_masks <= 16 ? 32 : 64;
    return n_masks <= 16 ? 16 : 32;
}

struct K8Ws {
    size_t pack_off, part_off, sig_off, total;
};

static K8Ws k8_ws(int64_t n_frames, int64_t sig_size, int n_masks, int bpp, bool with_sig) {
    K8Ws w;
    const int px = 128 / bpp;
    const int64_t sig_pad = ((sig_size + px - 1) / px) * px;
    w.pack_off = 0;
    const size_t pack = (size_t)k8_n(n_masks, bpp) * sig_pad * bpp;
    w.part_off = k8_align256(pack);
    const int64_t n_fb = (n_frames + K8_FB - 1) / K8_FB;
    const int ks = k8_choose_ksplit(n_fb, sig_size, sm_count(), px);
    const size_t part = ks > 1 ? (size_t)ks * n_frames * n_masks * sizeof(long long) : 0;
    w.sig_off = w.part_off + k8_align256(part);
    w.total = w.sig_off + (with_sig ? k8_align256((size_t)sig_size * 8) : 0);
    return w;
}

template <int N, int BPP>
static int k8_launch(const CUtensorMap& tmd, const CUtensorMap& tmm, const K8Params& p, int grid,
                     cudaStream_t st) {
    auto kern = k8_int_kernel<N, BPP>;
    const size_t smem = K8Smem::total(N);
    int dev = 0;
    LTB_CUDA_CHECK(cudaGetDevice(&dev));
    static thread_local int configured_dev = -1;
    if (configured_dev != dev) {
        LTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        configured_dev = dev;
    }
    kern<<<grid, K8_THREADS, smem, st>>>(tmd, tmm, p);
    count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

static bool k8_shape_ok(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
                        int n_masks, int bpp) {
    // 16-byte aligned rows for the TMA; signals beyond 65536 pixels are K-split so that the
    // int32 accumulators stay exact
    const int align = 16 / bpp;
    return sig_size % align == 0 && ld_tile % align == 0 && (uintptr_t)tile % 16 == 0 &&
           sig_size >= 4 * (128 / bpp) && sig_size <= K8_MAX_SIG && n_frames >= 1 &&
           n_frames < (1ll << 31) && n_masks >= 1 && n_masks <= K8_MAX_COLUMNS;
}

}  // namespace ltb

using namespace ltb;

extern "C" size_t ltb200_masks_dense_i8_workspace(int tile_dtype, int64_t n_frames,
                                                  int64_t sig_size, int n_masks,
                                                  int with_sig_sum) {
    if (n_frames <= 0 || sig_size <= 0 || n_masks <= 0 || n_masks > K8_MAX_COLUMNS) return 0;
    const int bpp = tile_dtype == LTB_U8 ? 1 : 2;
    return k8_ws(n_frames, sig_size, n_masks, bpp, with_sig_sum != 0).total;
}

extern "C" int ltb200_masks_dense_i8(const void* tile, int tile_dtype, int64_t n_frames,
                                     int64_t sig_size, int64_t ld_tile, const int8_t* masks,
                                     int n_masks, int64_t ld_masks, float* out, int64_t ld_out,
                                     int accumulate, float* sig_sum, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_size >= 0 && n_masks >= 0, "masks_dense_i8: negative size");
    LTB_REQUIRE(tile_dtype == LTB_U16 || tile_dtype == LTB_U8,
                "masks_dense_i8: uint16 or uint8 tiles, got dtype %d", tile_dtype);
    if (n_frames == 0 || n_masks == 0) return LTB_OK;
    LTB_REQUIRE(tile != nullptr && masks != nullptr && out != nullptr,
                "masks_dense_i8: NULL pointer");
    LTB_REQUIRE(ld_tile >= sig_size && ld_masks >= sig_size && ld_out >= n_masks,
                "masks_dense_i8: leading dimension too small");
    const int bpp = tile_dtype == LTB_U8 ? 1 : 2;
    const int px = 128 / bpp;
    if (!k8_shape_ok(tile, n_frames, sig_size, ld_tile, n_masks, bpp)) {
        set_error("masks_dense_i8: shape not supported by the int8 tensor-core path (sig_size "
                  "%lld, ld_tile %lld, %d columns; need 16-byte aligned rows, %d <= sig_size <= "
                  "%lld, 1..%d columns)", (long long)sig_size, (long long)ld_tile, n_masks,
                  4 * px, (long long)K8_MAX_SIG, K8_MAX_COLUMNS);
        return LTB_ERR_UNSUPPORTED;
    }
    const K8Ws wl = k8_ws(n_frames, sig_size, n_masks, bpp, sig_sum != nullptr);
    if (wl.total > workspace_bytes || workspace == nullptr) {
        set_error("masks_dense_i8: workspace of %zu B required, %zu B given", wl.total,
                  workspace_bytes);
        return LTB_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    const int sms = sm_count();
    const int n = k8_n(n_masks, bpp);
    const int64_t sig_pad = ((sig_size + px - 1) / px) * px;
    const int64_t n_fb = (n_frames + K8_FB - 1) / K8_FB;

    K8Params p;
    p.n_frames = n_frames;
    p.sig_size = sig_size;
    p.n_masks = n_masks;
    p.ksplit = k8_choose_ksplit(n_fb, sig_size, sms, px);
    const int64_t subs = (sig_size + px - 1) / px;
    p.k_per_split = ((subs + p.ksplit - 1) / p.ksplit) * px;
    p.n_items = n_fb * p.ksplit;
    p.out = out;
    p.ld_out = ld_out;
    p.part = (long long*)(ws + wl.part_off);
    p.accumulate = accumulate;
    p.sig_acc = nullptr;
    p.debug = 0;
    if (const char* e = getenv("LTB200_K8_DEBUG")) p.debug = atoi(e);
    const int grid = (int)(p.n_items < sms ? p.n_items : sms);
    if (sig_sum != nullptr) {
        p.sig_acc = (unsigned long long*)(ws + wl.sig_off);
        LTB_CUDA_CHECK(cudaMemsetAsync(p.sig_acc, 0, (size_t)sig_size * 8, st));
    }
    {
        const int64_t total = (int64_t)(bpp == 2 ? n / 2 : n) * sig_pad;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sms * 8) blocks = sms * 8;
        if (bpp == 2)
            k8_pack_masks_kernel<<<blocks, 256, 0, st>>>(masks, n_masks, ld_masks, sig_size,
                                                         sig_pad, n / 2,
                                                         (uint16_t*)(ws + wl.pack_off));
        else
            k8_pack_masks_u8_kernel<<<blocks, 256, 0, st>>>(masks, n_masks, ld_masks, sig_size,
                                                            sig_pad, n,
                                                            (int8_t*)(ws + wl.pack_off));
        count_launch();
    }
    CUtensorMap tmd, tmm;
    int rc = encode_tmap_2d_sw(&tmd, tile,
                               bpp == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                        : CU_TENSOR_MAP_DATA_TYPE_UINT8,
                               (uint64_t)sig_size, (uint64_t)n_frames, (uint64_t)ld_tile * bpp,
                               (uint32_t)px, K8_FB, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != LTB_OK) return rc;
    rc = encode_tmap_2d_sw(&tmm, ws + wl.pack_off, CU_TENSOR_MAP_DATA_TYPE_UINT8,
                           (uint64_t)sig_pad * bpp, (uint64_t)n, (uint64_t)sig_pad * bpp, 128,
                           (uint32_t)n, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != LTB_OK) return rc;
    if (bpp == 1)
        rc = n == 16 ? k8_launch<16, 1>(tmd, tmm, p, grid, st)
                     : k8_launch<32, 1>(tmd, tmm, p, grid, st);
    else
        rc = n == 16 ? k8_launch<16, 2>(tmd, tmm, p, grid, st)
             : n == 32 ? k8_launch<32, 2>(tmd, tmm, p, grid, st)
                       : k8_launch<64, 2>(tmd, tmm, p, grid, st);
    if (rc != LTB_OK) return rc;
    if (p.ksplit > 1) {
        const int64_t total = n_frames * n_masks;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sms * 8) blocks = sms * 8;
        k8_finalize_kernel<<<blocks, 256, 0, st>>>(p.part, p.ksplit, n_frames, n_masks, out,
                                                   ld_out, accumulate);
        count_launch();
    }
    if (sig_sum != nullptr) {
        k8_sig_finalize_kernel<<<(unsigned)((sig_size + 255) / 256), 256, 0, st>>>(
            p.sig_acc, sig_size, sig_sum);
        count_launch();
    }
    LTB_CUDA_CHECK(cudaGetLastError());
    set_last_kernel(8);
    return LTB_OK;
}
