// k6_tensor.cu -- K6: dense masked reduction on the 5th-gen tensor cores (tcgen05 + TMEM)
//
//     out[f, m] (+)= sum_k tile[f, k] * masks[m, k]          (same seam as K1, float32 tiles)
//
// Replaces ApplyMasksEngine.process_flat (reference udf/masks.py:31-83, the torch.mm / numpy
// GEMM of a (frames x sig_size) tile with the (sig_size x masks) stack) for float32 tiles.
//
// Arithmetic: split-TF32 ("3xTF32").  Every float32 value is written as hi + lo with
//   hi = the top 19 bits (exactly what kind::tf32 reads of a float32 word -- the tensor core
//        ignores the low 13 bits, measured: bit-identical results with and without masking),
//   lo = x - hi (exact in float32), rounded to nearest at TF32 precision,
// for the frames (in registers, per element: LOP3 + FADD + IADD) and for the masks (once per
// call, k6_pack_masks_kernel), and the tensor core accumulates
//   hi_d*hi_m + hi_d*lo_m + lo_d*hi_m + lo_d*lo_m
// in float32.  All partial products are exact (11 x 11 significant bits); the dropped part of x
// is <= 2^-22 |x|.  Integer-valued data below 2^22 times binary masks is therefore bit-exact.
// The float32 accumulate inside the tensor core TRUNCATES (measured: bias -4.5e-8 x adds x
// |running sum|), so the TMEM accumulation chain is cut every `chain` sub-stages (32 pixels, 8
// MMA adds each) and the chain totals are added in float32 registers with round-to-nearest.
// chain = 1 (default) gives 3-5e-7 of the sum|x||m| scale against float64 -- the level of the
// reference's BLAS sgemm (2-6e-7) and of the FFMA2 kernel; chain = 8 is 2 % faster at 2e-6.
//
// Mapping (sm_100a, one persistent CTA per SM, 12 warps):
//   * work item = (block of 256 frames, K split).  The frames are the M dimension of the MMA
//     (two groups of 128 TMEM lanes), the mask columns the N dimension: N = 2 * NH where rows
//     [0, NH) of the packed mask tile hold hi(mask) and rows [NH, 2NH) hold lo(mask); the two
//     halves of an accumulator row are added in the drain.
//   * warp 0 (one lane): TMA producer of the frame stream: [256 frames x 32 px] boxes with the
//     128-byte swizzle into a 5-deep ring (evict_first).  128-byte rows stream at the same
//     6.4-6.5 TB/s as the 512-byte rows of K1 (measured with the compute switched off).
//   * warp 1 (one lane): TMA producer of the packed mask tile [N x 32 px] (K-major, 128-byte
//     swizzle = the canonical UMMA smem layout), L2-resident (evict_last), 4-deep ring.
//   * warps 4..11: converters.  Thread <-> frame row (= TMEM lane).  Each reads its 128-byte row
//     of the stage (conflict-free thanks to the swizzle), splits hi/lo in registers and writes
//     both as the A operand into TMEM with tcgen05.st (2-deep ring of 128 columns).
//   * warp 2: issues tcgen05.mma.kind::tf32 (M = 128, K = 8) with A from TMEM and B (masks)
//     from shared memory: per sub-stage and frame group 4 k-steps x (hi, lo); tcgen05.commit
//     releases the A slot and the mask slot.  The whole warp runs the issue loop and one
//     ELECTed lane issues: inside a divergent `if (lane == 0)` region ptxas wraps every UTCHMMA
//     in an election loop (~90 cycles per issue), which made the first version issue-bound at
//     0.80 of the HBM roofline; with uniform control flow the 16 MMAs of a sub-stage issue
//     back to back and the kernel is HBM-bound.  (A from shared memory -- the raw frames are
//     a valid hi operand -- was measured slower than A from TMEM and dropped.)
//   * the converters also drain the accumulators: tcgen05.ld of a finished chain is issued
//     before the conversion of the next sub-stage and consumed after it; at the end of an item
//     they store the (frames x columns) block.
//   * N = 64 (25-32 columns): eight more warps drain the accumulators and warp 3 issues the
//     MMAs of frame group 1 (template parameter DWM below); N = 32 / 64: the lo(x) MMAs read
//     only the hi(mask) rows (three products, K6Params::three).
// The frames never pass through the FP32 FMA pipe and the shared-memory operand traffic of the
// tensor core is only the (small) mask tile, so the kernel stays HBM-bound up to 24 columns
// (0.95-0.99 of the measured copy bandwidth in bursts) and reaches 0.90 at 32 columns, where the
// FFMA2 kernel is at 0.72 / 0.37.  Sustained (seconds of back-to-back launches) the board sits
// at its power limit and every form loses ~10 % to the SM clock.
// Switches (read per call): LTB200_K6_CHAIN, LTB200_K6_DW (0/1/2), LTB200_K6_ISSUERS (1/2),
// LTB200_K6_THREE (0/1), LTB200_K6_DEBUG (bring-up).
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace ltb {

constexpr int K6_FB = 256;            // frames per item
constexpr int K6_KS = 32;             // pixels per sub-stage (one 128-byte swizzle row)
constexpr int K6_DS = 5;              // data ring depth
constexpr int K6_MS = 4;              // mask ring depth
constexpr int K6_THREADS = 384;
constexpr int K6_CONV_WARPS = 8;
constexpr uint32_t K6_DATA_BYTES = K6_FB * K6_KS * 4;   // 32 KiB per sub-stage
constexpr int K6_TMEM_COLS = 512;
constexpr int K6_AS = 2;              // TMEM ring of A operands
constexpr int K6_A_BASE = 256;        // TMEM columns [256, 512): 2 slots x 2 groups x (hi 32 | lo 32)
constexpr int K6_U16_MAX_COLUMNS = 16; // uint16 form: N in {16, 32}

struct K6Params {
    int64_t n_frames;
    int64_t sig_size;
    int n_masks;
    int ksplit;
    int64_t k_per_split;   // multiple of K6_KS
    int64_t n_items;
    float* out;
    int64_t ld_out;
    float* part;           // (ksplit, n_frames, n_masks) when ksplit > 1
    int accumulate;
    int chain;             // sub-stages per TMEM accumulation chain
    int debug;             // bring-up switches (LTB200_K6_DEBUG): 1 no convert, 2 no MMA, 4 no drain
    unsigned long long* sig_acc;   // uint16 tiles: (sig_size) exact integer frame sums, or NULL
    uint32_t zero;                 // 0, unknown to the compiler (stage release of the sum warps)
    int issuers;                   // 1: warp 2 issues every MMA; 2: warp 3 takes frame group 1
    int three;                     // 1: lo(x) meets hi(mask) only (N / 2 columns): "3xTF32"
};

// per input type: a data stage is one 128-byte row per frame = 32 float32 or 64 uint16 pixels,
// i.e. HALVES sub-stages of 32 pixels (the unit of the mask ring, the TMEM operand ring and the
// accumulation chains).  uint16 tiles get four extra warps for the fused frame sum (SumUDF).
template <typename TIN> struct K6In;
template <> struct K6In<float> {
    static constexpr int HALVES = 1;
    static constexpr int THREADS = K6_THREADS;
    static constexpr int SUM_WARPS = 0;
};
template <> struct K6In<uint16_t> {
    static constexpr int HALVES = 2;
    static constexpr int THREADS = K6_THREADS + 128;
    static constexpr int SUM_WARPS = 4;
};

// ---- tcgen05 wrappers ----------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void tc_mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, "
        "%11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
        "r"(v[15])
        : "memory");
}
// TMEM -> registers, 16 columns of this thread's lane; completion by tc_ld_fence16 below
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, "
        "%11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// pins the registers of an earlier tc_ld16 behind the tc_wait_ld that precedes this call
// (volatile asm statements keep their order; the in/out operands tie the uses to it)
__device__ __forceinline__ void tc_ld_fence16(uint32_t (&r)[16]) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]),
                   "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]),
                   "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tc_ldn(uint32_t taddr, uint32_t (&r)[16]) { tc_ld16(taddr, r); }
__device__ __forceinline__ void tc_ld_fencen(uint32_t (&r)[16]) { tc_ld_fence16(r); }
__device__ __forceinline__ void tc_ldn(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld_fencen(uint32_t (&r)[8]) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]),
                   "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
// acc (two columns) += x * hi(mask) + x * lo(mask) of a drained chain: columns c, c + 1 and
// nh + c, nh + c + 1 of the accumulator row, two packed adds (same rounding as the scalar form)
template <int Q>
__device__ __forceinline__ void k6_add_pair(float& a0, float& a1, const uint32_t (&v)[Q][16],
                                            int c, int nh) {
    const float2 t2 = __fadd2_rn(
        make_float2(__uint_as_float(v[c / 16][c % 16]), __uint_as_float(v[c / 16][c % 16 + 1])),
        make_float2(__uint_as_float(v[(nh + c) / 16][(nh + c) % 16]),
                    __uint_as_float(v[(nh + c) / 16][(nh + c) % 16 + 1])));
    const float2 a2 = __fadd2_rn(make_float2(a0, a1), t2);
    a0 = a2.x;
    a1 = a2.y;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
// (bits: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=2 [61,64))
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// packed[(r), k]: r in [0, NH) -> hi(mask r), r in [NH, 2NH) -> lo(mask r - NH); zero padded
__global__ void k6_pack_masks_kernel(const float* __restrict__ masks, int n_masks,
                                     int64_t ld_masks, int64_t sig_size, int64_t sig_pad, int nh,
                                     float* __restrict__ packed) {
    const int64_t total = (int64_t)nh * sig_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / sig_pad);
        const int64_t k = i % sig_pad;
        float m = 0.f;
        if (r < n_masks && k < sig_size) m = masks[(int64_t)r * ld_masks + k];
        const float hi = __uint_as_float(tf32_rna(m));
        const float lo = __uint_as_float(tf32_rna(m - hi));
        packed[(int64_t)r * sig_pad + k] = hi;
        packed[(int64_t)(r + nh) * sig_pad + k] = lo;
    }
}

__global__ void k6_finalize_kernel(const float* __restrict__ part, int ksplit, int64_t n_frames,
                                   int n_masks, float* __restrict__ out, int64_t ld_out,
                                   int accumulate) {
    const int64_t total = n_frames * n_masks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < ksplit; k++) s += part[(int64_t)k * total + i];
        float* o = out + (i / n_masks) * ld_out + (i % n_masks);
        *o = accumulate ? (*o + s) : s;
    }
}

struct K6Smem {
    // offsets from the 1024-byte aligned base
    static constexpr uint32_t data_off(int s) { return (uint32_t)s * K6_DATA_BYTES; }
    static constexpr uint32_t mask_off(int s, int n) {
        return K6_DS * K6_DATA_BYTES + (uint32_t)s * (uint32_t)n * 128u;
    }
    static constexpr uint32_t bar_off(int n) { return mask_off(K6_MS, n); }
    static constexpr uint32_t total(int n) { return bar_off(n) + 256 + 1024; }   // + align slack
};

// DW ("drain warps", float32 tiles with N = 64 only): the accumulators are drained by eight
// extra warps (12..19) instead of the converters.  At 25-32 columns the converters' instruction
// stream per sub-stage (conversion + 64-column drain, ~330 instructions at ~5 cycles each in a
// warp that owns its frame rows alone) was longer than the HBM time of the sub-stage; the drain
// is a third of it and runs concurrently here.  Handshake: the MMA warp commits `acc_full[b]` at
// the end of a chain and waits for `acc_free[b]` (8 drain warps) before it restarts buffer b;
// chains are numbered over the lifetime of the CTA, so every parity is (chain / 2) & 1.
// DWM = 2 additionally doubles the converter warps (4..19, drain warps 20..27): the two warps
// of a frame-row quarter split the 32 pixels of the sub-stage, which halves the instruction
// stream in front of every A-operand hand-over.  Validated (tests/test_k6_gpu.py) but measured
// SLOWER than DWM = 1 (0.73-0.75 vs 0.76-0.79 sustained at 25-32 columns): at the board's power
// limit more resident warps cost SM clock; kept as LTB200_K6_DW=2 for comparison.
constexpr int K6_DRAIN_WARPS = 8;
__host__ __device__ constexpr int k6_threads(int base, int dwm) {
    return base + (dwm > 0 ? K6_DRAIN_WARPS * 32 : 0) + (dwm == 2 ? K6_CONV_WARPS * 32 : 0);
}

template <int N, typename TIN, int DWM = 0>
__global__ void __launch_bounds__(k6_threads(K6In<TIN>::THREADS, DWM), 1)
k6_tensor_kernel(const __grid_constant__ CUtensorMap tm_data,
                 const __grid_constant__ CUtensorMap tm_mask, const K6Params p) {
    constexpr int NH = N / 2;
    constexpr int HALVES = K6In<TIN>::HALVES;
    constexpr int PX = K6_KS * HALVES;            // pixels per data stage
    constexpr uint32_t MASK_BYTES = (uint32_t)N * 128u;
    constexpr uint32_t IDESC = umma_idesc_tf32(N);
    static_assert(N % 16 == 0 && N >= 16 && N <= 64, "K6: N in {16, 32, 48, 64}");
    constexpr bool DW = DWM > 0;
    constexpr int CONVW = DWM == 2 ? 2 * K6_CONV_WARPS : K6_CONV_WARPS;
    // registers per role after setmaxnreg: 640 x 96 -> 40 / 88 / 128; 896 x 72 -> 40 / 64 / 104
    constexpr int DRAIN_COLS = DWM == 2 ? 8 : 16;     // columns per tcgen05.ld round of the drain
    static_assert(!DW || (N == 64 && K6In<TIN>::HALVES == 1), "K6 drain warps: float32, N = 64");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K6Smem::bar_off(N));
    uint64_t* data_full = bars;                       // [DS]  TMA landed
    uint64_t* data_free = data_full + K6_DS;          // [DS]  8 converter warps + the hi MMAs
    uint64_t* mask_full = data_free + K6_DS;          // [MS]
    uint64_t* mask_empty = mask_full + K6_MS;         // [MS]
    uint64_t* a_full = mask_empty + K6_MS;            // [AS]  lo parts written to TMEM
    uint64_t* mma_done = a_full + K6_AS;              // [AS]  MMAs of the sub-stage completed
    uint64_t* acc_full = mma_done + K6_AS;            // [2]   DW: chain in buffer b completed
    uint64_t* acc_free = acc_full + 2;                // [2]   DW: buffer b drained (8 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

    // warp index made warp-uniform for the compiler: the MMA issue loop must be uniform control
    // flow, otherwise every UTCHMMA is wrapped in an election loop (~90 cycles per issue)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K6_DS; s++) {
            mbar_init(&data_full[s], 1);
            mbar_init(&data_free[s],
                      CONVW + (p.sig_acc != nullptr ? K6In<TIN>::SUM_WARPS : 0));
        }
        for (int s = 0; s < K6_MS; s++) {
            mbar_init(&mask_full[s], 1);
            mbar_init(&mask_empty[s], p.issuers);
        }
        for (int s = 0; s < K6_AS; s++) {
            mbar_init(&a_full[s], CONVW);
            mbar_init(&mma_done[s], p.issuers);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(&acc_full[s], p.issuers);
            mbar_init(&acc_free[s], K6_DRAIN_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "n"(K6_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int chain = p.chain;

    // DW: 640 threads x 96 registers are re-dealt per role (setmaxnreg, whole warpgroups):
    // producers / issuer 40, converters 88, drain warps 128
    if (warp == 0) {
        // ===== frame stream producer =====
        if constexpr (DW) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (lane == 0) {
            prefetch_tmap(&tm_data);
            const uint64_t pol = l2_policy_evict_first();
            uint32_t it = 0;
            for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int64_t fb = item / p.ksplit;
                const int64_t k0 = (item % p.ksplit) * p.k_per_split;
                int64_t k1 = k0 + p.k_per_split;
                if (k1 > p.sig_size) k1 = p.sig_size;
                const int n_dsub = (int)((k1 - k0 + PX - 1) / PX);
                const int32_t f0 = (int32_t)(fb * K6_FB);
                for (int i = 0; i < n_dsub; i++, it++) {
                    const int ds = it % K6_DS;
                    mbar_wait(&data_free[ds], ((it / K6_DS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&data_full[ds], K6_DATA_BYTES);
                    tma_load_2d(smem + K6Smem::data_off(ds), &tm_data, (int32_t)(k0 + i * PX),
                                f0, &data_full[ds], pol);
                }
            }
        }
    } else if (warp == 1) {
        // ===== mask tile producer =====
        if constexpr (DW) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (lane == 0) {
            prefetch_tmap(&tm_mask);
            const uint64_t pol = l2_policy_evict_last();
            uint32_t it = 0;
            for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int64_t k0 = (item % p.ksplit) * p.k_per_split;
                int64_t k1 = k0 + p.k_per_split;
                if (k1 > p.sig_size) k1 = p.sig_size;
                const int n_sub = (int)((k1 - k0 + PX - 1) / PX) * HALVES;
                for (int i = 0; i < n_sub; i++, it++) {
                    const int ms = it % K6_MS;
                    mbar_wait(&mask_empty[ms], ((it / K6_MS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&mask_full[ms], MASK_BYTES);
                    tma_load_2d(smem + K6Smem::mask_off(ms, N), &tm_mask,
                                (int32_t)(k0 + i * K6_KS), 0, &mask_full[ms], pol);
                }
            }
        }
    } else if (warp == 2 || (warp == 3 && p.issuers == 2)) {
        // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
        // One thread issues a tcgen05.mma at most every ~45 cycles (scripts/ubench/
        // mma_rate_probe.cu), i.e. 16 MMAs = 710 cycles of the ~1400 a sub-stage has at the HBM
        // rate; with `issuers == 2` warp 3 issues the MMAs of frame group 1 and commits onto the
        // same mbarriers (count 2), which halves the issue phase of the hand-over loop.
        if constexpr (DW) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        // `three`: the lo(x) MMAs read only the hi(mask) rows [0, NH) of the tile (N = NH)
        const uint32_t idesc_lo =
            (NH % 16 == 0 && p.three) ? umma_idesc_tf32(NH % 16 == 0 ? NH : N) : IDESC;
        const int g_lo = (p.issuers == 2 && warp == 3) ? 1 : 0;
        const int g_hi = (p.issuers == 2 && warp == 2) ? 1 : 2;
        uint32_t it = 0;
        uint32_t gc = 0;                             // DW: chains started by this CTA
        for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int64_t k0 = (item % p.ksplit) * p.k_per_split;
            int64_t k1 = k0 + p.k_per_split;
            if (k1 > p.sig_size) k1 = p.sig_size;
            const int n_sub = (int)((k1 - k0 + PX - 1) / PX) * HALVES;
            int in_chain = 0, cbuf = 0;
            for (int i = 0; i < n_sub; i++, it++) {
                const int ms = it % K6_MS;
                const int as = it % K6_AS;
                mbar_wait(&mask_full[ms], (it / K6_MS) & 1);
                mbar_wait(&a_full[as], (it / K6_AS) & 1);
                bool chain_last = false;
                if constexpr (DW) {
                    if (in_chain == 0) {
                        // buffer gc & 1 restarts: its previous chain (gc - 2) must be drained
                        cbuf = (int)(gc & 1u);
                        mbar_wait(&acc_free[cbuf], ((gc >> 1) & 1u) ^ 1u);
                    }
                    chain_last = in_chain + 1 == chain || i == n_sub - 1;
                }
                tc_fence_after();
                const uint64_t bdesc0 =
                    umma_desc_k_sw128(smem_u32(smem + K6Smem::mask_off(ms, N)));
                const uint32_t a0 = tmem_base + (uint32_t)(K6_A_BASE + as * 128);
                const uint32_t d0 = tmem_base + (uint32_t)(cbuf * N);
                if (elect_one()) {
                    if (!(p.debug & 2)) {
#pragma unroll
                        for (int g = 0; g < 2; g++) {
                            if (g < g_lo || g >= g_hi) continue;
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) {
                                const uint64_t bdesc = bdesc0 + (uint64_t)(kk * 2);   // +32 bytes
                                tc_mma_tf32_ts(d0 + g * 2 * N, a0 + g * 64 + kk * 8, bdesc, IDESC,
                                               (in_chain | kk) != 0 ? 1u : 0u);
                                tc_mma_tf32_ts(d0 + g * 2 * N, a0 + g * 64 + 32 + kk * 8, bdesc,
                                               idesc_lo, 1u);
                            }
                        }
                    }
                    tc_commit(&mma_done[as]);
                    tc_commit(&mask_empty[ms]);
                    if (DW && chain_last) tc_commit(&acc_full[cbuf]);
                }
                __syncwarp();
                if constexpr (DW) {
                    in_chain++;
                    if (chain_last) {
                        in_chain = 0;
                        gc++;
                    }
                } else if (++in_chain == chain) {
                    in_chain = 0;
                    cbuf ^= 1;
                }
            }
        }
    } else if (DW && warp == 3) {
        // (setmaxnreg is warpgroup-wide: the TMEM-allocating warp of warpgroup 0 joins in)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    } else if (warp >= 4 && warp < 4 + CONVW) {
        // ===== converters / accumulator drain =====
        if constexpr (DWM == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        if constexpr (DWM == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const int cw = warp - 4;
        const int px_half = cw >> 3;                // DWM == 2: which 16 pixels of the sub-stage
        const int g = (cw >> 2) & 1;
        const int w = cw & 3;                       // == warp % 4: the TMEM lane quarter
        const int row = g * 128 + w * 32 + lane;    // frame row inside the item
        const uint32_t lane_sel = (uint32_t)(w * 32) << 16;
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t row_off = (uint32_t)row * 128u;

        uint32_t it = 0;                             // 32-pixel sub-stages
        uint32_t dit = 0;                            // data stages
        for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int64_t fb = item / p.ksplit;
            const int ksi = (int)(item % p.ksplit);
            const int64_t k0 = (int64_t)ksi * p.k_per_split;
            int64_t k1 = k0 + p.k_per_split;
            if (k1 > p.sig_size) k1 = p.sig_size;
            const int n_dsub = (int)((k1 - k0 + PX - 1) / PX);
            const int n_sub = n_dsub * HALVES;

            // Chain totals are summed in float32 registers (round to nearest).  Adding thousands
            // of them into ONE running sum is what dominated this kernel's rounding error
            // (1.4e-6 of sum|x||m| for a 256x256 signal, measured and emulated); with <= 16
            // columns there are registers for a second level: `acc` collects 32 chains, then
            // moves into `acc_hi` (3-4e-7, the level of the reference's blocked sgemm).
            constexpr bool TWO_LEVEL = (NH <= 16 && HALVES == 1) || NH <= 8;   // (uint16 form: 128 regs)
            float acc[NH];
            float acc_hi[TWO_LEVEL ? NH : 1];
#pragma unroll
            for (int c = 0; c < NH; c++) acc[c] = 0.f;
#pragma unroll
            for (int c = 0; c < (TWO_LEVEL ? NH : 1); c++) acc_hi[c] = 0.f;
            int in_block = 0;                        // chains collected in `acc` (TWO_LEVEL)
            auto level_up = [&]() {
                if constexpr (TWO_LEVEL) {
                    if (++in_block == 32) {
                        in_block = 0;
#pragma unroll
                        for (int c = 0; c < NH; c++) {
                            acc_hi[c] += acc[c];
                            acc[c] = 0.f;
                        }
                    }
                }
            };
            int next_chain = 0;                      // first chain not yet drained
            int chain_pos = 0;                       // i % chain, kept as a counter (no division)
            int nc_start = 0;                        // next_chain * chain, likewise
            auto next_end = [&]() {                  // last sub-stage of chain `next_chain`
                const int e = nc_start + chain;
                return (e < n_sub ? e : n_sub) - 1;
            };
            // drain every chain whose last sub-stage is <= done (their MMAs have completed)
            auto drain_upto = [&](int done) {
                while (nc_start < n_sub && next_end() <= done) {
                    const uint32_t d =
                        tmem_base + lane_sel + (uint32_t)((g * 2 + (next_chain & 1)) * N);
                    uint32_t v[N / 16][16];
#pragma unroll
                    for (int q = 0; q < N / 16; q++) tc_ld16(d + q * 16, v[q]);
                    tc_wait_ld();
#pragma unroll
                    for (int q = 0; q < N / 16; q++) tc_ld_fence16(v[q]);
#pragma unroll
                    for (int c = 0; c < NH; c += 2) k6_add_pair(acc[c], acc[c + 1], v, c, NH);
                    level_up();
                    next_chain++;
                    nc_start += chain;
                }
            };

            for (int di = 0; di < n_dsub; di++, dit++) {
                const int ds = dit % K6_DS;
                mbar_wait(&data_full[ds], (dit / K6_DS) & 1);
                const uint8_t* rp = smem + K6Smem::data_off(ds) + row_off;
                uint4 x[8];                           // this frame's 128-byte row of the stage
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (DWM == 2 && (j >> 2) != px_half) continue;     // (its 64-byte half)
                    x[j] = *reinterpret_cast<const uint4*>(rp + (((uint32_t)j ^ swz) << 4));
                }
#pragma unroll
                for (int h2 = 0; h2 < HALVES; h2++, it++) {
                    const int i = di * HALVES + h2;
                    const int as = it % K6_AS;
                    // lo slot `as` is free once the MMAs of sub-stage i - AS have completed
                    mbar_wait(&mma_done[as], ((it / K6_AS) & 1) ^ 1);
                    int known = i - K6_AS;
                    // chain i/chain - 2 shares its accumulator with the chain that starts at
                    // sub-stage i: it must be drained before this sub-stage is handed to the MMAs
                    // (DW: the MMA warp waits for the drain warps instead)
                    if (!DW && chain_pos == 0 && i >= 2 * chain) {
                        const int must = i - chain - 1;
                        if (must > known) {
                            const uint32_t itm = it - (uint32_t)(i - must);
                            mbar_wait(&mma_done[itm % K6_AS], (itm / K6_AS) & 1);
                            known = must;
                        }
                    }
                    tc_fence_after();
                    // at most one chain completes per sub-stage: issue its TMEM loads now, use
                    // them after the conversion below (the load latency hides behind the ALU work)
                    uint32_t v[N / 16][16];
                    bool pend = false;
                    if (!DW && known >= 0 && !(p.debug & 4) && nc_start < n_sub &&
                        next_end() <= known) {
                        const uint32_t d =
                            tmem_base + lane_sel + (uint32_t)((g * 2 + (next_chain & 1)) * N);
#pragma unroll
                        for (int q = 0; q < N / 16; q++) tc_ld16(d + q * 16, v[q]);
                        pend = true;
                        next_chain++;
                        nc_start += chain;
                    }
                    const uint32_t a =
                        tmem_base + lane_sel + (uint32_t)(K6_A_BASE + as * 128 + g * 64);
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (p.debug & 1) break;
                        if (DWM == 2 && h != px_half) continue;
                        uint32_t hi[16], lo[16];
                        if constexpr (HALVES == 1) {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const uint4 xv = x[h * 4 + j];
                                const uint32_t e[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                                for (int t = 0; t < 4; t += 2) {
                                    // hi = top 19 bits (a TF32 number); lo = x - hi exactly (two
                                    // pixels per FADD2), rounded to nearest at TF32 precision by
                                    // adding half a TF32 ulp (kind::tf32 ignores the low 13 bits)
                                    const uint32_t h0 = e[t] & 0xFFFFE000u;
                                    const uint32_t h1 = e[t + 1] & 0xFFFFE000u;
                                    const float2 d = __fadd2_rn(
                                        make_float2(__uint_as_float(e[t]), __uint_as_float(e[t + 1])),
                                        make_float2(-__uint_as_float(h0), -__uint_as_float(h1)));
                                    hi[j * 4 + t] = h0;
                                    hi[j * 4 + t + 1] = h1;
                                    lo[j * 4 + t] = __float_as_uint(d.x) + 0x1000u;
                                    lo[j * 4 + t + 1] = __float_as_uint(d.y) + 0x1000u;
                                }
                            }
                        } else {
                            // uint16 pixels: 0x4B00vvvv is the float 2^23 + v, so v as a float
                            // costs PRMT + FADD (exact for all 16-bit values); hi = its top 11
                            // significant bits, lo = v - hi has <= 5 bits: both exact TF32 numbers
#pragma unroll
                            for (int j = 0; j < 2; j++) {
                                const uint4 xv = x[h2 * 4 + h * 2 + j];
                                const uint32_t e[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                                for (int t = 0; t < 4; t++) {
#pragma unroll
                                    for (int u = 0; u < 2; u++) {
                                        const float f =
                                            __uint_as_float(__byte_perm(e[t], 0x4B000000u,
                                                                        u ? 0x7632 : 0x7610)) -
                                            8388608.f;
                                        const uint32_t hb = __float_as_uint(f) & 0xFFFFE000u;
                                        hi[j * 8 + t * 2 + u] = hb;
                                        lo[j * 8 + t * 2 + u] =
                                            __float_as_uint(f - __uint_as_float(hb));
                                    }
                                }
                            }
                        }
                        tc_st16(a + h * 16, hi);
                        tc_st16(a + 32 + h * 16, lo);
                    }
                    if (pend) {
                        tc_wait_ld();
#pragma unroll
                        for (int q = 0; q < N / 16; q++) tc_ld_fence16(v[q]);
#pragma unroll
                        for (int c = 0; c < NH; c += 2) k6_add_pair(acc[c], acc[c + 1], v, c, NH);
                        level_up();
                    }
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&a_full[as]);
                        if (h2 == HALVES - 1) mbar_arrive(&data_free[ds]);
                    }
                    if (++chain_pos == chain) chain_pos = 0;
                }
            }
            if constexpr (DW) continue;               // drained and stored by warps 12..19
            // item tail: the last commit covers every earlier MMA of the item
            {
                const uint32_t itl = it - 1;
                mbar_wait(&mma_done[itl % K6_AS], (itl / K6_AS) & 1);
            }
            tc_fence_after();
            if (!(p.debug & 4)) drain_upto(n_sub - 1);
            tc_fence_before();
            if constexpr (TWO_LEVEL) {
#pragma unroll
                for (int c = 0; c < NH; c++) acc[c] += acc_hi[c];
            }

            const int64_t f = fb * K6_FB + row;
            if (f < p.n_frames) {
                if (p.ksplit == 1) {
                    float* o = p.out + f * p.ld_out;
#pragma unroll
                    for (int c = 0; c < NH; c++)
                        if (c < p.n_masks) o[c] = p.accumulate ? (o[c] + acc[c]) : acc[c];
                } else {
                    float* o = p.part + ((int64_t)ksi * p.n_frames + f) * p.n_masks;
#pragma unroll
                    for (int c = 0; c < NH; c++)
                        if (c < p.n_masks) o[c] = acc[c];
                }
            }
        }
    } else if (DW && warp >= 4 + CONVW) {
        // ===== DW: accumulator drain (thread <-> frame row = TMEM lane, like the converters) =====
        if constexpr (DW) {
            if constexpr (DWM == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
            if constexpr (DWM == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
            const int cw = warp - (4 + CONVW);
            const int g = cw >> 2;
            const int w = cw & 3;                   // == warp % 4: the TMEM lane quarter
            const int row = g * 128 + w * 32 + lane;
            const uint32_t lane_sel = (uint32_t)(w * 32) << 16;
            uint32_t gc = 0;
            for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int64_t fb = item / p.ksplit;
                const int ksi = (int)(item % p.ksplit);
                const int64_t k0 = (int64_t)ksi * p.k_per_split;
                int64_t k1 = k0 + p.k_per_split;
                if (k1 > p.sig_size) k1 = p.sig_size;
                const int n_sub = (int)((k1 - k0 + PX - 1) / PX);
                const int n_chains = (n_sub + chain - 1) / chain;
                // two-level sum of the chain totals (see the converters): 32 chains, then up
                float acc[NH], acc_hi[NH];
#pragma unroll
                for (int c = 0; c < NH; c++) acc[c] = acc_hi[c] = 0.f;
                int in_block = 0;
                for (int c = 0; c < n_chains; c++, gc++) {
                    const uint32_t b = gc & 1u;
                    mbar_wait(&acc_full[b], (gc >> 1) & 1u);
                    tc_fence_after();
                    const uint32_t d = tmem_base + lane_sel + (uint32_t)((g * 2 + (int)b) * N);
                    // columns [0, 32): x * hi(mask), [32, 64): x * lo(mask); 16 columns of each
                    // per round keep the live registers at acc + acc_hi + 32
#pragma unroll
                    for (int h = 0; h < NH / DRAIN_COLS; h++) {
                        uint32_t vh[DRAIN_COLS], vl[DRAIN_COLS];
                        tc_ldn(d + h * DRAIN_COLS, vh);
                        tc_ldn(d + NH + h * DRAIN_COLS, vl);
                        tc_wait_ld();
                        tc_ld_fencen(vh);
                        tc_ld_fencen(vl);
#pragma unroll
                        for (int j = 0; j < DRAIN_COLS; j += 2) {
                            const float2 t2 = __fadd2_rn(
                                make_float2(__uint_as_float(vh[j]), __uint_as_float(vh[j + 1])),
                                make_float2(__uint_as_float(vl[j]), __uint_as_float(vl[j + 1])));
                            const float2 a2 = __fadd2_rn(
                                make_float2(acc[h * DRAIN_COLS + j], acc[h * DRAIN_COLS + j + 1]), t2);
                            acc[h * DRAIN_COLS + j] = a2.x;
                            acc[h * DRAIN_COLS + j + 1] = a2.y;
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_free[b]);
                    if (++in_block == 32) {
                        in_block = 0;
#pragma unroll
                        for (int cc = 0; cc < NH; cc++) {
                            acc_hi[cc] += acc[cc];
                            acc[cc] = 0.f;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < NH; c++) acc[c] += acc_hi[c];
                const int64_t f = fb * K6_FB + row;
                if (f < p.n_frames) {
                    if (p.ksplit == 1) {
                        float* o = p.out + f * p.ld_out;
#pragma unroll
                        for (int c = 0; c < NH; c++)
                            if (c < p.n_masks) o[c] = p.accumulate ? (o[c] + acc[c]) : acc[c];
                    } else {
                        float* o = p.part + ((int64_t)ksi * p.n_frames + f) * p.n_masks;
#pragma unroll
                        for (int c = 0; c < NH; c++)
                            if (c < p.n_masks) o[c] = acc[c];
                    }
                }
            }
        }
    } else if (K6In<TIN>::SUM_WARPS > 0 && warp >= 4 + K6_CONV_WARPS) {
        // ===== fused frame sum (SumUDF, reference udf/sum.py:44-49) of uint16 tiles =====
        // Warp sw owns the logical 16-byte chunks {2 sw, 2 sw + 1} (16 pixels) of every row of
        // the stage.  Per step its four lane octets read rows r, r + 2, r + 4, r + 6: the
        // 128-byte swizzle (chunk ^ (row & 7)) sends the same chunk pair of those rows to four
        // different bank groups, so the LDS.32 is conflict-free.  Sums are exact integers:
        // 64 rows x 65535 < 2^32 per lane and stage, then one 64-bit RED per pixel and stage.
        if (p.sig_acc != nullptr) {
            const int sw = warp - (4 + K6_CONV_WARPS);
            const int q = lane >> 3;                 // row octet of the step
            const int sub = lane & 7;
            const uint32_t chunk = (uint32_t)(2 * sw + (sub >> 2));
            const uint32_t word = (uint32_t)(sub & 3) * 4u;
            uint32_t dit = 0;
            for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int64_t k0 = (item % p.ksplit) * p.k_per_split;
                int64_t k1 = k0 + p.k_per_split;
                if (k1 > p.sig_size) k1 = p.sig_size;
                const int n_dsub = (int)((k1 - k0 + PX - 1) / PX);
                for (int di = 0; di < n_dsub; di++, dit++) {
                    const int ds = dit % K6_DS;
                    mbar_wait(&data_full[ds], (dit / K6_DS) & 1);
                    const uint8_t* base = smem + K6Smem::data_off(ds);
                    uint32_t s0 = 0, s1 = 0;
#pragma unroll 8
                    for (int st = 0; st < K6_FB / 4; st++) {
                        const uint32_t r = (uint32_t)((st >> 1) * 8 + (st & 1) + 2 * q);
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(
                            base + r * 128u + ((chunk ^ (r & 7u)) << 4) + word);
                        s0 += v & 0xFFFFu;
                        s1 += v >> 16;
                    }
                    __syncwarp();
                    // the release must not overtake the loads of this stage: its address depends
                    // on the sums (p.zero is 0 at run time only), see k7_group_tensor.cu
                    if (lane == 0) mbar_arrive(&data_free[ds] + ((s0 ^ s1) & p.zero));
                    s0 += __shfl_xor_sync(0xffffffffu, s0, 8);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
                    s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                    if (q == 0) {
                        const int64_t px = k0 + (int64_t)di * PX + chunk * 8 + (sub & 3) * 2;
                        if (px < p.sig_size) atomicAdd(p.sig_acc + px, (unsigned long long)s0);
                        if (px + 1 < p.sig_size)
                            atomicAdd(p.sig_acc + px + 1, (unsigned long long)s1);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(K6_TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int k6_choose_ksplit(int64_t n_fb, int64_t sig_size, int sms, int px) {
    int best = 1;
    double best_eff = 0.0;
    for (int ks = 1; ks <= 64; ks *= 2) {
        if (ks > 1 && sig_size / ks < 16 * px) break;
        const int64_t items = n_fb * ks;
        const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = ks;
        }
        if (eff >= 0.95) break;
    }
    return best;
}

static size_t k6_align256(size_t v) { return (v + 255) & ~(size_t)255; }

static int k6_nh(int n_masks) { return n_masks <= 8 ? 8 : n_masks <= 16 ? 16 : n_masks <= 24 ? 24 : 32; }

struct K6Ws {
    size_t pack_off, part_off, sig_off, total;
};

// px = pixels per data stage (32 for float32, 64 for uint16 tiles)
static K6Ws k6_ws(int64_t n_frames, int64_t sig_size, int n_masks, int px, bool with_sig) {
    K6Ws w;
    const int nm = n_masks > 32 ? 32 : n_masks;
    const int64_t sig_pad = ((sig_size + 63) / 64) * 64;
    w.pack_off = 0;
    const size_t pack = (size_t)2 * k6_nh(nm) * sig_pad * sizeof(float);
    w.part_off = k6_align256(pack);
    const int64_t n_fb = (n_frames + K6_FB - 1) / K6_FB;
    const int ks = k6_choose_ksplit(n_fb, sig_size, sm_count(), px);
    const size_t part = ks > 1 ? (size_t)ks * n_frames * nm * sizeof(float) : 0;
    w.sig_off = w.part_off + k6_align256(part);
    w.total = w.sig_off + (with_sig ? k6_align256((size_t)sig_size * 8) : 0);
    return w;
}

template <int N, typename TIN, int DWM = 0>
static int k6_launch(const CUtensorMap& tmd, const CUtensorMap& tmm, const K6Params& p, int grid,
                     cudaStream_t st) {
    auto kern = k6_tensor_kernel<N, TIN, DWM>;
    const size_t smem = K6Smem::total(N);
    int dev = 0;
    LTB_CUDA_CHECK(cudaGetDevice(&dev));
    static thread_local int configured_dev = -1;
    if (configured_dev != dev) {
        if constexpr (DWM > 0) {
            // setmaxnreg re-deals threads x compiled registers (40 / 88 / 128 or 40 / 64 / 104
            // per role): with fewer compiled registers setmaxnreg.inc would block for ever
            cudaFuncAttributes fa;
            LTB_CUDA_CHECK(cudaFuncGetAttributes(&fa, kern));
            const int need = DWM == 1 ? 128 * 40 + 256 * 88 + 256 * 128
                                      : 128 * 40 + 512 * 64 + 256 * 104;
            if (fa.numRegs * k6_threads(K6In<TIN>::THREADS, DWM) < need) {
                set_error("masks_dense_tc: drain-warp form compiled with %d registers per "
                          "thread; use LTB200_K6_DW=0", fa.numRegs);
                return LTB_ERR_UNSUPPORTED;
            }
        }
        LTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        configured_dev = dev;
    }
    kern<<<grid, k6_threads(K6In<TIN>::THREADS, DWM), smem, st>>>(tmd, tmm, p);
    count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

bool k6_shape_ok(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile) {
    return sig_size % 4 == 0 && ld_tile % 4 == 0 && (uintptr_t)tile % 16 == 0 &&
           sig_size >= 4 * K6_KS && n_frames >= 1 && sig_size < (1ll << 30) &&
           n_frames < (1ll << 31);
}

// uint16 tiles: 16-byte aligned rows; <= 16 columns (the 512-thread form has 128 registers)
bool k6_u16_shape_ok(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
                     int n_masks) {
    return sig_size % 8 == 0 && ld_tile % 8 == 0 && (uintptr_t)tile % 16 == 0 &&
           sig_size >= 8 * K6_KS && n_frames >= 1 && sig_size < (1ll << 30) &&
           n_frames < (1ll << 31) && n_masks >= 1 && n_masks <= K6_U16_MAX_COLUMNS;
}

size_t k6_workspace(int64_t n_frames, int64_t sig_size, int n_masks) {
    return k6_ws(n_frames, sig_size, n_masks, K6_KS, false).total;
}

size_t k6_u16_workspace(int64_t n_frames, int64_t sig_size, int n_masks, int with_sig_sum) {
    return k6_ws(n_frames, sig_size, n_masks, 2 * K6_KS, with_sig_sum != 0).total;
}

int k6_default_chain() {
    static int chain = -1;
    if (chain < 0) {
        chain = 1;
        if (const char* e = getenv("LTB200_K6_CHAIN")) {
            const int v = atoi(e);
            if (v >= 1 && v <= (1 << 20)) chain = v;
        }
    }
    return chain;
}

// 25-32 columns: drain warps on (LTB200_K6_DW=0 selects the 12-warp form, kept for comparison)
// 25-32 columns.  0: 12 warps, 1 (default): + 8 drain warps, 2: + 8 drain warps and 16
// converter warps (measured slower than 1: the kernel runs at the board's power limit there and
// the extra warps cost clock).  Read per call: the tests compare the forms in one process.
static int k6_drain_warps() {
    const char* e = getenv("LTB200_K6_DW");
    const int v = e == nullptr ? 1 : atoi(e);
    return v < 0 ? 0 : v > 2 ? 2 : v;
}

// MMA-issuing warps: two with the drain-warp forms (+6 % at 25-32 columns at steady state), one
// otherwise (no gain up to 24 columns, -1.6 % at 19); LTB200_K6_ISSUERS=1|2 overrides
static int k6_issuers(bool drain_warps) {
    if (const char* e = getenv("LTB200_K6_ISSUERS")) return atoi(e) == 2 ? 2 : 1;
    return drain_warps ? 2 : 1;
}

// float32 tiles, N = 32 or 64: the lo(x) x lo(mask) products are dropped (LTB200_K6_THREE=0
// keeps all four); a quarter of the tensor work for a term below 2^-21 |x||m| with lo(mask)
// rounded to nearest, i.e. of either sign -- the measured error against float64 is unchanged to
// three digits, and at the power limit the kernel runs 3-4 % faster
static int k6_three(int n, bool is_float) {
    if (!is_float || (n != 32 && n != 64)) return 0;
    if (const char* e = getenv("LTB200_K6_THREE")) return atoi(e) != 0;
    return 1;
}

// sig_sum[k] += exact integer frame sum (rounded once to float32)
__global__ void k6_sig_finalize_kernel(const unsigned long long* __restrict__ acc,
                                       int64_t sig_size, float* __restrict__ sig_sum) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < sig_size) sig_sum[k] += (float)acc[k];
}

// one pass over the frames for <= 32 mask columns
template <typename TIN>
static int k6_run_group(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
                        const float* mk, int nm, int64_t ld_masks, float* o, int64_t ld_out,
                        int accumulate, int chain, float* sig_sum, uint8_t* ws, const K6Ws& wl,
                        cudaStream_t st) {
    constexpr int PX = K6_KS * K6In<TIN>::HALVES;
    const int sms = sm_count();
    const int nh = k6_nh(nm);
    const int n = 2 * nh;
    const int64_t sig_pad = ((sig_size + 63) / 64) * 64;
    const int64_t n_fb = (n_frames + K6_FB - 1) / K6_FB;

    K6Params p;
    p.n_frames = n_frames;
    p.sig_size = sig_size;
    p.n_masks = nm;
    p.ksplit = k6_choose_ksplit(n_fb, sig_size, sms, PX);
    const int64_t subs = (sig_size + PX - 1) / PX;
    p.k_per_split = ((subs + p.ksplit - 1) / p.ksplit) * PX;
    p.n_items = n_fb * p.ksplit;
    p.out = o;
    p.ld_out = ld_out;
    p.part = (float*)(ws + wl.part_off);
    p.accumulate = accumulate;
    p.chain = chain > 0 ? chain : k6_default_chain();
    p.debug = 0;
    p.sig_acc = nullptr;
    p.zero = 0u;
    p.issuers = k6_issuers(sizeof(TIN) == 4 && n == 64 && k6_drain_warps() > 0);
    p.three = k6_three(n, sizeof(TIN) == 4);
    if (const char* e = getenv("LTB200_K6_DEBUG")) p.debug = atoi(e);
    const int grid = (int)(p.n_items < sms ? p.n_items : sms);
    if (sig_sum != nullptr) {
        p.sig_acc = (unsigned long long*)(ws + wl.sig_off);
        LTB_CUDA_CHECK(cudaMemsetAsync(p.sig_acc, 0, (size_t)sig_size * 8, st));
    }

    float* packed = (float*)(ws + wl.pack_off);
    {
        const int64_t total = (int64_t)nh * sig_pad;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sms * 8) blocks = sms * 8;
        k6_pack_masks_kernel<<<blocks, 256, 0, st>>>(mk, nm, ld_masks, sig_size, sig_pad, nh,
                                                     packed);
        count_launch();
    }
    CUtensorMap tmd, tmm;
    int rc = encode_tmap_2d_sw(&tmd, tile,
                               sizeof(TIN) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                                : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                               (uint64_t)sig_size, (uint64_t)n_frames,
                               (uint64_t)ld_tile * sizeof(TIN), PX, K6_FB,
                               CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != LTB_OK) return rc;
    rc = encode_tmap_2d_sw(&tmm, packed, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (uint64_t)sig_pad,
                           (uint64_t)n, (uint64_t)sig_pad * 4, K6_KS, (uint32_t)n,
                           CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != LTB_OK) return rc;
    if constexpr (sizeof(TIN) == 2) {
        switch (n) {
            case 16: rc = k6_launch<16, TIN>(tmd, tmm, p, grid, st); break;
            default: rc = k6_launch<32, TIN>(tmd, tmm, p, grid, st); break;
        }
    } else {
        switch (n) {
            case 16: rc = k6_launch<16, TIN>(tmd, tmm, p, grid, st); break;
            case 32: rc = k6_launch<32, TIN>(tmd, tmm, p, grid, st); break;
            case 48: rc = k6_launch<48, TIN>(tmd, tmm, p, grid, st); break;
            default:
                switch (k6_drain_warps()) {
                    case 0: rc = k6_launch<64, TIN, 0>(tmd, tmm, p, grid, st); break;
                    case 1: rc = k6_launch<64, TIN, 1>(tmd, tmm, p, grid, st); break;
                    default: rc = k6_launch<64, TIN, 2>(tmd, tmm, p, grid, st); break;
                }
                break;
        }
    }
    if (rc != LTB_OK) return rc;
    if (p.ksplit > 1) {
        const int64_t total = n_frames * nm;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sms * 8) blocks = sms * 8;
        k6_finalize_kernel<<<blocks, 256, 0, st>>>(p.part, p.ksplit, n_frames, nm, o, ld_out,
                                                   accumulate);
        count_launch();
    }
    if (sig_sum != nullptr) {
        k6_sig_finalize_kernel<<<(unsigned)((sig_size + 255) / 256), 256, 0, st>>>(
            p.sig_acc, sig_size, sig_sum);
        count_launch();
    }
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

int k6_run(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
           const float* masks, int n_masks, int64_t ld_masks, float* out, int64_t ld_out,
           int accumulate, int chain, void* workspace, cudaStream_t st) {
    const K6Ws wl = k6_ws(n_frames, sig_size, n_masks, K6_KS, false);
    for (int m0 = 0; m0 < n_masks; m0 += 32) {
        const int nm = (n_masks - m0) > 32 ? 32 : (n_masks - m0);
        int rc = k6_run_group<float>(tile, n_frames, sig_size, ld_tile,
                                     masks + (int64_t)m0 * ld_masks, nm, ld_masks, out + m0,
                                     ld_out, accumulate, chain, nullptr, (uint8_t*)workspace, wl,
                                     st);
        if (rc != LTB_OK) return rc;
    }
    set_last_kernel(6);
    return LTB_OK;
}

// uint16 tiles, <= K6_U16_MAX_COLUMNS columns, optional fused frame sum (sig_sum += sum_f tile)
int k6_run_u16(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
               const float* masks, int n_masks, int64_t ld_masks, float* out, int64_t ld_out,
               int accumulate, int chain, float* sig_sum, void* workspace, cudaStream_t st) {
    const K6Ws wl = k6_ws(n_frames, sig_size, n_masks, 2 * K6_KS, sig_sum != nullptr);
    int rc = k6_run_group<uint16_t>(tile, n_frames, sig_size, ld_tile, masks, n_masks, ld_masks,
                                    out, ld_out, accumulate, chain, sig_sum, (uint8_t*)workspace,
                                    wl, st);
    if (rc != LTB_OK) return rc;
    set_last_kernel(6);
    return LTB_OK;
}

}  // namespace ltb

using namespace ltb;

extern "C" size_t ltb200_masks_dense_tc_workspace(int64_t n_frames, int64_t sig_size,
                                                  int n_masks) {
    if (n_frames <= 0 || sig_size <= 0 || n_masks <= 0) return 0;
    return k6_workspace(n_frames, sig_size, n_masks);
}

extern "C" int ltb200_masks_dense_tc(const float* tile, int64_t n_frames, int64_t sig_size,
                                     int64_t ld_tile, const float* masks, int n_masks,
                                     int64_t ld_masks, float* out, int64_t ld_out, int accumulate,
                                     int chain, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_size >= 0 && n_masks >= 0, "masks_dense_tc: negative size");
    if (n_frames == 0 || n_masks == 0) return LTB_OK;
    LTB_REQUIRE(tile != nullptr && masks != nullptr && out != nullptr,
                "masks_dense_tc: NULL pointer");
    LTB_REQUIRE(ld_tile >= sig_size && ld_masks >= sig_size && ld_out >= n_masks,
                "masks_dense_tc: leading dimension too small");
    if (!k6_shape_ok(tile, n_frames, sig_size, ld_tile)) {
        set_error("masks_dense_tc: tile shape/alignment not supported by the tensor-core path "
                  "(sig_size %lld, ld_tile %lld)", (long long)sig_size, (long long)ld_tile);
        return LTB_ERR_UNSUPPORTED;
    }
    const size_t need = k6_workspace(n_frames, sig_size, n_masks);
    if (need > workspace_bytes || workspace == nullptr) {
        set_error("masks_dense_tc: workspace of %zu B required, %zu B given", need,
                  workspace_bytes);
        return LTB_ERR_WORKSPACE;
    }
    return k6_run(tile, n_frames, sig_size, ld_tile, masks, n_masks, ld_masks, out, ld_out,
                  accumulate, chain, workspace, (cudaStream_t)stream);
}

extern "C" size_t ltb200_masks_dense_tc_u16_workspace(int64_t n_frames, int64_t sig_size,
                                                      int n_masks, int with_sig_sum) {
    if (n_frames <= 0 || sig_size <= 0 || n_masks <= 0) return 0;
    return k6_u16_workspace(n_frames, sig_size, n_masks, with_sig_sum);
}

extern "C" int ltb200_masks_dense_tc_u16(const uint16_t* tile, int64_t n_frames,
                                         int64_t sig_size, int64_t ld_tile, const float* masks,
                                         int n_masks, int64_t ld_masks, float* out,
                                         int64_t ld_out, int accumulate, int chain,
                                         float* sig_sum, void* workspace,
                                         size_t workspace_bytes, void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_size >= 0 && n_masks >= 0,
                "masks_dense_tc_u16: negative size");
    if (n_frames == 0 || n_masks == 0) return LTB_OK;
    LTB_REQUIRE(tile != nullptr && masks != nullptr && out != nullptr,
                "masks_dense_tc_u16: NULL pointer");
    LTB_REQUIRE(ld_tile >= sig_size && ld_masks >= sig_size && ld_out >= n_masks,
                "masks_dense_tc_u16: leading dimension too small");
    if (!k6_u16_shape_ok(tile, n_frames, sig_size, ld_tile, n_masks)) {
        set_error("masks_dense_tc_u16: shape not supported by the tensor-core path (sig_size "
                  "%lld, ld_tile %lld, %d columns; need sig_size %% 8 == 0, >= 256, <= %d "
                  "columns)", (long long)sig_size, (long long)ld_tile, n_masks,
                  K6_U16_MAX_COLUMNS);
        return LTB_ERR_UNSUPPORTED;
    }
    const size_t need = k6_u16_workspace(n_frames, sig_size, n_masks, sig_sum != nullptr);
    if (need > workspace_bytes || workspace == nullptr) {
        set_error("masks_dense_tc_u16: workspace of %zu B required, %zu B given", need,
                  workspace_bytes);
        return LTB_ERR_WORKSPACE;
    }
    return k6_run_u16(tile, n_frames, sig_size, ld_tile, masks, n_masks, ld_masks, out, ld_out,
                      accumulate, chain, sig_sum, workspace, (cudaStream_t)stream);
}
