// k7_group_tensor.cu -- K7: group-sparse masked reduction on the tensor cores
//                       (RadialFourierAnalysis hot path, tcgen05 form of K4)
//
// Replaces ApplyMasksUDF over radial_mask_factory masks (reference
// analysis/radialfourier.py:106-146: n_bins*(max_order+1) complex64 masks
// ring_b(r)*exp(i*o*phi), applied through the CSR rmatmul, udf/masks.py:68-69).
//
// All (max_order+1) masks of ring b share ONE support, so per ring the contraction is a dense
// GEMM  C_b[F x 2G] = I[:, ring pixels] . T_b  over the ring's pixels only (G complex columns =
// 2G real columns).  K4 runs that GEMM on the FP32 pipe (FFMA2) and is bound by it at 0.20 of
// the HBM roofline; here it runs on tcgen05.mma.kind::tf32 with the split-TF32 scheme of K6
// (k6_tensor.cu: hi = top 19 bits, lo = x - hi; exact products, float32 accumulate in TMEM cut
// into short chains that are summed in registers).
//
// Work item = (ring, block of 128 frames), ring-major so that all CTAs read the same slice of
// the weight table at the same time (the table -- N x n_entries floats, 131 MB for cfg4 -- is
// larger than L2, a ring's slice is 4 MB).  Items are fetched dynamically; their metadata
// travels with the pipeline stage.  13 warps:
//   * warps 0..3  producers: 4-byte cp.async gather of 128 frames x 64 ring entries per stage
//     (lanes walk the ring's ascending pixel list -> runs coalesce) into two 128-byte-swizzled
//     [128 x 32] sub-tiles, + TMA of the matching [N x 32] slices of the weight table (K-major,
//     128-byte swizzle = canonical UMMA layout); both complete on one mbarrier.
//   * warps 4..11 converters: thread <-> frame row (TMEM lane); warps 4..7 take entries 0..15
//     of a sub-tile, warps 8..11 entries 16..31: LDS.128, hi/lo split, tcgen05.st into a 4-slot
//     TMEM ring; they also drain the accumulators (each warp half of the columns) and store.
//   * warp 12     MMA issuer (warp-uniform loop, one elected lane): per sub-tile 4 k-steps x
//     (hi, lo) MMAs of M = 128, K = 8, N; tcgen05.commit frees the TMEM slot / the stage.
// Measured (B200, cfg4 geometry, 8192 frames): 3.9 ms vs 6.3 ms for K4 (0.34 vs 0.21 of the HBM
// roofline on the algorithmic bytes).  The bound is DRAM traffic, not the tensor pipe (24 %
// busy): B200 fills L2 from DRAM in 128-byte lines (cudaLimitMaxL2FetchGranularity has no
// effect), and a ring crosses an image row in runs of ~18 pixels, so a ring-by-ring gather moves
// 2.2-2.7x the bytes it uses; scheduling adjacent rings of the same frames back to back
// (rgroup) recovers part of it through L2.
// Weight-table rows are ordered [hi(cols 0..N/4) | lo(cols 0..N/4) | hi(cols N/4..N/2) | lo(..)]
// so that each half of the accumulator row holds the hi and lo parts of the same real columns.
#include "common.cuh"
#include <cstdlib>

namespace ltb {

constexpr int K7_FB = 128;            // frames per item (TMEM lanes)
constexpr int K7_KT = 64;             // ring entries per stage (2 sub-tiles of 32)
constexpr int K7_DSTAGES = 3;          // gathered data stages (ring size; p.dstages in use)
constexpr int K7_TSTAGES = 3;          // weight-table stages (ring size; p.tstages in use)
constexpr int K7_QLEN = 8;             // items queued from the gather producers to the table warp
constexpr int K7_AS = 4;              // TMEM operand ring (sub-tiles)
// gather modes (template parameter GM): 0 = cp.async.cg 16 B (4 producer warps), 1 = cp.async.ca
// (A/B, measured slower), 2 = LDG.128 into registers + STS.128 (8 producer warps, quad plans):
// scripts/ubench/gather_mix_probe.cu measures 2.0 vs 1.1 copies per clock per SM for 2 vs 0
// on L2-resident data.  Producer warps come in multiples of 4 (warp % 4 = TMEM lane quarter of
// the converter / drain warps that follow).
__host__ __device__ constexpr int k7_pwarps(int gm) { return gm == 2 ? 8 : 4; }
constexpr int K7_CWARPS = 8;
constexpr int K7_DWARPS = 4;          // accumulator drain warps, one per TMEM lane quarter
__host__ __device__ constexpr int k7_threads(int gm) {      // + MMA warp + table warp
    return (k7_pwarps(gm) + K7_CWARPS + K7_DWARPS + 2) * 32;
}
constexpr uint32_t K7_SUB_BYTES = K7_FB * 32 * 4;          // 16 KiB per sub-tile
constexpr uint32_t K7_DATA_BYTES = 2 * K7_SUB_BYTES;
constexpr int K7_TMEM_COLS = 512;
constexpr int K7_A_BASE = 256;        // TMEM columns [256, 512): 4 slots x (hi 32 | lo 32)

struct K7Params {
    const float* tile;
    int64_t n_frames, ld_tile;
    const int32_t* entry_px;       // (n_entries_padded): pixel index of every ring entry
    const int32_t* group_off;      // (n_groups + 1): entry offsets, multiples of K7_KT
    int n_groups, n_pairs;
    float* out;                    // (n_frames, ld_out) floats = complex64 (n_groups*n_pairs)
    int64_t ld_out;
    int accumulate;
    int64_t n_fb;                  // frame blocks
    int64_t n_items;
    int* counter;
    int chain;                     // sub-tiles per TMEM accumulation chain
    int rgroup;                    // adjacent rings scheduled back to back (L2 sharing)
    int fbgroup;                   // > 0: banded plan, frame blocks per scheduling group
    int n_bands;                   // banded plan: groups = n_bands x n_rings, band-major
    int prefetch;                  // banded plan: dense L2 prefetch of the next round's data
    int band_major;                // banded plan: item order [band][frame-block group][ring][fb]
    int quad;                      // entry_px lists QUADS (4 consecutive, 16-byte aligned pixels)
    int64_t sig_size;
    uint32_t zero;                 // 0, unknown to the compiler (stage release in the converters)
};

constexpr int K7_PF_PX = 256;         // pixels per L2-prefetch box (x 128 frames = 128 KiB)

// DRAM -> L2 only: the gathers that follow hit L2 (UTMAPF in SASS)
__device__ __forceinline__ void k7_tma_prefetch_2d(const CUtensorMap* map, int32_t c0,
                                                   int32_t c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map),
                 "r"(c0), "r"(c1)
                 : "memory");
}

// Item order: ring groups of `rgroup` adjacent rings outermost, frame blocks next, the rings of
// the group innermost.  CTAs that fetch consecutive items therefore gather ADJACENT rings of the
// SAME frames within microseconds of each other, so the 128-byte lines that straddle two rings
// are served by L2 instead of a second DRAM read (DRAM traffic 2.6x -> see DESIGN.md), while the
// weight-table slices in use at any time stay at rgroup x 4 MB.
__device__ __forceinline__ void k7_decode_item(const K7Params& p, int64_t item, int& g,
                                               int64_t& fb) {
    if (p.fbgroup > 0 && p.band_major) {
        // band outermost: the band's slice of the weight table (table / n_bands) stays
        // L2-resident for the whole pass over the frame blocks, and at any time the CTAs work on
        // ONE band of `fbgroup` frame blocks -- a working set small enough for L2, which the
        // producers prefetch densely one round ahead (see the prefetch in the producer loop)
        const int n_rings = p.n_groups / p.n_bands;
        const int64_t per_band = p.n_fb * n_rings;
        const int64_t b = item / per_band;
        const int64_t rem = item - b * per_band;
        const int64_t per = (int64_t)p.fbgroup * n_rings;
        const int64_t j = rem / per;
        const int64_t rem2 = rem - j * per;
        int64_t f_here = p.n_fb - j * p.fbgroup;
        if (f_here > p.fbgroup) f_here = p.fbgroup;
        const int64_t ring = rem2 / f_here;
        g = (int)(b * n_rings + ring);
        fb = j * p.fbgroup + (rem2 - ring * f_here);
        return;
    }
    if (p.fbgroup > 0) {
        // banded plan (groups = (pixel band, ring), band-major): `fbgroup` frame blocks
        // outermost, groups next, the frame blocks of the group innermost.  All CTAs then work
        // on the SAME pixel band of the same few frame blocks at the same time: a 128-byte line
        // of a frame row is fetched from DRAM by the first ring that touches it and served by
        // L2 to the other rings crossing it, and the weight-table slices of the band in use
        // stay L2-resident.
        const int64_t per = (int64_t)p.fbgroup * p.n_groups;
        const int64_t j = item / per;
        const int64_t rem = item - j * per;
        int64_t f_here = p.n_fb - j * p.fbgroup;
        if (f_here > p.fbgroup) f_here = p.fbgroup;
        const int64_t gg = rem / f_here;
        g = (int)gg;
        fb = j * p.fbgroup + (rem - gg * f_here);
        return;
    }
    const int64_t per_group = p.n_fb * p.rgroup;
    const int64_t j = item / per_group;
    const int64_t rem = item - j * per_group;
    int r_here = p.n_groups - (int)j * p.rgroup;
    if (r_here > p.rgroup) r_here = p.rgroup;
    fb = rem / r_here;
    g = (int)j * p.rgroup + (int)(rem - fb * r_here);
}

struct K7Meta {
    int item;      // < 0: no more work
    int chunk;     // stage index inside the item
    int nchunks;   // stages of the item
    int pad;
};

// ---- PTX wrappers (same forms as k6_tensor.cu) ----------------------------------------------
__device__ __forceinline__ void k7_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void k7_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void k7_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void k7_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void k7_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, "
        "%11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
        "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void k7_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void k7_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void k7_ld_fence4(uint32_t (&r)[4]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) : : "memory");
}
__device__ __forceinline__ void k7_ld_fence8(uint32_t (&r)[8]) {
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]),
                   "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
__device__ __forceinline__ void k7_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void k7_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool k7_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t k7_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t k7_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void k7_cp_async_4(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void k7_cp_async_16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc)
                 : "memory");
}
// A/B variant (LTB200_K7_CA=1): the quad gather through L1
__device__ __forceinline__ void k7_cp_async_16_ca(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void k7_cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// Shared memory: a ring of K7_DSTAGES gathered data stages and a separate, shorter ring of
// K7_TSTAGES weight-table stages.  The gathers are latency-bound (every stage is thousands of
// small copies through L1), so what matters is the number of data stages in flight; the table
// slices arrive by TMA from L2 and need no depth.
template <int NMMA>
struct K7Smem {
    static constexpr uint32_t TABLE_BYTES = 2u * NMMA * 128u;              // two [NMMA x 32] slices
    static constexpr uint32_t TABLE_OFF = K7_DSTAGES * K7_DATA_BYTES;
    static constexpr uint32_t BAR_OFF = TABLE_OFF + K7_TSTAGES * TABLE_BYTES;
    static constexpr uint32_t TOTAL = BAR_OFF + 1024 + 1024;               // + alignment slack
};

struct K7QItem {
    int item;      // < 0: no more work
    int e0;        // first entry of the item's group
    int nchunks;
    int pad;
};

// SYM (mirror-symmetric plan, N = 128): a stage holds 32 ORBITS -- sub-tile 0 the pixels p,
// sub-tile 1 their mirror images p' (same columns, row sy - y) -- whose masks obey
// m(p') = conj(m(p)).  The converters form I(p) + I(p') and I(p) - I(p'); the sums meet the REAL
// parts of the weights in accumulator columns [0, 64), the differences the IMAGINARY parts in
// [64, 128): two MMAs of N = 64 per k-step instead of two of N = 112 per pixel pair, and a
// weight table of 128 rows per orbit instead of 112 per pixel.
__device__ __forceinline__ uint4 k7_ldg128(const void* src) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(src));
    return v;
}
__device__ __forceinline__ void k7_sts128(uint32_t dst, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

template <int N, bool SYM, int GM = 0>
__global__ void __launch_bounds__(k7_threads(GM), 1)
k7_group_tensor_kernel(const __grid_constant__ CUtensorMap tm_table,
                       const __grid_constant__ CUtensorMap tm_tile, const K7Params p) {
    constexpr int NMMA = SYM ? 64 : N;    // columns of one MMA
    constexpr int K7_PWARPS = k7_pwarps(GM);
    constexpr bool CA = GM == 1;
    using SM = K7Smem<NMMA>;
    constexpr int NHALF = N / 2;          // accumulator columns drained per converter warp
    constexpr int NQ = N / 4;             // real columns per half
    constexpr uint32_t IDESC = k7_idesc_tf32(NMMA);
    static_assert(N % 16 == 0 && N >= 16 && N <= 128, "K7: N in 16..128 step 16");
    static_assert(!SYM || N == 128, "K7: the symmetric plan uses 2 x 64 accumulator columns");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

    uint64_t* data_full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);  // [DSTAGES]
    uint64_t* data_free = data_full + K7_DSTAGES;                           // [DSTAGES]
    uint64_t* tab_full = data_free + K7_DSTAGES;                            // [TSTAGES]
    uint64_t* tab_free = tab_full + K7_TSTAGES;                             // [TSTAGES]
    uint64_t* a_full = tab_free + K7_TSTAGES;                               // [AS]
    uint64_t* mma_done = a_full + K7_AS;                                    // [AS]
    uint64_t* acc_full = mma_done + K7_AS;                                  // [2]
    uint64_t* acc_free = acc_full + 2;                                      // [2]
    uint64_t* q_full = acc_free + 2;                                        // [QLEN]
    uint64_t* q_free = q_full + K7_QLEN;                                    // [QLEN]
    K7Meta* meta = reinterpret_cast<K7Meta*>(q_free + K7_QLEN);             // [DSTAGES]
    K7Meta* tmeta = meta + K7_DSTAGES;                                      // [TSTAGES]
    K7QItem* queue = reinterpret_cast<K7QItem*>(tmeta + K7_TSTAGES);        // [QLEN]
    int* cur_item = reinterpret_cast<int*>(queue + K7_QLEN);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cur_item + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    constexpr int DRAIN_WARP0 = K7_PWARPS + K7_CWARPS;        // a multiple of 4: warp % 4 =
    constexpr int MMA_WARP = DRAIN_WARP0 + K7_DWARPS;         // TMEM lane quarter
    constexpr int TABLE_WARP = MMA_WARP + 1;
    static_assert(DRAIN_WARP0 % 4 == 0 && K7_DWARPS == 4, "K7: one drain warp per lane quarter");

    if (threadIdx.x == 0) {
        for (int s = 0; s < K7_DSTAGES; s++) {
            mbar_init(&data_full[s], 1 + K7_PWARPS * 32);
            mbar_init(&data_free[s], K7_CWARPS);
        }
        for (int s = 0; s < K7_TSTAGES; s++) {
            mbar_init(&tab_full[s], 1);
            mbar_init(&tab_free[s], 1);
        }
        for (int s = 0; s < K7_AS; s++) {
            mbar_init(&a_full[s], K7_CWARPS);
            mbar_init(&mma_done[s], 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_free[s], K7_DWARPS);
        }
        for (int s = 0; s < K7_QLEN; s++) {
            mbar_init(&q_full[s], 1);
            mbar_init(&q_free[s], 1 + K7_DWARPS);     // table warp + the drain warps read it
        }
        fence_mbar_init();
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "n"(K7_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    k7_fence_before();
    __syncthreads();
    k7_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int chain = p.chain;

    if (warp < K7_PWARPS) {
        // ===== gather producers =====
        const int pt = threadIdx.x;                       // 0 .. 32 * K7_PWARPS - 1
        const int el = pt & 31;                           // entry inside the sub-tile
        const int sub = (pt >> 5) & 1;                    // sub-tile of the stage
        constexpr int FPT = K7_FB / (K7_PWARPS / 2);      // frames per thread
        const int fq = pt >> 6;                           // rows fq*FPT .. +FPT-1
        // 128-byte swizzle: 16-byte chunk (el / 4) of row r lands at chunk ^ (r & 7).  The rows
        // of a thread start at a multiple of 8, so the offset pattern repeats every 8 rows:
        // eight precomputed offsets + an immediate per copy.
        uint32_t sw[8];
#pragma unroll
        for (int c = 0; c < 8; c++)
            sw[c] = (uint32_t)sub * K7_SUB_BYTES + (uint32_t)(fq * FPT + c) * 128u +
                    ((((uint32_t)(el >> 2)) ^ (uint32_t)c) << 4) + (uint32_t)(el & 3) * 4u;
        // quad plan: 16 quads per stage; thread = (quad qd of sub-tile qsub, rows rg + 8 j)
        const int qd = pt & 7, qsub = (pt >> 3) & 1, rg = pt >> 4;
        uint32_t qn = 0;
        int stage = 0;                 // data ring position (no division in the stage loop)
        uint32_t dphase = 0;
        while (true) {
            if (pt == 0) *cur_item = atomicAdd(p.counter, 1);
            named_bar_sync(2, K7_PWARPS * 32);
            const int item = *cur_item;
            named_bar_sync(2, K7_PWARPS * 32);
            const bool done = item >= p.n_items;
            int64_t fb = 0;
            int e0 = 0, nchunks = 1;
            if (!done) {
                int g;
                k7_decode_item(p, item, g, fb);
                if (p.prefetch && pt == 0) {
                    // dense DRAM -> L2 prefetch, one round ahead: the items of a round
                    // (band, frame-block group) share out the NEXT round's data -- item
                    // (ring r, frame block f of the group) prefetches, for the matching frame
                    // block of the next round, every n_rings-th 256-pixel box of the band.
                    // The gathers of the next round then hit L2 (dense 1 KiB rows from DRAM
                    // instead of scattered 128-byte lines).
                    const int n_rings = p.n_groups / p.n_bands;
                    int nb = g / n_rings;
                    const int r = g % n_rings;
                    int64_t pfb;
                    if (p.band_major) {
                        pfb = fb + p.fbgroup;               // same band, next frame-block group
                        if ((fb / p.fbgroup + 1) * p.fbgroup >= p.n_fb) {
                            nb += 1;                        // first group of the next band
                            pfb = fb % p.fbgroup;
                        }
                    } else {
                        nb += 1;                            // next band of the same frame blocks
                        pfb = fb;
                        if (nb == p.n_bands) {
                            nb = 0;
                            pfb = fb + p.fbgroup;
                        }
                    }
                    if (pfb < p.n_fb && nb < p.n_bands) {
                        const int64_t b0 = (p.sig_size * nb) / p.n_bands;
                        const int64_t b1 = (p.sig_size * (nb + 1)) / p.n_bands;
                        for (int64_t x = b0 + (int64_t)r * K7_PF_PX; x < b1;
                             x += (int64_t)n_rings * K7_PF_PX)
                            k7_tma_prefetch_2d(&tm_tile, (int32_t)x, (int32_t)(pfb * K7_FB));
                    }
                }
                e0 = p.group_off[g];
                nchunks = (p.group_off[g + 1] - e0) / K7_KT;
                if (nchunks == 0) continue;   // empty ring: nothing to add
            }
            if (pt == 0) {
                // hand the item to the weight-table warp and the drain warps
                const int q = qn % K7_QLEN;
                mbar_wait(&q_free[q], ((qn / K7_QLEN) & 1) ^ 1);
                queue[q] = K7QItem{done ? -1 : item, e0, nchunks, 0};
                mbar_arrive(&q_full[q]);
            }
            qn++;
            const int64_t f_first = fb * K7_FB + fq * FPT;
            const bool ragged = fb * K7_FB + K7_FB > p.n_frames;
            // the pixel index of this thread's entry is loaded one stage ahead (an L2 round
            // trip that would otherwise sit between the stage becoming free and its copies)
            int px_next = done ? 0
                               : (p.quad ? p.entry_px[(e0 >> 2) + qsub * 8 + qd]
                                         : p.entry_px[e0 + sub * 32 + el]);
            for (int c = 0; c < nchunks; c++) {
                mbar_wait(&data_free[stage], dphase ^ 1);
                uint8_t* dst = smem + (size_t)stage * K7_DATA_BYTES;
                uint64_t* full = &data_full[stage];
                if (pt == 0) {
                    // sentinel stage (done): tells the converters to stop (all arrivals, no data)
                    meta[stage] = K7Meta{done ? -1 : item, c, nchunks, 0};
                    mbar_arrive(full);
                }
                if (++stage == K7_DSTAGES) {
                    stage = 0;
                    dphase ^= 1;
                }
                if (done) {
                    mbar_arrive(full);
                    continue;
                }
                const int ebase = e0 + c * K7_KT;
                const int px = px_next;
                if (c + 1 < nchunks)
                    px_next = p.quad ? p.entry_px[((ebase + K7_KT) >> 2) + qsub * 8 + qd]
                                     : p.entry_px[ebase + K7_KT + sub * 32 + el];
                const uint32_t sdst = smem_u32(dst);
                if (p.quad) {
                    // one 16-byte copy moves 4 consecutive pixels of a frame into one swizzled
                    // chunk: a quarter of the copy instructions and of the shared-memory write
                    // wavefronts of the 4-byte gather
                    constexpr int RS = K7_PWARPS * 2;          // frame rows per copy step
                    const uint32_t d0 = sdst + (uint32_t)qsub * K7_SUB_BYTES + (uint32_t)rg * 128u +
                                        ((uint32_t)(qd ^ (rg & 7)) << 4);
                    const int64_t fr0 = fb * K7_FB + rg;
                    if constexpr (GM == 2) {
                        // synchronous gather: all copies of the thread in flight as LDG.128,
                        // then STS.128 into the swizzled stage; the plain arrive below releases
                        // the stores to the converters
                        uint4 v[K7_FB / RS];
                        const float* src = p.tile + px;
#pragma unroll
                        for (int j = 0; j < K7_FB / RS; j++) {
                            int64_t fr = fr0 + RS * j;
                            if (ragged && fr >= p.n_frames) fr = p.n_frames - 1;
                            v[j] = k7_ldg128(src + fr * p.ld_tile);
                        }
#pragma unroll
                        for (int j = 0; j < K7_FB / RS; j++)
                            k7_sts128(d0 + (uint32_t)j * (RS * 128u), v[j]);
                    } else if (!ragged) {
                        const float* src = p.tile + px + fr0 * p.ld_tile;
                        const int64_t step = RS * p.ld_tile;
#pragma unroll 4
                        for (int j = 0; j < K7_FB / RS; j++) {
                            if constexpr (CA)
                                k7_cp_async_16_ca(d0 + (uint32_t)j * (RS * 128u), src);
                            else
                                k7_cp_async_16(d0 + (uint32_t)j * (RS * 128u), src);
                            src += step;
                        }
                    } else {
#pragma unroll 4
                        for (int j = 0; j < K7_FB / RS; j++) {
                            int64_t fr = fr0 + RS * j;
                            if (fr >= p.n_frames) fr = p.n_frames - 1;
                            k7_cp_async_16(d0 + (uint32_t)j * (RS * 128u),
                                           p.tile + px + fr * p.ld_tile);
                        }
                    }
                } else if (!ragged) {
                    const float* src = p.tile + px + f_first * p.ld_tile;
#pragma unroll 1
                    for (int f8 = 0; f8 < FPT / 8; f8++) {
#pragma unroll
                        for (int c8 = 0; c8 < 8; c8++) {
                            k7_cp_async_4(sdst + sw[c8] + (uint32_t)f8 * 1024u, src);
                            src += p.ld_tile;
                        }
                    }
                } else {
                    // last frame block: rows past the end re-read the last frame (never stored)
                    const float* src = p.tile + px;
#pragma unroll 1
                    for (int f8 = 0; f8 < FPT / 8; f8++) {
#pragma unroll
                        for (int c8 = 0; c8 < 8; c8++) {
                            int64_t fr = f_first + f8 * 8 + c8;
                            if (fr >= p.n_frames) fr = p.n_frames - 1;
                            k7_cp_async_4(sdst + sw[c8] + (uint32_t)f8 * 1024u,
                                          src + fr * p.ld_tile);
                        }
                    }
                }
                if constexpr (GM == 2)
                    mbar_arrive(full);
                else
                    k7_cp_async_mbar_arrive_noinc(full);
            }
            if (done) break;
        }
    } else if (warp == TABLE_WARP) {
        // ===== weight-table producer: TMA of the [N x 32] table slices of every stage =====
        if (lane == 0) {
            prefetch_tmap(&tm_table);
            const uint64_t pol_keep = l2_policy_evict_last();
            int ts = 0;
            uint32_t tphase = 0;
            for (uint32_t qn = 0;; qn++) {
                const int q = qn % K7_QLEN;
                mbar_wait(&q_full[q], (qn / K7_QLEN) & 1);
                const K7QItem qi = queue[q];
                mbar_arrive(&q_free[q] + ((uint32_t)qi.item & p.zero));
                const int nch = qi.item < 0 ? 1 : qi.nchunks;
                for (int c = 0; c < nch; c++) {
                    mbar_wait(&tab_free[ts], tphase ^ 1);
                    uint8_t* dst = smem + SM::TABLE_OFF + (size_t)ts * SM::TABLE_BYTES;
                    uint64_t* full = &tab_full[ts];
                    tmeta[ts] = K7Meta{qi.item, c, qi.nchunks, 0};
                    if (++ts == K7_TSTAGES) {
                        ts = 0;
                        tphase ^= 1;
                    }
                    if (qi.item < 0) {
                        mbar_arrive(full);
                        continue;
                    }
                    const int ebase = qi.e0 + c * K7_KT;
                    mbar_arrive_expect_tx(full, SM::TABLE_BYTES);
                    if constexpr (SYM) {
                        // 32 orbits: rows [0, 64) = real parts (hi | lo), [64, 128) = imaginary
                        tma_load_2d(dst, &tm_table, ebase >> 1, 0, full, pol_keep);
                        tma_load_2d(dst + NMMA * 128, &tm_table, ebase >> 1, 64, full, pol_keep);
                    } else {
                        tma_load_2d(dst, &tm_table, ebase, 0, full, pol_keep);
                        tma_load_2d(dst + N * 128, &tm_table, ebase + 32, 0, full, pol_keep);
                    }
                }
                if (qi.item < 0) break;
            }
        }
    } else if (warp == MMA_WARP) {
        // ===== MMA issuer =====
        // Accumulators: two buffers of N TMEM columns; a CHAIN of `chain` sub-tiles accumulates
        // into one of them and is then handed to the drain warps (acc_full / acc_free), so the
        // float32 accumulate in the tensor core -- which truncates -- never runs long.
        int ts = 0, as = 0, cbuf = 0, in_chain = 0;
        uint32_t tphase = 0, aphase = 0, cuse = 0;      // cuse bit b: parity of the uses of buffer b
        for (;;) {
            mbar_wait(&tab_full[ts], tphase);
            const K7Meta m = tmeta[ts];
            if (m.item < 0) break;
            const bool last_stage = m.chunk == m.nchunks - 1;
            const uint32_t tbl = smem_u32(smem + SM::TABLE_OFF + (size_t)ts * SM::TABLE_BYTES);
#pragma unroll
            for (int s = 0; s < 2; s++) {
                if (in_chain == 0) {
                    // the drain warps have read the previous chain out of this buffer
                    mbar_wait(&acc_free[cbuf], ((cuse >> cbuf) & 1) ^ 1);
                }
                mbar_wait(&a_full[as], aphase);
                k7_fence_after();
                const uint64_t bdesc0 = k7_desc_k_sw128(tbl + (uint32_t)s * NMMA * 128u);
                const uint32_t a0 = tmem_base + (uint32_t)(K7_A_BASE + as * 64);
                // SYM: sub-tile 0 (sums) -> columns [0, 64), sub-tile 1 (differences) -> [64, 128);
                // a chain starts with the first STAGE (both sub-tiles zero-initialise)
                const uint32_t d0 = tmem_base + (uint32_t)(cbuf * N + (SYM ? s * 64 : 0));
                const int first = SYM ? (in_chain < 2 ? 0 : 1) : in_chain;
                in_chain++;
                const bool chain_done = in_chain == chain || (last_stage && s == 1);
                if (k7_elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const uint64_t bdesc = bdesc0 + (uint64_t)(kk * 2);
                        k7_mma_tf32_ts(d0, a0 + kk * 8, bdesc, IDESC, (first | kk) != 0 ? 1u : 0u);
                        k7_mma_tf32_ts(d0, a0 + 32 + kk * 8, bdesc, IDESC, 1u);
                    }
                    k7_commit(&mma_done[as]);
                    if (s == 1) k7_commit(&tab_free[ts]);
                    if (chain_done) k7_commit(&acc_full[cbuf]);
                }
                __syncwarp();
                if (chain_done) {
                    cuse ^= 1u << cbuf;
                    cbuf ^= 1;
                    in_chain = 0;
                }
                if (++as == K7_AS) {
                    as = 0;
                    aphase ^= 1;
                }
            }
            if (++ts == K7_TSTAGES) {
                ts = 0;
                tphase ^= 1;
            }
        }
    } else if (warp >= DRAIN_WARP0) {
        // ===== accumulator drain: chain totals -> float32 registers (round to nearest) -> out
        // One warp per TMEM lane quarter; thread <-> frame row.  The item sequence comes from
        // the same queue the table warp reads, the chain segmentation is the MMA warp's.
        const int dw = warp - DRAIN_WARP0;
        const int row = dw * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(dw * 32) << 16;
        int cbuf = 0;
        uint32_t duse = 0;
        for (uint32_t qn = 0;; qn++) {
            const int q = qn % K7_QLEN;
            mbar_wait(&q_full[q], (qn / K7_QLEN) & 1);
            const K7QItem qi = queue[q];
            __syncwarp();
            if (lane == 0) mbar_arrive(&q_free[q] + ((uint32_t)qi.item & p.zero));
            if (qi.item < 0) break;
            float acc[2][NQ];
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int c = 0; c < NQ; c++) acc[h][c] = 0.f;
            const int n_sub = 2 * qi.nchunks;
            for (int done = 0; done < n_sub; done += chain) {
                mbar_wait(&acc_full[cbuf], (duse >> cbuf) & 1);
                k7_fence_after();
                const uint32_t d = tmem_base + lane_sel + (uint32_t)(cbuf * N);
#pragma unroll
                for (int h = 0; h < 2; h++) {
#pragma unroll
                    for (int jb0 = 0; jb0 < NQ / 4; jb0 += 4) {
                        uint32_t vh[4][4], vl[4][4];
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (jb0 + u < NQ / 4) {
                                k7_ld4(d + h * NHALF + (jb0 + u) * 4, vh[u]);
                                k7_ld4(d + h * NHALF + NQ + (jb0 + u) * 4, vl[u]);
                            }
                        k7_wait_ld();
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (jb0 + u < NQ / 4) {
                                k7_ld_fence4(vh[u]);
                                k7_ld_fence4(vl[u]);
#pragma unroll
                                for (int t = 0; t < 4; t++)
                                    acc[h][(jb0 + u) * 4 + t] +=
                                        __uint_as_float(vh[u][t]) + __uint_as_float(vl[u][t]);
                            }
                    }
                }
                k7_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_free[cbuf]);
                duse ^= 1u << cbuf;
                cbuf ^= 1;
            }
            int g;
            int64_t fb;
            k7_decode_item(p, qi.item, g, fb);
            const int64_t f = fb * K7_FB + row;
            if (f < p.n_frames) {
                float* o = p.out + f * p.ld_out + (int64_t)g * p.n_pairs * 2;
                if constexpr (SYM) {
                    // half 0 holds the real parts, half 1 the imaginary parts
#pragma unroll
                    for (int c = 0; c < NQ; c++)
                        if (c < p.n_pairs) {
                            o[2 * c] = p.accumulate ? (o[2 * c] + acc[0][c]) : acc[0][c];
                            o[2 * c + 1] = p.accumulate ? (o[2 * c + 1] + acc[1][c]) : acc[1][c];
                        }
                } else {
#pragma unroll
                    for (int h = 0; h < 2; h++)
#pragma unroll
                        for (int c = 0; c < NQ; c++)
                            if (h * NQ + c < 2 * p.n_pairs)
                                o[h * NQ + c] =
                                    p.accumulate ? (o[h * NQ + c] + acc[h][c]) : acc[h][c];
                }
            }
        }
    } else {
        // ===== converters: gathered float32 pixels -> split-TF32 A operand in TMEM =====
        // thread <-> frame row (TMEM lane); warps 4..7 take entries 0..15 of a sub-tile, warps
        // 8..11 entries 16..31.  Nothing else happens here: this loop is what bounds the kernel
        // (profiles/r2_k7_*), so ring positions are wrapping counters and the accumulator
        // drain lives in its own warps.
        const int cw = warp - K7_PWARPS;
        const int w = cw & 3;                       // TMEM lane quarter (== warp % 4)
        const int hh = cw >> 2;
        const int row = w * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(w * 32) << 16;
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t row_off = (uint32_t)row * 128u;
        int stage = 0, as = 0;
        uint32_t dphase = 0, aphase = 0;
        for (;;) {
            mbar_wait(&data_full[stage], dphase);
            if (meta[stage].item < 0) break;
            const uint8_t* dbase = smem + (size_t)stage * K7_DATA_BYTES + row_off;
            float4 x[2][4];
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    x[s][j] = *reinterpret_cast<const float4*>(
                        dbase + s * K7_SUB_BYTES + (((uint32_t)(hh * 4 + j) ^ swz) << 4));
            // the stage is handed back to the gather producers as soon as the loads have
            // RETURNED: the barrier address of the arrive depends on one word of every LDS.128
            // (`p.zero` is 0, but only at run time), which puts the scoreboard wait of the loads
            // in front of the arrive.  An empty asm that only names the registers does not --
            // the arrive could overtake the loads (found with K10, csrc/k10_walk.cu).
            uint32_t dep = 0;
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int j = 0; j < 4; j++) dep ^= __float_as_uint(x[s][j].x);
            dep &= p.zero;
            __syncwarp();
            if (lane == 0) mbar_arrive(&data_free[stage] + dep);
            if (++stage == K7_DSTAGES) {
                stage = 0;
                dphase ^= 1;
            }
            if constexpr (SYM) {
                // butterflies of the mirror pairs: sub-tile 0 <- I(p) + I(p'), 1 <- I(p) - I(p')
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 u = x[0][j], v = x[1][j];
                    x[0][j] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
                    x[1][j] = make_float4(u.x - v.x, u.y - v.y, u.z - v.z, u.w - v.w);
                }
            }
            // both sub-tiles are written before ONE tcgen05.wait::st: the store latency is paid
            // once per stage
            int as_used[2];
#pragma unroll
            for (int s = 0; s < 2; s++) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float e[4] = {x[s][j].x, x[s][j].y, x[s][j].z, x[s][j].w};
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const uint32_t hb = __float_as_uint(e[t]) & 0xFFFFE000u;
                        hi[j * 4 + t] = hb;
                        lo[j * 4 + t] = __float_as_uint(e[t] - __uint_as_float(hb)) + 0x1000u;
                    }
                }
                // slot `as` is free once the MMAs of the sub-tile K7_AS back have completed
                mbar_wait(&mma_done[as], aphase ^ 1);
                k7_fence_after();
                const uint32_t a = tmem_base + lane_sel + (uint32_t)(K7_A_BASE + as * 64 + hh * 16);
                k7_st16(a, hi);
                k7_st16(a + 32, lo);
                as_used[s] = as;
                if (++as == K7_AS) {
                    as = 0;
                    aphase ^= 1;
                }
            }
            k7_wait_st();
            k7_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a_full[as_used[0]]);
                mbar_arrive(&a_full[as_used[1]]);
            }
        }
    }

    k7_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        k7_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(K7_TMEM_COLS)
                     : "memory");
    }
}

template <int N, bool SYM, int GM = 0>
static int k7_launch(const CUtensorMap& tm, const CUtensorMap& tmt, const K7Params& p, int grid,
                     cudaStream_t st) {
    auto kern = k7_group_tensor_kernel<N, SYM, GM>;
    const size_t smem = K7Smem<SYM ? 64 : N>::TOTAL;
    int dev = 0;
    LTB_CUDA_CHECK(cudaGetDevice(&dev));
    static thread_local int configured_dev = -1;
    if (configured_dev != dev) {
        LTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        configured_dev = dev;
    }
    kern<<<grid, k7_threads(GM), smem, st>>>(tm, tmt, p);
    count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

}  // namespace ltb

using namespace ltb;

extern "C" int ltb200_group_masks_tc_columns(int n_pairs) {
    // accumulator columns N (rows of the split weight table) for n_pairs complex columns
    const int need = 4 * n_pairs;             // (re, im) x (hi, lo)
    if (n_pairs < 1 || need > 112) return 0;
    if (need <= 16) return 16;
    if (need <= 32) return 32;
    if (need <= 64) return 64;
    return 112;
}

// out[f, ring, c] (+)= sum over the non-empty bands of part[f, band * n_rings + ring, c]
// (fixed order: deterministic)
__global__ void k7_band_reduce_kernel(const float* __restrict__ part, const int32_t* group_off,
                                      int64_t n_frames, int n_rings, int n_bands, int cols,
                                      float* __restrict__ out, int64_t ld_out, int accumulate) {
    const int64_t row = (int64_t)n_rings * cols;
    const int64_t total = n_frames * row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / row;
        const int rc = (int)(i - f * row);
        const int ring = rc / cols;
        const float* src = part + f * row * n_bands + rc;
        float s = 0.f;
        for (int b = 0; b < n_bands; b++) {
            const int g = b * n_rings + ring;
            if (group_off[g + 1] > group_off[g]) s += src[(int64_t)b * row];
        }
        float* o = out + f * ld_out + rc;
        *o = accumulate ? (*o + s) : s;
    }
}

extern "C" size_t ltb200_group_masks_tc_workspace(int64_t n_frames, int n_groups, int n_pairs,
                                                  int n_bands) {
    if (n_bands <= 1) return 256;
    return 256 + (size_t)n_frames * (size_t)n_groups * (size_t)n_pairs * 2 * sizeof(float);
}

static int k7_run(const float* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
                  const int32_t* entry_px, const float* table_split,
                  const int32_t* group_off_host, const int32_t* group_off_dev, int n_groups,
                  int n_pairs, int n_bands, int quad_plan, int sym, float* out, int64_t ld_out,
                  int accumulate, int chain, void* workspace, size_t workspace_bytes,
                  void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_size > 0 && n_groups > 0, "group_masks_tc: bad sizes");
    LTB_REQUIRE(n_bands >= 1 && n_groups % n_bands == 0,
                "group_masks_tc: n_groups must be n_bands x rings");
    LTB_REQUIRE(!sym || quad_plan, "group_masks_tc: the symmetric plan is a quad plan");
    // symmetric plan: 2 x (hi 32 | lo 32) accumulator columns, table of 128 rows per orbit
    const int n = sym ? (n_pairs >= 1 && n_pairs <= 28 ? 128 : 0)
                      : ltb200_group_masks_tc_columns(n_pairs);
    LTB_REQUIRE(n > 0, "group_masks_tc: 1..28 complex columns per group, got %d", n_pairs);
    if (n_frames == 0) return LTB_OK;
    LTB_REQUIRE(tile && entry_px && table_split && group_off_host && group_off_dev && out,
                "group_masks_tc: NULL pointer");
    const size_t need = ltb200_group_masks_tc_workspace(n_frames, n_groups, n_pairs, n_bands);
    LTB_REQUIRE(workspace != nullptr && workspace_bytes >= need,
                "group_masks_tc: workspace of %zu B required, %zu B given", need,
                workspace_bytes);
    const int n_rings = n_groups / n_bands;
    LTB_REQUIRE(ld_out >= (int64_t)n_rings * n_pairs * 2, "group_masks_tc: ld_out too small");
    LTB_REQUIRE((uintptr_t)table_split % 16 == 0, "group_masks_tc: table must be 16 B aligned");
    LTB_REQUIRE((uintptr_t)tile % 4 == 0, "group_masks_tc: tile must be 4 B aligned");
    const int64_t n_entries = group_off_host[n_groups];
    for (int g = 0; g <= n_groups; g++)
        LTB_REQUIRE(group_off_host[g] % K7_KT == 0,
                    "group_masks_tc: offsets must be multiples of %d", K7_KT);
    LTB_REQUIRE(n_entries > 0 && n_entries % 4 == 0, "group_masks_tc: empty entry list");
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tm;
    // table: (n, n_entries) floats; symmetric plan: (128, n_entries / 2) -- one column per orbit,
    // fetched as two [64 x 32] boxes (real rows, imaginary rows) per stage
    const uint64_t tcols = sym ? (uint64_t)n_entries / 2 : (uint64_t)n_entries;
    int rc = encode_tmap_2d_sw(&tm, table_split, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, tcols,
                               (uint64_t)n, tcols * 4, 32, (uint32_t)(sym ? 64 : n),
                               CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != LTB_OK) return rc;
    const bool banded = n_bands > 1;     // band partial sums + reduction
    const bool quad = quad_plan != 0;    // entry_px lists quads (16-byte gathers)
    LTB_REQUIRE(quad || !banded, "group_masks_tc: banded plans are quad plans");
    float* part = (float*)((uint8_t*)workspace + 256);
    K7Params p;
    p.tile = tile;
    p.n_frames = n_frames;
    p.ld_tile = ld_tile;
    p.entry_px = entry_px;
    p.group_off = group_off_dev;
    p.n_groups = n_groups;
    p.n_pairs = n_pairs;
    p.out = banded ? part : out;
    p.ld_out = banded ? (int64_t)n_groups * n_pairs * 2 : ld_out;
    p.accumulate = banded ? 0 : accumulate;
    p.n_fb = (n_frames + K7_FB - 1) / K7_FB;
    p.n_items = p.n_fb * n_groups;
    p.counter = (int*)workspace;
    p.chain = chain > 0 ? chain : 2;
    if (chain <= 0)
        if (const char* e = getenv("LTB200_K7_CHAIN"))
            if (atoi(e) > 0) p.chain = atoi(e);
    if (sym) p.chain = ((p.chain + 1) / 2) * 2;     // whole stages per chain
    p.rgroup = 4;
    if (const char* e = getenv("LTB200_K7_RGROUP"))
        if (atoi(e) > 0) p.rgroup = atoi(e);
    if (p.rgroup > n_groups) p.rgroup = n_groups;
    p.fbgroup = 0;
    p.quad = quad ? 1 : 0;
    if (quad)
        LTB_REQUIRE((uintptr_t)tile % 16 == 0 && ld_tile % 4 == 0,
                    "group_masks_tc: the quad plan needs 16-byte aligned frame rows");
    p.n_bands = n_bands;
    p.prefetch = 0;
    p.band_major = 0;
    p.sig_size = sig_size;
    p.zero = 0u;
    CUtensorMap tmt = tm;
    if (quad) {
        p.fbgroup = 8;
        if (const char* e = getenv("LTB200_K7_FBG"))
            if (atoi(e) > 0) p.fbgroup = atoi(e);
        if (const char* e = getenv("LTB200_K7_BM")) p.band_major = atoi(e) != 0;
        int want_pf = 0;
        if (const char* e = getenv("LTB200_K7_PF")) want_pf = atoi(e);
        if (want_pf && (uintptr_t)tile % 16 == 0 && ld_tile % 4 == 0 && sig_size >= K7_PF_PX &&
            sig_size < (1ll << 31)) {
            rc = encode_tmap_2d_sw(&tmt, tile, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (uint64_t)sig_size,
                                   (uint64_t)n_frames, (uint64_t)ld_tile * 4, K7_PF_PX, K7_FB,
                                   CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc != LTB_OK) return rc;
            p.prefetch = 1;
        }
    }
    LTB_REQUIRE(p.n_items < (1ll << 31), "group_masks_tc: too many work items");
    LTB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 4, st));
    if (!accumulate && !banded) {
        // rings without entries leave their columns untouched: define them as zero
        LTB_CUDA_CHECK(cudaMemset2DAsync(out, ld_out * sizeof(float), 0,
                                         (size_t)n_groups * n_pairs * 2 * sizeof(float), n_frames,
                                         st));
    }
    int grid = sm_count();
    if (p.n_items < grid) grid = (int)p.n_items;
    int gm = 0;                             // gather mode of the quad plans (see k7_pwarps)
    if (const char* e = getenv("LTB200_K7_GM")) gm = atoi(e);
    if (const char* e = getenv("LTB200_K7_CA"))
        if (atoi(e) != 0) gm = 1;
    if (gm != 0 && quad && (sym || n == 112)) {
        if (gm == 2)
            rc = sym ? k7_launch<128, true, 2>(tm, tmt, p, grid, st)
                     : k7_launch<112, false, 2>(tm, tmt, p, grid, st);
        else
            rc = sym ? k7_launch<128, true, 1>(tm, tmt, p, grid, st)
                     : k7_launch<112, false, 1>(tm, tmt, p, grid, st);
    } else
    switch (sym ? 128 : n) {
        case 16: rc = k7_launch<16, false>(tm, tmt, p, grid, st); break;
        case 32: rc = k7_launch<32, false>(tm, tmt, p, grid, st); break;
        case 64: rc = k7_launch<64, false>(tm, tmt, p, grid, st); break;
        case 128: rc = k7_launch<128, true>(tm, tmt, p, grid, st); break;
        default: rc = k7_launch<112, false>(tm, tmt, p, grid, st); break;
    }
    if (rc != LTB_OK) return rc;
    set_last_kernel(sym ? 71 : (quad ? 70 : 7));
    if (banded) {
        const int64_t total = n_frames * (int64_t)n_rings * n_pairs * 2;
        int64_t blocks = (total + 255) / 256;
        if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
        k7_band_reduce_kernel<<<(int)blocks, 256, 0, st>>>(part, group_off_dev, n_frames, n_rings,
                                                           n_bands, n_pairs * 2, out, ld_out,
                                                           accumulate);
        count_launch();
        LTB_CUDA_CHECK(cudaGetLastError());
    }
    return LTB_OK;
}

extern "C" int ltb200_group_masks_tc_banded(const float* tile, int64_t n_frames, int64_t sig_size,
                                            int64_t ld_tile, const int32_t* entry_px,
                                            const float* table_split,
                                            const int32_t* group_off_host,
                                            const int32_t* group_off_dev, int n_groups,
                                            int n_pairs, int n_bands, float* out, int64_t ld_out,
                                            int accumulate, int chain, void* workspace,
                                            size_t workspace_bytes, void* stream) {
    return k7_run(tile, n_frames, sig_size, ld_tile, entry_px, table_split, group_off_host,
                  group_off_dev, n_groups, n_pairs, n_bands, 1, 0, out, ld_out, accumulate, chain,
                  workspace, workspace_bytes, stream);
}

extern "C" int ltb200_group_masks_tc_sym(const float* tile, int64_t n_frames, int64_t sig_size,
                                         int64_t ld_tile, const int32_t* entry_px,
                                         const float* table_sym, const int32_t* group_off_host,
                                         const int32_t* group_off_dev, int n_groups, int n_pairs,
                                         int n_bands, float* out, int64_t ld_out, int accumulate,
                                         int chain, void* workspace, size_t workspace_bytes,
                                         void* stream) {
    return k7_run(tile, n_frames, sig_size, ld_tile, entry_px, table_sym, group_off_host,
                  group_off_dev, n_groups, n_pairs, n_bands, 1, 1, out, ld_out, accumulate, chain,
                  workspace, workspace_bytes, stream);
}

extern "C" int ltb200_group_masks_tc(const float* tile, int64_t n_frames, int64_t sig_size,
                                     int64_t ld_tile, const int32_t* entry_px,
                                     const float* table_split, const int32_t* group_off_host,
                                     const int32_t* group_off_dev, int n_groups, int n_pairs,
                                     float* out, int64_t ld_out, int accumulate, int chain,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    return k7_run(tile, n_frames, sig_size, ld_tile, entry_px, table_split, group_off_host,
                  group_off_dev, n_groups, n_pairs, 1, 0, 0, out, ld_out, accumulate, chain,
                  workspace, workspace_bytes, stream);
}
