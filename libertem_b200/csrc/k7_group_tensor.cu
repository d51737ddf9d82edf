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
constexpr int K7_STAGES = 3;
constexpr int K7_AS = 4;              // TMEM operand ring (sub-tiles)
constexpr int K7_PWARPS = 4;
constexpr int K7_CWARPS = 8;
constexpr int K7_THREADS = (K7_PWARPS + K7_CWARPS + 1) * 32;
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
};

// Item order: ring groups of `rgroup` adjacent rings outermost, frame blocks next, the rings of
// the group innermost.  CTAs that fetch consecutive items therefore gather ADJACENT rings of the
// SAME frames within microseconds of each other, so the 128-byte lines that straddle two rings
// are served by L2 instead of a second DRAM read (DRAM traffic 2.6x -> see DESIGN.md), while the
// weight-table slices in use at any time stay at rgroup x 4 MB.
__device__ __forceinline__ void k7_decode_item(const K7Params& p, int64_t item, int& g,
                                               int64_t& fb) {
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
__device__ __forceinline__ void k7_cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

template <int N>
struct K7Smem {
    static constexpr uint32_t TABLE_BYTES = 2u * N * 128u;                 // two [N x 32] slices
    static constexpr uint32_t STAGE_BYTES = K7_DATA_BYTES + TABLE_BYTES;
    static constexpr uint32_t BAR_OFF = K7_STAGES * STAGE_BYTES;
    static constexpr uint32_t TOTAL = BAR_OFF + 512 + 1024;                // + alignment slack
};

template <int N>
__global__ void __launch_bounds__(K7_THREADS, 1)
k7_group_tensor_kernel(const __grid_constant__ CUtensorMap tm_table, const K7Params p) {
    using SM = K7Smem<N>;
    constexpr int NHALF = N / 2;          // accumulator columns drained per converter warp
    constexpr int NQ = N / 4;             // real columns per half
    constexpr uint32_t IDESC = k7_idesc_tf32(N);
    static_assert(N % 16 == 0 && N >= 16 && N <= 112, "K7: N in 16..112 step 16");

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);   // [STAGES]
    uint64_t* free_bar = full_bar + K7_STAGES;                              // [STAGES]
    uint64_t* a_full = free_bar + K7_STAGES;                                // [AS]
    uint64_t* mma_done = a_full + K7_AS;                                    // [AS]
    K7Meta* meta = reinterpret_cast<K7Meta*>(mma_done + K7_AS);             // [STAGES]
    int* cur_item = reinterpret_cast<int*>(meta + K7_STAGES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cur_item + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K7_STAGES; s++) {
            mbar_init(&full_bar[s], 1 + K7_PWARPS * 32);
            mbar_init(&free_bar[s], K7_CWARPS + 1);
        }
        for (int s = 0; s < K7_AS; s++) {
            mbar_init(&a_full[s], K7_CWARPS);
            mbar_init(&mma_done[s], 1);
        }
        fence_mbar_init();
    }
    if (warp == K7_PWARPS + K7_CWARPS) {
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
        // ===== producers =====
        const int pt = threadIdx.x;                       // 0..255
        const int el = pt & 31;                           // entry inside the sub-tile
        const int sub = (pt >> 5) & 1;                    // sub-tile of the stage
        constexpr int FPT = K7_FB / (K7_PWARPS / 2);      // frames per thread (32)
        const int fq = pt >> 6;                           // rows fq*FPT .. +FPT-1
        const uint64_t pol_keep = l2_policy_evict_last();
        if (pt == 0) prefetch_tmap(&tm_table);
        // 128-byte swizzle: 16-byte chunk (el / 4) of row r lands at chunk ^ (r & 7).  The rows
        // of a thread start at a multiple of 8, so the offset pattern repeats every 8 rows:
        // eight precomputed offsets + an immediate per copy.
        uint32_t sw[8];
#pragma unroll
        for (int c = 0; c < 8; c++)
            sw[c] = (uint32_t)sub * K7_SUB_BYTES + (uint32_t)(fq * FPT + c) * 128u +
                    ((((uint32_t)(el >> 2)) ^ (uint32_t)c) << 4) + (uint32_t)(el & 3) * 4u;
        uint32_t it = 0;
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
                e0 = p.group_off[g];
                nchunks = (p.group_off[g + 1] - e0) / K7_KT;
                if (nchunks == 0) continue;   // empty ring: nothing to add
            }
            const int64_t f_first = fb * K7_FB + fq * FPT;
            const bool ragged = fb * K7_FB + K7_FB > p.n_frames;
            // the pixel index of this thread's entry is loaded one stage ahead (an L2 round
            // trip that would otherwise sit between the stage becoming free and its copies)
            int px_next = done ? 0 : p.entry_px[e0 + sub * 32 + el];
            for (int c = 0; c < nchunks; c++, it++) {
                const int stage = it % K7_STAGES;
                mbar_wait(&free_bar[stage], ((it / K7_STAGES) & 1) ^ 1);
                uint8_t* dst = smem + (size_t)stage * SM::STAGE_BYTES;
                if (done) {
                    // sentinel stage: tells the consumers to stop (all arrivals, no data)
                    if (pt == 0) {
                        meta[stage] = K7Meta{-1, 0, 0, 0};
                        mbar_arrive(&full_bar[stage]);
                    }
                    mbar_arrive(&full_bar[stage]);
                    continue;
                }
                const int ebase = e0 + c * K7_KT;
                if (pt == 0) {
                    meta[stage] = K7Meta{item, c, nchunks, 0};
                    mbar_arrive_expect_tx(&full_bar[stage], SM::TABLE_BYTES);
                    tma_load_2d(dst + K7_DATA_BYTES, &tm_table, ebase, 0, &full_bar[stage],
                                pol_keep);
                    tma_load_2d(dst + K7_DATA_BYTES + N * 128, &tm_table, ebase + 32, 0,
                                &full_bar[stage], pol_keep);
                }
                const int px = px_next;
                if (c + 1 < nchunks) px_next = p.entry_px[ebase + K7_KT + sub * 32 + el];
                const uint32_t sdst = smem_u32(dst);
                if (!ragged) {
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
                k7_cp_async_mbar_arrive_noinc(&full_bar[stage]);
            }
            if (done) break;
        }
    } else if (warp == K7_PWARPS + K7_CWARPS) {
        // ===== MMA issuer =====
        uint32_t it = 0;       // stages
        uint32_t st = 0;       // sub-tiles (TMEM slots)
        int in_chain = 0, cbuf = 0;
        for (;; it++) {
            const int stage = it % K7_STAGES;
            mbar_wait(&full_bar[stage], (it / K7_STAGES) & 1);
            const K7Meta m = meta[stage];
            if (m.item < 0) break;
            if (m.chunk == 0) {
                in_chain = 0;
                cbuf = 0;
            }
            const uint32_t tbl = smem_u32(smem + (size_t)stage * SM::STAGE_BYTES + K7_DATA_BYTES);
#pragma unroll
            for (int s = 0; s < 2; s++, st++) {
                const int as = st % K7_AS;
                mbar_wait(&a_full[as], (st / K7_AS) & 1);
                k7_fence_after();
                const uint64_t bdesc0 = k7_desc_k_sw128(tbl + (uint32_t)s * N * 128u);
                const uint32_t a0 = tmem_base + (uint32_t)(K7_A_BASE + as * 64);
                const uint32_t d0 = tmem_base + (uint32_t)(cbuf * N);
                if (k7_elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const uint64_t bdesc = bdesc0 + (uint64_t)(kk * 2);
                        k7_mma_tf32_ts(d0, a0 + kk * 8, bdesc, IDESC, (in_chain | kk) != 0 ? 1u : 0u);
                        k7_mma_tf32_ts(d0, a0 + 32 + kk * 8, bdesc, IDESC, 1u);
                    }
                    k7_commit(&mma_done[as]);
                    if (s == 1) k7_commit(&free_bar[stage]);
                }
                __syncwarp();
                if (++in_chain == chain) {
                    in_chain = 0;
                    cbuf ^= 1;
                }
            }
        }
    } else {
        // ===== converters / accumulator drain =====
        const int cw = warp - K7_PWARPS;
        const int w = cw & 3;                       // TMEM lane quarter (== warp % 4)
        const int hh = cw >> 2;                     // entries hh*16 .. +15 of a sub-tile; drains
                                                    // accumulator columns [hh*N/2, (hh+1)*N/2)
        const int row = w * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(w * 32) << 16;
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t row_off = (uint32_t)row * 128u;

        float acc[NQ];
        uint32_t it = 0, st = 0;
        int i = 0, n_sub = 0, next_chain = 0;

        auto chain_end = [&](int c) {
            const int e = (c + 1) * chain;
            return (e < n_sub ? e : n_sub) - 1;
        };
        auto drain_one = [&]() {
            const uint32_t d = tmem_base + lane_sel + (uint32_t)((next_chain & 1) * N + hh * NHALF);
            uint32_t v[NHALF / 8][8];
#pragma unroll
            for (int q = 0; q < NHALF / 8; q++) k7_ld8(d + q * 8, v[q]);
            k7_wait_ld();
#pragma unroll
            for (int q = 0; q < NHALF / 8; q++) k7_ld_fence8(v[q]);
#pragma unroll
            for (int c = 0; c < NQ; c++)
                acc[c] += __uint_as_float(v[c / 8][c % 8]) +
                          __uint_as_float(v[(NQ + c) / 8][(NQ + c) % 8]);
            next_chain++;
        };

        for (;; it++) {
            const int stage = it % K7_STAGES;
            mbar_wait(&full_bar[stage], (it / K7_STAGES) & 1);
            const K7Meta m = meta[stage];
            if (m.item < 0) break;
            if (m.chunk == 0) {
#pragma unroll
                for (int c = 0; c < NQ; c++) acc[c] = 0.f;
                i = 0;
                n_sub = 2 * m.nchunks;
                next_chain = 0;
            }
            const uint8_t* dbase = smem + (size_t)stage * SM::STAGE_BYTES + row_off;
            float4 x[2][4];
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    x[s][j] = *reinterpret_cast<const float4*>(
                        dbase + s * K7_SUB_BYTES + (((uint32_t)(hh * 4 + j) ^ swz) << 4));
#pragma unroll
            for (int s = 0; s < 2; s++, st++, i++) {
                const int as = st % K7_AS;
                // slot `as` is free once the MMAs of sub-tile st - AS have completed
                mbar_wait(&mma_done[as], ((st / K7_AS) & 1) ^ 1);
                int known = i - K7_AS;
                if (i % chain == 0 && i >= 2 * chain) {
                    // the chain that shares its accumulator with the one starting now
                    const int must = i - chain - 1;
                    if (must > known) {
                        const uint32_t stm = st - (uint32_t)(i - must);
                        mbar_wait(&mma_done[stm % K7_AS], (stm / K7_AS) & 1);
                        known = must;
                    }
                }
                k7_fence_after();
                while (next_chain * chain < n_sub && chain_end(next_chain) <= known) drain_one();
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
                const uint32_t a = tmem_base + lane_sel + (uint32_t)(K7_A_BASE + as * 64 + hh * 16);
                k7_st16(a, hi);
                k7_st16(a + 32, lo);
                k7_wait_st();
                k7_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&a_full[as]);
                    if (s == 1) mbar_arrive(&free_bar[stage]);
                }
            }
            if (m.chunk == m.nchunks - 1) {
                // item tail: the last commit covers every earlier MMA of the item
                const uint32_t stl = st - 1;
                mbar_wait(&mma_done[stl % K7_AS], (stl / K7_AS) & 1);
                k7_fence_after();
                while (next_chain * chain < n_sub) drain_one();
                k7_fence_before();
                int g;
                int64_t fb;
                k7_decode_item(p, m.item, g, fb);
                const int64_t f = fb * K7_FB + row;
                if (f < p.n_frames) {
                    float* o = p.out + f * p.ld_out + (int64_t)g * p.n_pairs * 2 + hh * NQ;
#pragma unroll
                    for (int c = 0; c < NQ; c++)
                        if (hh * NQ + c < 2 * p.n_pairs)
                            o[c] = p.accumulate ? (o[c] + acc[c]) : acc[c];
                }
            }
        }
    }

    k7_fence_before();
    __syncthreads();
    if (warp == K7_PWARPS + K7_CWARPS) {
        k7_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(K7_TMEM_COLS)
                     : "memory");
    }
}

template <int N>
static int k7_launch(const CUtensorMap& tm, const K7Params& p, int grid, cudaStream_t st) {
    auto kern = k7_group_tensor_kernel<N>;
    const size_t smem = K7Smem<N>::TOTAL;
    int dev = 0;
    LTB_CUDA_CHECK(cudaGetDevice(&dev));
    static thread_local int configured_dev = -1;
    if (configured_dev != dev) {
        LTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        configured_dev = dev;
    }
    kern<<<grid, K7_THREADS, smem, st>>>(tm, p);
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

extern "C" int ltb200_group_masks_tc(const float* tile, int64_t n_frames, int64_t sig_size,
                                     int64_t ld_tile, const int32_t* entry_px,
                                     const float* table_split, const int32_t* group_off_host,
                                     const int32_t* group_off_dev, int n_groups, int n_pairs,
                                     float* out, int64_t ld_out, int accumulate, int chain,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_size > 0 && n_groups > 0, "group_masks_tc: bad sizes");
    const int n = ltb200_group_masks_tc_columns(n_pairs);
    LTB_REQUIRE(n > 0, "group_masks_tc: 1..28 complex columns per group, got %d", n_pairs);
    if (n_frames == 0) return LTB_OK;
    LTB_REQUIRE(tile && entry_px && table_split && group_off_host && group_off_dev && out,
                "group_masks_tc: NULL pointer");
    LTB_REQUIRE(workspace != nullptr && workspace_bytes >= 256, "group_masks_tc: workspace");
    LTB_REQUIRE(ld_out >= (int64_t)n_groups * n_pairs * 2, "group_masks_tc: ld_out too small");
    LTB_REQUIRE((uintptr_t)table_split % 16 == 0, "group_masks_tc: table must be 16 B aligned");
    LTB_REQUIRE((uintptr_t)tile % 4 == 0, "group_masks_tc: tile must be 4 B aligned");
    const int64_t n_entries = group_off_host[n_groups];
    for (int g = 0; g <= n_groups; g++)
        LTB_REQUIRE(group_off_host[g] % K7_KT == 0,
                    "group_masks_tc: offsets must be multiples of %d", K7_KT);
    LTB_REQUIRE(n_entries > 0 && n_entries % 4 == 0, "group_masks_tc: empty entry list");
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tm;
    int rc = encode_tmap_2d_sw(&tm, table_split, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                               (uint64_t)n_entries, (uint64_t)n, (uint64_t)n_entries * 4, 32,
                               (uint32_t)n, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != LTB_OK) return rc;
    K7Params p;
    p.tile = tile;
    p.n_frames = n_frames;
    p.ld_tile = ld_tile;
    p.entry_px = entry_px;
    p.group_off = group_off_dev;
    p.n_groups = n_groups;
    p.n_pairs = n_pairs;
    p.out = out;
    p.ld_out = ld_out;
    p.accumulate = accumulate;
    p.n_fb = (n_frames + K7_FB - 1) / K7_FB;
    p.n_items = p.n_fb * n_groups;
    p.counter = (int*)workspace;
    p.chain = chain > 0 ? chain : 2;
    if (chain <= 0)
        if (const char* e = getenv("LTB200_K7_CHAIN"))
            if (atoi(e) > 0) p.chain = atoi(e);
    p.rgroup = 4;
    if (const char* e = getenv("LTB200_K7_RGROUP"))
        if (atoi(e) > 0) p.rgroup = atoi(e);
    if (p.rgroup > n_groups) p.rgroup = n_groups;
    LTB_REQUIRE(p.n_items < (1ll << 31), "group_masks_tc: too many work items");
    LTB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 4, st));
    if (!accumulate) {
        // rings without entries leave their columns untouched: define them as zero
        LTB_CUDA_CHECK(cudaMemset2DAsync(out, ld_out * sizeof(float), 0,
                                         (size_t)n_groups * n_pairs * 2 * sizeof(float), n_frames,
                                         st));
    }
    int grid = sm_count();
    if (p.n_items < grid) grid = (int)p.n_items;
    switch (n) {
        case 16: return k7_launch<16>(tm, p, grid, st);
        case 32: return k7_launch<32>(tm, p, grid, st);
        case 64: return k7_launch<64>(tm, p, grid, st);
        default: return k7_launch<112>(tm, p, grid, st);
    }
}
