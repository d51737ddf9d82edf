// k10_walk.cu -- K10: group-sparse masked reduction on the tensor cores, DENSE-WALK form
//                 (RadialFourierAnalysis hot path; successor of K7's gathered form)
//
// Replaces ApplyMasksUDF over radial_mask_factory masks (reference
// analysis/radialfourier.py:106-146: n_bins*(max_order+1) complex64 masks
// ring_b(r)*exp(i*o*phi), applied through the CSR rmatmul, udf/masks.py:68-69).
//
// K7 gathers every ring's pixels with 16-byte cp.async copies and is bound by the rate of
// those copies (1.1 per clock per SM).  Here every pixel of a frame crosses the L2 -> SM fabric
// once, inside a dense TMA box [128 frames x 32 px]; the boxes are visited sorted by the first
// group (ring) they touch, so only a window of 4 groups has an open accumulator at any time,
// and the separation into groups is done by the weights: an OP is (slice of 8 consecutive
// pixels, group touching it) with an [8 x 2G] weight block that is zero outside the group.
// Everything order-dependent is decided on the host (libertem_b200/walk_plan.py): the kernel
// interprets static lists ("microcode") that are the same for every block of 128 frames.  The
// groups of even and odd id are two independent pipelines (MMA issuer warp, accumulator
// buffers, drain warps, weight-table stream) that share only the boxes and their A stages.
//
// 24 warps (22 with a role):
//   * warps 0..7   converters: thread <-> frame row (TMEM lane); the slices of a box alternate
//     between the two sets of 4 warps: two LDS.128 of the row's 32 bytes, hi/lo split (hi = top
//     19 bits, lo = x - hi rounded to TF32), tcgen05.st of both into the box's A-operand stage
//     (two stages of 4 slices x (hi 8 | lo 8) columns);
//   * warps 8..15  accumulator drain: chain totals (TMEM) -> float32 registers; groups of even
//     / odd id go to warps 8..11 / 12..15, each thread keeps two groups (window of 4); the last
//     chain of a group writes the result row;
//   * warp 20      box producer (TMA, frame stream, evict_first) + work-item fetch;
//   * warp 21      weight-table producer: one contiguous 14 KiB bulk copy per 4 ops of an
//     issuer (the tables are stored as the byte image of the swizzled stage), four rings of 2;
//   * warps 16..19 MMA issuers (groups g % 4; two per pipeline): per op three tcgen05.mma.kind::tf32 of
//     M 128, N 64, K 8 into the op's accumulator buffer (x_hi.m_hi + x_hi.m_lo + x_lo.m_hi;
//     the table rows are [hi(0..55) | lo(0..55)], the lo product reads rows 56..119).
// TMEM: 5 accumulator buffers of 64 columns (one pool; chains of <= 8 ops, the float32
// accumulate of the tensor core truncates -- see k6_tensor.cu) + 3 A stages of 4 x (hi 8 | lo 8)
// columns: the hand-over of an A stage (MMAs complete -> converters store -> MMA issuers)
// takes about as long as the tensor pipe needs for two boxes, so with two stages the pipe sat
// idle half the time.
#include "common.cuh"
#include <cstdlib>

namespace ltb {

constexpr int K10_FB = 128;                      // frames per item (TMEM lanes)
constexpr uint32_t K10_BOX_BYTES = K10_FB * 128;  // [128 frames x 32 px] float32
constexpr int K10_DSTAGES = 6;
constexpr int K10_HR = 56;                       // weight rows per half
constexpr uint32_t K10_TAB_BYTES = 2 * K10_HR * 128;        // 14 KiB copied per stage
constexpr uint32_t K10_TAB_STRIDE = K10_TAB_BYTES + 1024;   // + 8 zero rows (rows 112..119)
constexpr int K10_TSTAGES = 2;                    // per issuer (4 rings)
constexpr int K10_AS = 3;                        // A-operand stages: 4 slices x (hi 8 | lo 8)
constexpr int K10_NBUF = 5;                      // accumulator buffers: one pool (4 live + 1 in drain)
constexpr int K10_ACC_COLS = 64;
constexpr int K10_A_BASE = K10_NBUF * K10_ACC_COLS;   // 320
constexpr int K10_QLEN = 4;
constexpr int K10_MAXSEG = 8;
constexpr int K10_THREADS = 768;                  // 24 warps (22 with a role)
constexpr int K10_TMEM_COLS = 512;

constexpr uint32_t K10_OP_FIRST = 1u << 3, K10_OP_COMMIT = 1u << 4,
                   K10_OP_NEW = 1u << 5, K10_OP_END = 1u << 6,   // first / last word in a box
                   K10_OP_NOMMA = 1u << 7;                       // marker / padding
constexpr int K10_OP_SLICE_SHIFT = 8;            // bits 8-9: slice of the box
constexpr int K10_OP_PARITY_SHIFT = 10;          // FIRST: mbarrier parity of the wait for the drain
constexpr int K10_OP_ASTAGE_SHIFT = 11;          // bits 11-12: A stage of the box
constexpr int K10_OP_APARITY_SHIFT = 13;         // mbarrier parity of the A stage
constexpr uint32_t K10_EV_SLOT = 1u << 3, K10_EV_LAST = 1u << 4;
constexpr int K10_EV_PARITY_SHIFT = 5;
constexpr int K10_EV_NEXT_SHIFT = 6;             // bits 6-7: issuer (g % 4) that uses the buffer next

struct K10Params {
    const uint32_t* boxes;         // per visit: first pixel | slice mask (low 4 bits)
    const uint32_t* ops[4];        // per MMA issuer (g % 4): its words (ops / box markers)
    const uint32_t* events[2];     // per pipeline: one word per chain
    const float* table[4];         // per issuer: stage images of K10_TAB_BYTES (4 ops each)
    int visit_off[K10_MAXSEG + 1]; // multiples of 6 visits per segment
    int op_off[4][K10_MAXSEG + 1]; // multiples of 4 words
    int tab_off[4][K10_MAXSEG + 1];   // table stages
    int ev_off[2][K10_MAXSEG + 1];
    int n_seg;
    int n_cols;                    // 2 * n_pairs real columns per group
    float* out;
    int64_t ld_out;
    int64_t n_frames, n_fb, n_items;
    int* counter;
    int atomic_out;                // several segments: red.add into a zeroed / staged result
    int fgroup;                    // frame blocks per scheduling group
    int tab_policy;                // L2 policy of the table stream: 0 evict_last, 1 normal, 2 evict_first
    uint32_t zero;                 // 0, unknown to the compiler (see the converters)
};

struct K10QItem {
    int item;                      // < 0: no more work
    int seg;
    int fb;
    int pad;
};

__device__ __forceinline__ void k10_decode(const K10Params& p, int item, int& seg, int& fb) {
    // `fgroup` frame blocks outermost, segments next, the frame blocks of the group innermost:
    // all CTAs walk the same part of the weight table at the same time (served by L2)
    const int per = p.fgroup * p.n_seg;
    const int j = item / per;
    const int rem = item - j * per;
    int f_here = (int)p.n_fb - j * p.fgroup;
    if (f_here > p.fgroup) f_here = p.fgroup;
    seg = rem / f_here;
    fb = j * p.fgroup + (rem - seg * f_here);
}

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void k10_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void k10_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void k10_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void k10_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with the descriptor given as its two 32-bit halves (the low half is a running value)
__device__ __forceinline__ void k10_mma2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo,
                                         uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void k10_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                     taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
                 "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void k10_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
        "%12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void k10_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7])
        : "r"(taddr)
        : "memory");
}
template <int NR>
__device__ __forceinline__ void k10_pin(uint32_t (&r)[NR]) {
#pragma unroll
    for (int i = 0; i < NR; i++) asm volatile("" : "+r"(r[i])::"memory");
}
__device__ __forceinline__ void k10_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void k10_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool k10_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// K-major operand, 128-byte swizzle, 8-row atoms 1024 B apart (canonical UMMA layout)
__device__ __forceinline__ uint64_t k10_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t k10_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void k10_bulk_load(void* dst, const void* src, uint32_t bytes,
                                              uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

struct K10Smem {
    static constexpr uint32_t TABLE_OFF = K10_DSTAGES * K10_BOX_BYTES;
    static constexpr uint32_t BAR_OFF = TABLE_OFF + 4 * K10_TSTAGES * K10_TAB_STRIDE;
    static constexpr uint32_t TOTAL = BAR_OFF + 1024 + 1024;            // + alignment slack
};

__global__ void __launch_bounds__(K10_THREADS, 1)
k10_walk_kernel(const __grid_constant__ CUtensorMap tm_tile, const K10Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

    uint64_t* data_full = reinterpret_cast<uint64_t*>(smem + K10Smem::BAR_OFF);   // [DSTAGES]
    uint64_t* data_free = data_full + K10_DSTAGES;                                // [DSTAGES]
    uint64_t* tab_full = data_free + K10_DSTAGES;                                 // [4][TSTAGES]
    uint64_t* tab_free = tab_full + 4 * K10_TSTAGES;                              // [4][TSTAGES]
    uint64_t* a_full = tab_free + 4 * K10_TSTAGES;                                // [AS]
    uint64_t* mma_done = a_full + K10_AS;                                         // [AS]
    uint64_t* acc_full = mma_done + K10_AS;                                       // [NBUF][2]
    uint64_t* acc_free = acc_full + 2 * K10_NBUF;                                 // [NBUF][4]
    uint64_t* q_full = acc_free + 4 * K10_NBUF;                                   // [QLEN]
    uint64_t* q_free = q_full + K10_QLEN;                                         // [QLEN]
    int* meta = reinterpret_cast<int*>(q_free + K10_QLEN);                        // [DSTAGES]
    K10QItem* queue = reinterpret_cast<K10QItem*>(meta + 8);                      // [QLEN]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(queue + K10_QLEN);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    constexpr int DRAIN_WARP0 = 8, MMA_WARP = 16, DATA_WARP = 20, TABLE_WARP = 21;   // MMA: 16..19

    if (threadIdx.x == 0) {
        for (int s = 0; s < K10_DSTAGES; s++) {
            mbar_init(&data_full[s], 1);
            mbar_init(&data_free[s], 8);               // converter warps
        }
        for (int s = 0; s < 4 * K10_TSTAGES; s++) {
            mbar_init(&tab_full[s], 1);
            mbar_init(&tab_free[s], 1);
        }
        for (int s = 0; s < K10_AS; s++) {
            mbar_init(&a_full[s], 8);                  // converter warps
            mbar_init(&mma_done[s], 4);                // the four MMA issuers
        }
        for (int s = 0; s < K10_NBUF; s++) {
            // The buffers are ONE pool for both pipelines and all four issuers, which can be
            // a (short) chain apart: on a common barrier the parity of one's wait could be
            // satisfied by another's phase.  So "chain complete" has one barrier per buffer
            // and drain group, and "drained" one per buffer and WAITING issuer: the drain of a
            // chain arrives on the barrier of the issuer that uses the buffer next (static).
            mbar_init(&acc_full[2 * s], 1);
            mbar_init(&acc_full[2 * s + 1], 1);
            for (int c = 0; c < 4; c++) mbar_init(&acc_free[4 * s + c], 4);   // a drain group
        }
        for (int s = 0; s < K10_QLEN; s++) {
            mbar_init(&q_full[s], 1);
            mbar_init(&q_free[s], 13);                 // table warp, 4 MMA warps, 8 drain warps
        }
        fence_mbar_init();
    }
    // rows 112..119 of every table stage (read by the lo product, N = 64 from row 56) stay zero
    for (int i = threadIdx.x; i < 4 * K10_TSTAGES * 256; i += K10_THREADS) {
        const int s = i >> 8, w = i & 255;
        reinterpret_cast<uint32_t*>(smem + K10Smem::TABLE_OFF + (size_t)s * K10_TAB_STRIDE +
                                    K10_TAB_BYTES)[w] = 0u;
    }
    fence_proxy_async();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "n"(K10_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    k10_fence_before();
    __syncthreads();
    k10_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < DRAIN_WARP0) {
        // ===== converters =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int set = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t swz = (uint32_t)(row & 7);
        const uint32_t row_off = (uint32_t)row * 128u;
        int stage = 0;
        uint32_t dphase = 0;
        uint32_t ast = 0, aphase = 0;                   // A stage of the next box and its parity
        for (;;) {
            mbar_wait(&data_full[stage], dphase);
            const int m = meta[stage];
            if (m < 0) break;
            const uint8_t* dbase = smem + (size_t)stage * K10_BOX_BYTES + row_off;
            // the used slices of a box alternate between the two sets (at most 2 of the 4 each):
            // load, release the stage (once the loads have returned), then convert
            float4 x[2][2];
            int jj[2];
            int n_mine = 0, c = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if ((m >> j) & 1) {
                    if ((c & 1) == set) {
                        const float4 v0 = *reinterpret_cast<const float4*>(
                            dbase + (((uint32_t)(2 * j) ^ swz) << 4));
                        const float4 v1 = *reinterpret_cast<const float4*>(
                            dbase + (((uint32_t)(2 * j + 1) ^ swz) << 4));
                        if (n_mine == 0) {
                            x[0][0] = v0;
                            x[0][1] = v1;
                            jj[0] = j;
                        } else {
                            x[1][0] = v0;
                            x[1][1] = v1;
                            jj[1] = j;
                        }
                        n_mine++;
                    }
                    c++;
                }
            }
            // The stage may be handed back as soon as the loads have RETURNED.  An empty asm
            // that only names the registers does not make the hardware wait for them: the arrive
            // then overtakes the LDS and the TMA refill of the stage (fast when the next box is
            // L2-resident and boxes carry a single op) lands before the load has read it --
            // wrong rows in 5-50 % of launches on narrow-ring stacks.  The barrier ADDRESS of the
            // arrive is therefore made to depend on one word of every LDS.128 (`p.zero` is 0,
            // but only at run time; a value that is merely computed and not used is dropped by
            // ptxas): the scoreboard wait of the loads sits in front of the arrive.
            uint32_t dep = 0;
            if (n_mine > 0) dep = __float_as_uint(x[0][0].x) ^ __float_as_uint(x[0][1].x);
            if (n_mine > 1) dep ^= __float_as_uint(x[1][0].x) ^ __float_as_uint(x[1][1].x);
            dep &= p.zero;
            __syncwarp();
            if (lane == 0) mbar_arrive(&data_free[stage] + dep);
            if (++stage == K10_DSTAGES) {
                stage = 0;
                dphase ^= 1;
            }
            // the A stage of this box is free once the MMAs of the box K10_AS back have completed
            mbar_wait(&mma_done[ast], aphase ^ 1u);
            k10_fence_after();
#pragma unroll
            for (int s = 0; s < 2; s++) {
                if (s < n_mine) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float e[4] = {x[s][h].x, x[s][h].y, x[s][h].z, x[s][h].w};
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const uint32_t hb = __float_as_uint(e[t]) & 0xFFFFE000u;
                            hi[h * 4 + t] = hb;
                            lo[h * 4 + t] = __float_as_uint(e[t] - __uint_as_float(hb)) + 0x1000u;
                        }
                    }
                    const uint32_t a =
                        tmem_base + lane_sel + (uint32_t)(K10_A_BASE + ast * 64 + jj[s] * 16);
                    k10_st8(a, hi);
                    k10_st8(a + 8, lo);
                }
            }
            k10_wait_st();
            k10_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[ast]);
            if (++ast == K10_AS) {
                ast = 0;
                aphase ^= 1u;
            }
        }
    } else if (warp < MMA_WARP) {
        // ===== accumulator drain =====
        // the CTA owns 768 x 80 = 61 440 registers (what the launch allocates: see -Xptxas -v):
        // 8 x 32 x 144 here + 8 x 32 x 56 (converters) + 8 x 32 x 40 (warps 16..23)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        const int grp = (warp - DRAIN_WARP0) >> 2;
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
        float acc[2][K10_HR];
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int c = 0; c < K10_HR; c++) acc[s][c] = 0.f;
        for (uint32_t qn = 0;; qn++) {
            const int q = qn % K10_QLEN;
            mbar_wait(&q_full[q], (qn / K10_QLEN) & 1);
            const K10QItem qi = queue[q];
            __syncwarp();
            if (lane == 0) mbar_arrive(&q_free[q] + ((uint32_t)qi.item & p.zero));
            if (qi.item < 0) break;
            const int e0 = p.ev_off[grp][qi.seg], e1 = p.ev_off[grp][qi.seg + 1];
            const uint32_t* __restrict__ events = p.events[grp];
            const int64_t f = (int64_t)qi.fb * K10_FB + row;
            uint32_t w_next = e0 + lane < e1 ? events[e0 + lane] : 0u;
            for (int base = e0; base < e1; base += 32) {
                const uint32_t w = w_next;
                if (base + 32 < e1)
                    w_next = base + 32 + lane < e1 ? events[base + 32 + lane] : 0u;
                const int n = e1 - base < 32 ? e1 - base : 32;
                for (int j = 0; j < n; j++) {
                    const uint32_t ev = __shfl_sync(0xffffffffu, w, j);
                    const uint32_t buf = ev & 7u;
                    const uint32_t g = ev >> 8;
                    // (every buffer is used an even number of times per segment: static parity)
                    mbar_wait(&acc_full[2 * buf + grp], (ev >> K10_EV_PARITY_SHIFT) & 1u);
                    k10_fence_after();
                    const uint32_t d = tmem_base + lane_sel + buf * K10_ACC_COLS;
                    const bool slot1 = (ev & K10_EV_SLOT) != 0;
                    // 16 columns at a time: the register budget of these warps is the two
                    // 56-column accumulators (setmaxnreg below the CTA's launch allocation)
#pragma unroll
                    for (int c0 = 0; c0 < 48; c0 += 16) {
                        uint32_t v[16];
                        k10_ld16(d + c0, v);
                        k10_wait_ld();
                        k10_pin(v);
                        if (slot1) {
#pragma unroll
                            for (int c = 0; c < 16; c++) acc[1][c0 + c] += __uint_as_float(v[c]);
                        } else {
#pragma unroll
                            for (int c = 0; c < 16; c++) acc[0][c0 + c] += __uint_as_float(v[c]);
                        }
                    }
                    {
                        uint32_t v[8];
                        k10_ld8(d + 48, v);
                        k10_wait_ld();
                        k10_pin(v);
                        if (slot1) {
#pragma unroll
                            for (int c = 0; c < 8; c++) acc[1][48 + c] += __uint_as_float(v[c]);
                        } else {
#pragma unroll
                            for (int c = 0; c < 8; c++) acc[0][48 + c] += __uint_as_float(v[c]);
                        }
                    }
                    k10_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(&acc_free[4 * buf + ((ev >> K10_EV_NEXT_SHIFT) & 3u)]);
                    if (ev & K10_EV_LAST) {
                        // the group's sum over this segment is complete: write the row
                        float* o = p.out + f * p.ld_out + (int64_t)g * p.n_cols;
                        const bool store = f < p.n_frames;
                        if (slot1) {
#pragma unroll
                            for (int c = 0; c < K10_HR; c++) {
                                if (store && c < p.n_cols) {
                                    if (p.atomic_out) atomicAdd(o + c, acc[1][c]);
                                    else o[c] = acc[1][c];
                                }
                                acc[1][c] = 0.f;
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < K10_HR; c++) {
                                if (store && c < p.n_cols) {
                                    if (p.atomic_out) atomicAdd(o + c, acc[0][c]);
                                    else o[c] = acc[0][c];
                                }
                                acc[0][c] = 0.f;
                            }
                        }
                    }
                }
            }
        }
    } else {
        // (whole warpgroups 16..19 and 20..23: setmaxnreg is a warpgroup-wide instruction)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == DATA_WARP) {
            // ===== box producer + work-item fetch =====
            if (lane == 0) {
                prefetch_tmap(&tm_tile);
                const uint64_t pol = l2_policy_evict_first();
                int stage = 0;
                uint32_t dphase = 0;
                for (uint32_t qn = 0;; qn++) {
                    const int item = atomicAdd(p.counter, 1);
                    const bool done = item >= p.n_items;
                    int seg = 0, fb = 0;
                    if (!done) k10_decode(p, item, seg, fb);
                    const int q = qn % K10_QLEN;
                    mbar_wait(&q_free[q], ((qn / K10_QLEN) & 1) ^ 1);
                    queue[q] = K10QItem{done ? -1 : item, seg, fb, 0};
                    mbar_arrive(&q_full[q]);
                    if (done) {
                        // sentinel stage: tells the converters to stop
                        mbar_wait(&data_free[stage], dphase ^ 1);
                        meta[stage] = -1;
                        mbar_arrive(&data_full[stage]);
                        break;
                    }
                    const int v0 = p.visit_off[seg], v1 = p.visit_off[seg + 1];
                    uint32_t word_next = v0 < v1 ? p.boxes[v0] : 0u;
                    for (int v = v0; v < v1; v++) {
                        const uint32_t word = word_next;
                        if (v + 1 < v1) word_next = p.boxes[v + 1];
                        mbar_wait(&data_free[stage], dphase ^ 1);
                        meta[stage] = (int)(word & 15u);
                        mbar_arrive_expect_tx(&data_full[stage], K10_BOX_BYTES);
                        tma_load_2d(smem + (size_t)stage * K10_BOX_BYTES, &tm_tile,
                                    (int32_t)(word & ~31u), fb * K10_FB, &data_full[stage], pol);
                        if (++stage == K10_DSTAGES) {
                            stage = 0;
                            dphase ^= 1;
                        }
                    }
                }
            }
        } else if (warp == TABLE_WARP) {
            // ===== weight-table producer: one ring of K10_TSTAGES stages per issuer, served by
            // one lane that polls whichever ring has a free stage =====
            if (lane == 0) {
                uint64_t pol_keep = l2_policy_evict_last();
                if (p.tab_policy == 1)
                    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
                else if (p.tab_policy == 2)
                    pol_keep = l2_policy_evict_first();
                int ts[4] = {0, 0, 0, 0};
                uint32_t tphase[4] = {0, 0, 0, 0};
                for (uint32_t qn = 0;; qn++) {
                    const int q = qn % K10_QLEN;
                    mbar_wait(&q_full[q], (qn / K10_QLEN) & 1);
                    const K10QItem qi = queue[q];
                    mbar_arrive(&q_free[q] + ((uint32_t)qi.item & p.zero));
                    if (qi.item < 0) break;
                    int t[4], t1[4];
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        t[s] = p.tab_off[s][qi.seg];
                        t1[s] = p.tab_off[s][qi.seg + 1];
                    }
                    while (t[0] < t1[0] || t[1] < t1[1] || t[2] < t1[2] || t[3] < t1[3]) {
#pragma unroll
                        for (int s = 0; s < 4; s++) {
                            if (t[s] >= t1[s]) continue;
                            const int slot = s * K10_TSTAGES + ts[s];
                            if (!mbar_try_wait(&tab_free[slot], tphase[s] ^ 1)) continue;
                            mbar_arrive_expect_tx(&tab_full[slot], K10_TAB_BYTES);
                            k10_bulk_load(smem + K10Smem::TABLE_OFF + (size_t)slot * K10_TAB_STRIDE,
                                          p.table[s] + (size_t)t[s] * (K10_TAB_BYTES / 4),
                                          K10_TAB_BYTES, &tab_full[slot], pol_keep);
                            t[s]++;
                            if (++ts[s] == K10_TSTAGES) {
                                ts[s] = 0;
                                tphase[s] ^= 1;
                            }
                        }
                    }
                }
            }
        } else if (warp < DATA_WARP) {
            // ===== MMA issuers (one per group class g % 4 = warp - 16; two per pipeline) =====
            // One op = 96 tensor-pipe cycles, but a single thread retires an instruction only
            // every 6-8 cycles in this code and cannot issue a tcgen05.mma more often than every
            // 44.5 cycles (scripts/ubench/mma_rate_probe.cu), so the op stream is split over
            // four issuers with PRIVATE word lists and weight-table streams: nothing is walked
            // that is not the issuer's own, every word is final (buffer, A stage, all mbarrier
            // parities are static), four words per unrolled round, the words of the next round
            // are shuffled out while this one runs; the position in the table ring (stage,
            // slot of 4) lives in the elected lane's registers.
            const int me = warp - MMA_WARP;             // issuer 0..3
            const int pipe = me & 1;
            constexpr uint32_t IDESC = k10_idesc_tf32(K10_ACC_COLS);
            const uint32_t tb0 =
                smem_u32(smem + K10Smem::TABLE_OFF + (size_t)me * K10_TSTAGES * K10_TAB_STRIDE);
            const uint32_t desc_hi32 = (uint32_t)(k10_desc_k_sw128(0) >> 32);
            const uint32_t desc_lo0 = (uint32_t)k10_desc_k_sw128(tb0);
            constexpr uint32_t DESC_STAGE = K10_TAB_STRIDE >> 4, DESC_LO_HALF = (K10_HR * 128) >> 4;
            uint64_t* my_tab_full = tab_full + me * K10_TSTAGES;
            uint64_t* my_tab_free = tab_free + me * K10_TSTAGES;
            const uint32_t* __restrict__ ops = p.ops[me];
            const uint32_t a_base = tmem_base + (uint32_t)K10_A_BASE;
            // state of the elected lane (the same lane every time: the warp is converged)
            uint32_t desc = desc_lo0;                   // table stage `ts`, rows 0.., op `slot`
            int ts = 0, slot = 0;
            uint32_t tphase = 0;
            for (uint32_t qn = 0;; qn++) {
                const int q = qn % K10_QLEN;
                mbar_wait(&q_full[q], (qn / K10_QLEN) & 1);
                const K10QItem qi = queue[q];
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_free[q] + ((uint32_t)qi.item & p.zero));
                if (qi.item < 0) break;
                const int o0 = p.op_off[me][qi.seg], o1 = p.op_off[me][qi.seg + 1];
                uint32_t w_next = o0 + lane < o1 ? ops[o0 + lane] : K10_OP_NOMMA;
                for (int base = o0; base < o1; base += 32) {
                    const uint32_t w = w_next;
                    if (base + 32 < o1)
                        w_next = base + 32 + lane < o1 ? ops[base + 32 + lane] : K10_OP_NOMMA;
                    const int n = o1 - base < 32 ? o1 - base : 32;      // a multiple of 4
                    uint32_t nx[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) nx[u] = __shfl_sync(0xffffffffu, w, u);
                    for (int j = 0; j < n; j += 4) {
                        uint32_t op[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) op[u] = nx[u];
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            nx[u] = __shfl_sync(0xffffffffu, w, (j + 4 + u) & 31);
                        const bool last_round = base + j + 4 >= o1;
                        if (k10_elect_one()) {
#pragma unroll
                            for (int sub = 0; sub < 4; sub++) {
                                const uint32_t o = op[sub];
                                const uint32_t ast = (o >> K10_OP_ASTAGE_SHIFT) & 3u;
                                if (o & K10_OP_NEW) {
                                    mbar_wait(&a_full[ast], (o >> K10_OP_APARITY_SHIFT) & 1u);
                                    k10_fence_after();
                                }
                                if (!(o & K10_OP_NOMMA)) {
                                    const uint32_t buf = o & 7u;
                                    if (o & K10_OP_FIRST) {
                                        mbar_wait(&acc_free[4 * buf + (uint32_t)me],
                                                  (o >> K10_OP_PARITY_SHIFT) & 1u);
                                        k10_fence_after();
                                    }
                                    if (slot == 0) mbar_wait(&my_tab_full[ts], tphase);
                                    const uint32_t d = tmem_base + buf * K10_ACC_COLS;
                                    const uint32_t a_hi =
                                        a_base + ((o >> (K10_OP_ASTAGE_SHIFT - 6)) & 192u) +
                                        ((o >> (K10_OP_SLICE_SHIFT - 4)) & 48u);
                                    k10_mma2(d, a_hi, desc, desc_hi32, IDESC,
                                             (o & K10_OP_FIRST) ? 0u : 1u);
                                    k10_mma2(d, a_hi, desc + DESC_LO_HALF, desc_hi32, IDESC, 1u);
                                    k10_mma2(d, a_hi + 8, desc, desc_hi32, IDESC, 1u);
                                    if (o & K10_OP_COMMIT) k10_commit(&acc_full[2 * buf + pipe]);
                                    desc += 2;
                                    if (++slot == 4) {
                                        k10_commit(&my_tab_free[ts]);
                                        slot = 0;
                                        desc += DESC_STAGE - 8;
                                        if (++ts == K10_TSTAGES) {
                                            ts = 0;
                                            tphase ^= 1;
                                            desc = desc_lo0;
                                        }
                                    }
                                }
                                // this issuer's MMAs of the box (if any) release the A stage
                                if (o & K10_OP_END) k10_commit(&mma_done[ast]);
                            }
                            if (last_round && slot != 0) {
                                // the last table stage of the segment is a partial one
                                k10_commit(&my_tab_free[ts]);
                                desc += DESC_STAGE - 2u * (uint32_t)slot;
                                slot = 0;
                                if (++ts == K10_TSTAGES) {
                                    ts = 0;
                                    tphase ^= 1;
                                    desc = desc_lo0;
                                }
                            }
                        }
                        __syncwarp();
#pragma unroll
                        for (int u = 0; u < 4; u++) asm volatile("" : "+r"(nx[u])::"memory");
                    }
                }
            }
        }
    }

    k10_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        k10_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(K10_TMEM_COLS)
                     : "memory");
    }
}

// out[f, c] += part[f, c] (accumulate with several segments: the partial sums of the segments
// meet in a zeroed staging buffer first, so that the result does not depend on their order)
__global__ void k10_add_kernel(const float* __restrict__ part, int64_t n_frames, int cols,
                               float* __restrict__ out, int64_t ld_out) {
    const int64_t total = n_frames * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / cols;
        const int c = (int)(i - f * cols);
        out[f * ld_out + c] += part[i];
    }
}

}  // namespace ltb

using namespace ltb;

extern "C" size_t ltb200_group_masks_walk_workspace(int64_t n_frames, int n_groups, int n_pairs,
                                                    int accumulate) {
    size_t need = 256;
    if (accumulate)
        need += (size_t)n_frames * (size_t)n_groups * (size_t)n_pairs * 2 * sizeof(float);
    return need;
}

extern "C" int ltb200_group_masks_walk(const float* tile, int64_t n_frames, int64_t sig_size,
                                       int64_t ld_tile, const uint32_t* boxes,
                                       const uint32_t* ops0, const uint32_t* ops1,
                                       const uint32_t* ops2, const uint32_t* ops3,
                                       const uint32_t* events0, const uint32_t* events1,
                                       const float* table0, const float* table1,
                                       const float* table2, const float* table3,
                                       const int32_t* seg_off_host,
                                       int n_segments, int n_groups, int n_pairs, float* out,
                                       int64_t ld_out, int accumulate, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_size > 0 && n_groups > 0, "group_masks_walk: bad sizes");
    LTB_REQUIRE(n_pairs >= 1 && 2 * n_pairs <= K10_HR,
                "group_masks_walk: 1..%d complex columns per group, got %d", K10_HR / 2, n_pairs);
    LTB_REQUIRE(n_segments >= 1 && n_segments <= K10_MAXSEG,
                "group_masks_walk: 1..%d segments", K10_MAXSEG);
    if (n_frames == 0) return LTB_OK;
    LTB_REQUIRE(tile && boxes && ops0 && ops1 && ops2 && ops3 && events0 && events1 && table0 && table1 && table2 && table3 &&
                    seg_off_host && out,
                "group_masks_walk: NULL pointer");
    LTB_REQUIRE(sig_size % 32 == 0 && sig_size < (1ll << 31),
                "group_masks_walk: sig_size must be a multiple of 32");
    LTB_REQUIRE((uintptr_t)tile % 16 == 0 && ld_tile % 4 == 0 && ld_tile >= sig_size,
                "group_masks_walk: frame rows must be 16-byte aligned");
    LTB_REQUIRE((uintptr_t)table0 % 16 == 0 && (uintptr_t)table1 % 16 == 0 &&
                    (uintptr_t)table2 % 16 == 0 && (uintptr_t)table3 % 16 == 0,
                "group_masks_walk: tables must be 16 B aligned");
    LTB_REQUIRE(ld_out >= (int64_t)n_groups * n_pairs * 2, "group_masks_walk: ld_out too small");
    const size_t need = ltb200_group_masks_walk_workspace(n_frames, n_groups, n_pairs, accumulate);
    LTB_REQUIRE(workspace != nullptr && workspace_bytes >= need,
                "group_masks_walk: workspace of %zu B required, %zu B given", need,
                workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    K10Params p;
    p.boxes = boxes;
    p.ops[0] = ops0;
    p.ops[1] = ops1;
    p.ops[2] = ops2;
    p.ops[3] = ops3;
    p.events[0] = events0;
    p.events[1] = events1;
    p.table[0] = table0;
    p.table[1] = table1;
    p.table[2] = table2;
    p.table[3] = table3;
    const int ns1 = n_segments + 1;
    // seg_off_host rows: 0 visits, 1..4 words of issuer 0..3, 5..8 table stages, 9..10 events
    for (int s = 0; s <= K10_MAXSEG; s++) {
        const int t = s <= n_segments ? s : n_segments;
        p.visit_off[s] = seg_off_host[t];
        for (int c = 0; c < 4; c++) {
            p.op_off[c][s] = seg_off_host[(1 + c) * ns1 + t];
            p.tab_off[c][s] = seg_off_host[(5 + c) * ns1 + t];
            LTB_REQUIRE(p.op_off[c][s] % 4 == 0,
                        "group_masks_walk: word offsets must be multiples of 4");
        }
        for (int k = 0; k < 2; k++) p.ev_off[k][s] = seg_off_host[(9 + k) * ns1 + t];
        LTB_REQUIRE(p.visit_off[s] % (2 * K10_AS) == 0,
                    "group_masks_walk: visit offsets must be multiples of %d", 2 * K10_AS);
    }
    p.n_seg = n_segments;
    p.n_cols = 2 * n_pairs;
    const int cols = n_groups * n_pairs * 2;
    float* part = (float*)((uint8_t*)workspace + 256);
    p.out = accumulate ? part : out;
    p.ld_out = accumulate ? (int64_t)cols : ld_out;
    p.n_frames = n_frames;
    p.n_fb = (n_frames + K10_FB - 1) / K10_FB;
    p.n_items = p.n_fb * n_segments;
    LTB_REQUIRE(p.n_items < (1ll << 30), "group_masks_walk: too many work items");
    p.counter = (int*)workspace;
    p.atomic_out = 1;
    int grid = sm_count();
    p.fgroup = grid;
    if (const char* e = getenv("LTB200_K10_FGROUP"))
        if (atoi(e) > 0) p.fgroup = atoi(e);
    p.tab_policy = 0;
    p.zero = 0u;
    if (const char* e = getenv("LTB200_K10_TABPOL")) p.tab_policy = atoi(e);
    if (p.n_items < grid) grid = (int)p.n_items;
    CUtensorMap tm;
    // 128-byte L2 promotion: the boxes are 128-byte row pieces visited in ring order; with the
    // 256-byte default every box pulled its row neighbour out of DRAM too (1.44x the frame
    // bytes in ncu), which is gone from L2 (evict_first) before the walk gets there
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (const char* e = getenv("LTB200_K10_PROMO"))
        promo = atoi(e) == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                               : atoi(e) == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : promo;
    int rc = encode_tmap_2d_sw_promo(&tm, tile, CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                                     (uint64_t)sig_size, (uint64_t)n_frames,
                                     (uint64_t)ld_tile * 4, 32, K10_FB,
                                     CU_TENSOR_MAP_SWIZZLE_128B, promo);
    if (rc != LTB_OK) return rc;
    LTB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 4, st));
    // the segments add their partial sums into a zeroed result (<= 2 addends per element: the
    // order does not matter); groups without ops keep the zeros
    LTB_CUDA_CHECK(cudaMemset2DAsync(p.out, p.ld_out * sizeof(float), 0,
                                     (size_t)cols * sizeof(float), n_frames, st));
    static thread_local int configured_dev = -1;
    int dev = 0;
    LTB_CUDA_CHECK(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        // setmaxnreg re-deals the CTA's registers (768 threads x the compiled count) as 56 / 144 /
        // 40 per warpgroup pair: a build with fewer compiled registers would block in
        // setmaxnreg.inc for ever -- refuse it instead
        cudaFuncAttributes fa;
        LTB_CUDA_CHECK(cudaFuncGetAttributes(&fa, k10_walk_kernel));
        if (fa.numRegs * K10_THREADS < 256 * (56 + 144 + 40)) {
            set_error("group_masks_walk: kernel compiled with %d registers per thread, the "
                      "setmaxnreg layout needs 80", fa.numRegs);
            return LTB_ERR_UNSUPPORTED;
        }
        LTB_CUDA_CHECK(cudaFuncSetAttribute(k10_walk_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)K10Smem::TOTAL));
        configured_dev = dev;
    }
    k10_walk_kernel<<<grid, K10_THREADS, K10Smem::TOTAL, st>>>(tm, p);
    count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    set_last_kernel(10);
    if (accumulate) {
        const int64_t total = n_frames * (int64_t)cols;
        int64_t blocks = (total + 255) / 256;
        if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
        k10_add_kernel<<<(int)blocks, 256, 0, st>>>(part, n_frames, cols, out, ld_out);
        count_launch();
        LTB_CUDA_CHECK(cudaGetLastError());
    }
    return LTB_OK;
}
