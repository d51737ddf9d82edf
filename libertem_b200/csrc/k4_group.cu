// k4_group.cu -- K4: group-sparse masked reduction (RadialFourierAnalysis hot path).
//
// Replaces ApplyMasksUDF over radial_mask_factory masks (reference
// analysis/radialfourier.py:106-146: n_bins*(max_order+1) complex64 masks
// ring_b(r)*exp(i*o*phi), applied through the CSR rmatmul, udf/masks.py:68-69).
//
// Structure exploited: all (max_order+1) masks of ring b share ONE support (the ring's pixels),
// so per ring the contraction is a dense GEMM  C_b[F x 2G] = I[:, ring pixels] . T_b  with
// G complex columns.  Work item = (64-frame block, ring):
//   * producer warpgroup: gathers I[f, px(e)] for 64 frames x 128 ring entries per stage with
//     4-byte cp.async (lanes walk the ring's ascending pixel list -> runs coalesce) and TMA-loads
//     the matching [pair][entry][2] slice of the packed table (the exact complex64 mask values
//     of the reference, pair p = (re, im) of order p); completion of both on one mbarrier.
//     Items are fetched dynamically (atomic counter), their (item, first/last chunk) metadata
//     travels with the stage, so consumers need no scheduler of their own.
//   * consumer warps = (frame group fg, pair group mg): the mask-pair FFMA2 register tile of
//     k1_pair.cuh (8 frames x 7 pairs per lane), blocked accumulation, no cross-warp reduction.
// FP32-pipe bound (about 54 FMA per pixel): see DESIGN.md for the roofline discussion.
#include "common.cuh"

namespace ltb {

constexpr int K4_FB = 64;            // frames per item
constexpr int K4_KT = 128;           // ring entries per stage
constexpr int K4_NP = 7;             // pairs per consumer lane
constexpr int K4_MG = 4;             // pair groups  -> up to 28 complex columns per ring
constexpr int K4_FR = 8;             // frames per lane
constexpr int K4_CWARPS = 8;
constexpr int K4_PWARPS = 4;
constexpr int K4_THREADS = (K4_CWARPS + K4_PWARPS) * 32;
constexpr int K4_STAGES = 3;            // 3 x 60 KiB of the 227 KiB
constexpr int K4_NPR = K4_NP * K4_MG;                 // 28 pair rows
constexpr size_t K4_DATA_BYTES = (size_t)K4_FB * K4_KT * 4;          // 32 KiB
constexpr size_t K4_MASK_BYTES = (size_t)K4_NPR * 2 * K4_KT * 4;     // 28 KiB
constexpr size_t K4_STAGE_BYTES = K4_DATA_BYTES + K4_MASK_BYTES;

struct K4Params {
    const float* tile;
    int64_t n_frames, ld_tile;
    const int32_t* entry_px;       // (n_entries_padded): pixel index of every ring entry
    const int32_t* group_off;      // (n_groups + 1): entry offsets, multiples of K4_KT
    int n_groups, n_pairs;         // rings, complex columns per ring (<= 28)
    float* out;                    // (n_frames, ld_out) floats = complex64 (n_groups*n_pairs)
    int64_t ld_out;
    int accumulate;
    int64_t n_items;
    int* counter;
    uint32_t zero;                 // 0, unknown to the compiler (stage release after the loads)
};

struct K4Meta {
    int item;      // < 0: no more work
    int first;     // first chunk of the item
    int last;      // last chunk of the item
    int pad;
};

template <int N>
__device__ __forceinline__ void k4_xreduce_half(const float (&v)[N], float (&r)[N / 2],
                                                bool upper, int lane_xor) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        float mine = upper ? v[i + N / 2] : v[i];
        float theirs = upper ? v[i] : v[i + N / 2];
        r[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, lane_xor);
    }
}

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc)
                 : "memory");
}

__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(K4_THREADS, 1)
k4_group_kernel(const __grid_constant__ CUtensorMap tm_table, const K4Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + K4_STAGES * K4_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + K4_STAGES;
    K4Meta* meta = reinterpret_cast<K4Meta*>(empty_bar + K4_STAGES);
    int* cur_item = reinterpret_cast<int*>(meta + K4_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K4_STAGES; s++) {
            mbar_init(&full_bar[s], 1 + K4_PWARPS * 32);
            mbar_init(&empty_bar[s], K4_CWARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp < K4_PWARPS) {
        // ===== producers: dynamic item fetch, cp.async gather of the frames, TMA of the table
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int pt = threadIdx.x;                       // 0..127 = entry within the stage
        const uint64_t pol_keep = l2_policy_evict_last();
        if (pt == 0) prefetch_tmap(&tm_table);
        uint32_t it = 0;
        while (true) {
            if (pt == 0) *cur_item = atomicAdd(p.counter, 1);
            named_bar_sync(2, K4_PWARPS * 32);
            const int item = *cur_item;
            named_bar_sync(2, K4_PWARPS * 32);
            const bool done = item >= p.n_items;
            int64_t fb = 0;
            int e0 = 0, nchunks = 1;
            if (!done) {
                fb = item / p.n_groups;
                const int g = item % p.n_groups;
                e0 = p.group_off[g];
                nchunks = (p.group_off[g + 1] - e0) / K4_KT;
                if (nchunks == 0) continue;   // empty ring: nothing to add (out rows stay as is)
            }
            for (int c = 0; c < nchunks; c++, it++) {
                const int stage = it % K4_STAGES;
                mbar_wait(&empty_bar[stage], ((it / K4_STAGES) & 1) ^ 1);
                uint8_t* dst = smem + (size_t)stage * K4_STAGE_BYTES;
                if (done) {
                    // sentinel stage: tells the consumers to stop (129 arrivals, no data)
                    if (pt == 0) {
                        meta[stage] = K4Meta{-1, 0, 0, 0};
                        mbar_arrive(&full_bar[stage]);
                    }
                    mbar_arrive(&full_bar[stage]);
                    continue;
                }
                const int e = e0 + c * K4_KT + pt;
                if (pt == 0) {
                    meta[stage] = K4Meta{item, c == 0, c == nchunks - 1, 0};
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)K4_MASK_BYTES);
                    tma_load_2d(dst + K4_DATA_BYTES, &tm_table, 2 * (e0 + c * K4_KT), 0,
                                &full_bar[stage], pol_keep);
                }
                const int px = p.entry_px[e];
                const float* src = p.tile + px;
                float* drow = reinterpret_cast<float*>(dst) + pt;
#pragma unroll 8
                for (int f = 0; f < K4_FB; f++) {
                    int64_t fr = fb * K4_FB + f;
                    if (fr >= p.n_frames) fr = p.n_frames - 1;
                    cp_async_4(drow + f * K4_KT, src + fr * p.ld_tile);
                }
                // arrive when this thread's copies have landed (128 of the 129 arrivals; the
                // 129th is thread 0's expect_tx arrive above)
                cp_async_mbar_arrive_noinc(&full_bar[stage]);
            }
            if (done) break;
        }
    } else {
        // ===== consumers =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        constexpr int NP = K4_NP, FR = K4_FR;
        constexpr int NV = 2 * NP;
        const int cw = warp - K4_PWARPS;
        const int fg = cw / K4_MG;           // 0..1: 32-frame group
        const int mg = cw % K4_MG;           // pair group
        const int fl = lane >> 3;
        const int q = lane & 7;
        const int row_base = fg * 32 + fl;

        float2 acc[FR][NP];
        float tot[NV];
        int since_flush = 0;
        uint32_t it = 0;

        auto flush = [&]() {
            float v[FR * 2 * NP];
#pragma unroll
            for (int j = 0; j < FR; j++)
#pragma unroll
                for (int pp = 0; pp < NP; pp++) {
                    v[j * 2 * NP + 2 * pp] = acc[j][pp].x;
                    v[j * 2 * NP + 2 * pp + 1] = acc[j][pp].y;
                    acc[j][pp] = make_float2(0.f, 0.f);
                }
            float r1[FR * NP], r2[FR * NP / 2], r3[NV];
            k4_xreduce_half<FR * 2 * NP>(v, r1, (q & 4) != 0, 4);
            k4_xreduce_half<FR * NP>(r1, r2, (q & 2) != 0, 2);
            k4_xreduce_half<FR * NP / 2>(r2, r3, (q & 1) != 0, 1);
#pragma unroll
            for (int i = 0; i < NV; i++) tot[i] += r3[i];
        };

        for (;; it++) {
            const int stage = it % K4_STAGES;
            mbar_wait(&full_bar[stage], (it / K4_STAGES) & 1);
            const K4Meta m = meta[stage];
            if (m.item < 0) break;
            if (m.first) {
#pragma unroll
                for (int i = 0; i < NV; i++) tot[i] = 0.f;
#pragma unroll
                for (int j = 0; j < FR; j++)
#pragma unroll
                    for (int pp = 0; pp < NP; pp++) acc[j][pp] = make_float2(0.f, 0.f);
                since_flush = 0;
            }
            const float* d = reinterpret_cast<const float*>(smem + (size_t)stage * K4_STAGE_BYTES);
            const float* mk = d + K4_FB * K4_KT + (size_t)mg * NP * 256;
#pragma unroll 1
            for (int s = 0; s < K4_KT / 32; s++) {
                const int kk = s * 32 + q * 4;
                float4 dv[FR];
#pragma unroll
                for (int j = 0; j < FR; j++) dv[j] = lds128(d + (row_base + j * 4) * K4_KT + kk);
#pragma unroll
                for (int pp = 0; pp < NP; pp++) {
                    const float4 m01 = lds128(mk + pp * 256 + 2 * s * 32 + 4 * q);
                    const float4 m23 = lds128(mk + pp * 256 + 2 * s * 32 + 32 + 4 * q);
                    const float2 ma = make_float2(m01.x, m01.y), mb = make_float2(m01.z, m01.w);
                    const float2 mc = make_float2(m23.x, m23.y), md = make_float2(m23.z, m23.w);
#pragma unroll
                    for (int j = 0; j < FR; j++) {
                        float2 a = acc[j][pp];
                        a = __ffma2_rn(make_float2(dv[j].x, dv[j].x), ma, a);
                        a = __ffma2_rn(make_float2(dv[j].y, dv[j].y), mb, a);
                        a = __ffma2_rn(make_float2(dv[j].z, dv[j].z), mc, a);
                        a = __ffma2_rn(make_float2(dv[j].w, dv[j].w), md, a);
                        acc[j][pp] = a;
                    }
                }
            }
            // release after the loads have returned (see k1_dense_tma_kernel)
            const uint32_t dep = __float_as_uint(acc[FR - 1][NP - 1].y) & p.zero;
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage] + dep);
            since_flush++;
            if (since_flush == 16 || m.last) {      // chains of <= 256 terms
                flush();
                since_flush = 0;
            }
            if (m.last) {
                const int64_t fb = m.item / p.n_groups;
                const int g = m.item % p.n_groups;
                const int64_t f = fb * K4_FB + fg * 32 + q * 4 + fl;
                if (f < p.n_frames) {
                    float* o = p.out + f * p.ld_out + ((int64_t)g * p.n_pairs + mg * NP) * 2;
#pragma unroll
                    for (int i = 0; i < NV; i++) {
                        if (mg * NP + i / 2 < p.n_pairs)
                            o[i] = p.accumulate ? (o[i] + tot[i]) : tot[i];
                    }
                }
            }
        }
    }
}

}  // namespace ltb

using namespace ltb;

extern "C" size_t ltb200_group_masks_workspace(void) { return 256; }

extern "C" int ltb200_group_masks(const void* tile, int tile_dtype, int64_t n_frames,
                                  int64_t sig_size, int64_t ld_tile, const int32_t* entry_px,
                                  const float* table_packed, const int32_t* group_off_host,
                                  const int32_t* group_off_dev, int n_groups, int n_pairs,
                                  float* out, int64_t ld_out, int accumulate, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    LTB_REQUIRE(tile_dtype == LTB_F32, "group_masks: only float32 tiles are supported");
    LTB_REQUIRE(n_frames >= 0 && sig_size > 0 && n_groups > 0, "group_masks: bad sizes");
    LTB_REQUIRE(n_pairs >= 1 && n_pairs <= K4_NPR, "group_masks: 1..%d columns per group",
                K4_NPR);
    if (n_frames == 0) return LTB_OK;
    LTB_REQUIRE(tile && entry_px && table_packed && group_off_host && group_off_dev && out,
                "group_masks: NULL pointer");
    LTB_REQUIRE(workspace != nullptr && workspace_bytes >= 256, "group_masks: workspace");
    LTB_REQUIRE(ld_out >= (int64_t)n_groups * n_pairs * 2, "group_masks: ld_out too small");
    LTB_REQUIRE((uintptr_t)table_packed % 16 == 0, "group_masks: table must be 16 B aligned");
    const int64_t n_entries = group_off_host[n_groups];
    for (int g = 0; g <= n_groups; g++)
        LTB_REQUIRE(group_off_host[g] % K4_KT == 0, "group_masks: offsets must be multiples of %d",
                    K4_KT);
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tm;
    int rc = encode_tmap_2d(&tm, table_packed, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                            (uint64_t)n_entries * 2, (uint64_t)K4_NPR, (uint64_t)n_entries * 8,
                            256, K4_NPR);
    if (rc != LTB_OK) return rc;
    K4Params p;
    p.zero = 0u;
    p.tile = (const float*)tile;
    p.n_frames = n_frames;
    p.ld_tile = ld_tile;
    p.entry_px = entry_px;
    p.group_off = group_off_dev;
    p.n_groups = n_groups;
    p.n_pairs = n_pairs;
    p.out = out;
    p.ld_out = ld_out;
    p.accumulate = accumulate;
    p.n_items = ((n_frames + K4_FB - 1) / K4_FB) * n_groups;
    p.counter = (int*)workspace;
    LTB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 4, st));
    if (!accumulate) {
        // rings without entries leave their columns untouched: define them as zero
        LTB_CUDA_CHECK(cudaMemset2DAsync(out, ld_out * sizeof(float), 0,
                                         (size_t)n_groups * n_pairs * 2 * sizeof(float), n_frames,
                                         st));
    }
    const size_t smem = K4_STAGES * K4_STAGE_BYTES + 2 * K4_STAGES * sizeof(uint64_t) +
                        K4_STAGES * sizeof(K4Meta) + 16;
    int dev = 0;
    LTB_CUDA_CHECK(cudaGetDevice(&dev));
    static thread_local int configured_dev = -1;
    if (configured_dev != dev) {
        LTB_CUDA_CHECK(cudaFuncSetAttribute(k4_group_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        configured_dev = dev;
    }
    int grid = sm_count();
    if (p.n_items < grid) grid = (int)p.n_items;
    k4_group_kernel<<<grid, K4_THREADS, smem, st>>>(tm, p);
    count_launch();
    set_last_kernel(4);
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}
