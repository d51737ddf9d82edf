// k1_pair.cuh -- K1 "mask-pair" variant of the dense masked reduction (included by k1_dense.cu)
//
// Same pipeline as k1_dense_tma_kernel (persistent CTAs, producer warpgroup issuing 2D TMA into
// an mbarrier ring, FFMA2 consumers) with a different register tile:
//   * an accumulator pair holds TWO MASKS for one frame: acc[j][p] = (col 2p, col 2p+1), updated
//     with fma.rn.f32x2 (d, d) * (m_2p[k], m_2p+1[k]).  No even/odd doubling -> a lane can own
//     FR = 8 (or 16) frames x 12 columns in 96 registers, so every mask LDS.128 (which costs 4
//     wavefronts whether or not the quarter warps read the same bytes -- ncu, profiles/) feeds
//     twice as many FMAs and the LSU pipe stops being the bound for >= 12 columns.
//   * masks are pair-interleaved in shared memory ([pair][pixel][2]); a tiny pack kernel builds
//     that layout in the workspace per call (<= 24 x sig_size floats).
//   * templated on the input type: float32 or uint16 (converted in registers with the
//     0x4B000000 magic-number trick, exact for all 16-bit values) -- the u16 ingest variant (K3).
//   * optional fused SumUDF: when the frame block covers all 64 rows in one warp (FR = 16) and
//     sig_size*4 fits, per-pixel sums over the tile's frames are accumulated in shared memory
//     and flushed to per-CTA partial rows (deterministic), see `sig_part`.
#pragma once

namespace ltb {

template <typename TIN>
struct K1In;

template <>
struct K1In<float> {
    static constexpr CUtensorMapDataType TMAP = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    static constexpr bool INTEGER = false;
    __device__ static __forceinline__ float4 load4(const uint8_t* base, int row, int kk,
                                                    int kt) {
        return *reinterpret_cast<const float4*>(base + ((size_t)row * kt + kk) * 4);
    }
};

template <>
struct K1In<uint16_t> {
    static constexpr CUtensorMapDataType TMAP = CU_TENSOR_MAP_DATA_TYPE_UINT16;
    static constexpr bool INTEGER = true;
    static constexpr uint32_t MAGIC = 0x4B000000u;     // float bits of 2^23
    // 4 pixels as "magic words" 0x4B00vvvv: as floats they are 2^23 + v exactly, as integers
    // their payloads add without carries for up to 128 terms
    __device__ static __forceinline__ uint4 load_magic(const uint8_t* base, int row, int kk,
                                                       int kt) {
        const uint2 w = *reinterpret_cast<const uint2*>(base + ((size_t)row * kt + kk) * 2);
        uint4 m;
        m.x = __byte_perm(w.x, MAGIC, 0x7610);
        m.y = __byte_perm(w.x, MAGIC, 0x7632);
        m.z = __byte_perm(w.y, MAGIC, 0x7610);
        m.w = __byte_perm(w.y, MAGIC, 0x7632);
        return m;
    }
    __device__ static __forceinline__ float4 magic_to_float(const uint4 m) {
        float4 r;
        r.x = __uint_as_float(m.x) - 8388608.f;
        r.y = __uint_as_float(m.y) - 8388608.f;
        r.z = __uint_as_float(m.z) - 8388608.f;
        r.w = __uint_as_float(m.w) - 8388608.f;
        return r;
    }
    __device__ static __forceinline__ float4 load4(const uint8_t* base, int row, int kk,
                                                    int kt) {
        return magic_to_float(load_magic(base, row, kk, kt));
    }
};

// Packed mask layout: per pair row, per group of 32 pixels, 64 floats arranged so that the two
// LDS.128 of pixel lane q (pixels 4q..4q+3) are bank-conflict free across the 8 pixel lanes:
//   float index in group = (c/2)*32 + q*4 + (c%2)*2 + which   (pixel 4q+c, mask 2p+which)
// rows are padded to a multiple of 32 pixels (zeros), so sig_pad = ceil(sig_size/32)*32.
__global__ void k1_pack_masks_kernel(const float* __restrict__ masks, int n_masks,
                                     int64_t ld_masks, int64_t sig_size, int64_t sig_pad,
                                     int n_pairs, float* __restrict__ packed) {
    const int64_t total = (int64_t)n_pairs * sig_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int pr = (int)(i / sig_pad);
        const int64_t k = i % sig_pad;
        const int c0 = 2 * pr, c1 = 2 * pr + 1;
        float2 v = make_float2(0.f, 0.f);
        if (k < sig_size) {
            v.x = c0 < n_masks ? masks[(int64_t)c0 * ld_masks + k] : 0.f;
            v.y = c1 < n_masks ? masks[(int64_t)c1 * ld_masks + k] : 0.f;
        }
        const int r = (int)(k & 31), q = r >> 2, c = r & 3;
        const int64_t dst = (int64_t)pr * sig_pad + (k - r) + (c >> 1) * 16 + q * 2 + (c & 1);
        reinterpret_cast<float2*>(packed)[dst] = v;
    }
}

template <typename TIN, int NP, int FR, int MG>
struct K1PairCfg {
    static constexpr int KT = FR == 16 ? 256 : 128;      // pixels per pipeline chunk
    static constexpr int MH = 2 * KT / 256;              // 256-float wide mask TMA boxes per stage
    static constexpr int FG = K1_FB / (4 * FR);          // frame groups (warps along frames)
    static constexpr int KS = K1_CWARPS / (FG * MG);     // pixel split across warps
    static constexpr int KW = KT / KS;                   // pixels per warp per chunk
    static constexpr int SPC = KW / 32;                  // steps per chunk
    static constexpr int NPR = NP * MG;                  // pair rows per stage
    static constexpr int NV = (FR / 8) * 2 * NP;         // totals per lane after the transpose
    static constexpr size_t DATA_BYTES = (size_t)K1_FB * KT * sizeof(TIN);
    static constexpr size_t MASK_BYTES = (size_t)NPR * 2 * KT * 4;
    static constexpr size_t STAGE_BYTES = DATA_BYTES + MASK_BYTES;
    static constexpr int RED_ROWS = K1_CWARPS - FG * MG > 0 ? K1_CWARPS - FG * MG : 1;
    static constexpr int RED_BUFS = 2;                   // double-buffered by item parity
    static constexpr size_t FIXED_BYTES =
        2 * K1_MAX_STAGES * sizeof(uint64_t) + (size_t)RED_BUFS * RED_ROWS * NV * 32 * 4;
    static_assert(FG >= 1 && KS >= 1 && SPC >= 1, "bad K1 pair geometry");
    static_assert(FG * MG * KS == K1_CWARPS, "warps must tile the CTA");
};

template <typename TIN, int NP, int FR, int MG>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_pair_kernel(const __grid_constant__ CUtensorMap tm_data,
               const __grid_constant__ CUtensorMap tm_mask, const K1Params p) {
    using C = K1PairCfg<TIN, NP, FR, MG>;
    constexpr int FG = C::FG, KS = C::KS, KW = C::KW, SPC = C::SPC, NV = C::NV, KT = C::KT;
    constexpr int NPR = C::NPR;
    constexpr int FLUSH_EVERY = (K1_CHAIN / (4 * SPC)) > 0 ? (K1_CHAIN / (4 * SPC)) : 1;
    constexpr size_t DATA_BYTES = C::DATA_BYTES, STAGE_BYTES = C::STAGE_BYTES;

    extern __shared__ __align__(128) uint8_t smem[];
    const int S = p.n_stages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)S * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + K1_MAX_STAGES;
    float* red_base = reinterpret_cast<float*>(empty_bar + K1_MAX_STAGES);  // [2][RED_ROWS][NV][32]
    float* sig_s = red_base + C::RED_BUFS * C::RED_ROWS * NV * 32;          // [sig_size] (optional)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool do_sig = p.sig_part != nullptr;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], K1_CWARPS);
        }
        fence_mbar_init();
    }
    if (do_sig)
        for (int64_t i = threadIdx.x; i < p.sig_size; i += K1_THREADS) sig_s[i] = 0.f;
    __syncthreads();

    if (warp < K1_PWARPS) {
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
                const int nchunks = (int)((k1 - k0 + KT - 1) / KT);
                const int32_t f0 = (int32_t)(fb * K1_FB);
                for (int c = 0; c < nchunks; c++, it++) {
                    const int stage = it % S;
                    mbar_wait(&empty_bar[stage], ((it / S) & 1) ^ 1);
                    uint8_t* dst = smem + (size_t)stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)STAGE_BYTES);
                    const int32_t kc = (int32_t)(k0 + (int64_t)c * KT);
                    tma_load_2d(dst, &tm_data, kc, f0, &full_bar[stage], pol_stream);
#pragma unroll
                    for (int h = 0; h < C::MH; h++)   // mask smem layout: [h][pair row][256]
                        tma_load_2d(dst + DATA_BYTES + (size_t)h * NPR * 1024, &tm_mask,
                                    2 * kc + h * 256, 0, &full_bar[stage], pol_keep);
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int cw = warp - K1_PWARPS;
        const int fg = cw / (MG * KS);
        const int mg = (cw / KS) % MG;
        const int ks = cw % KS;
        const int fl = lane >> 3;
        const int q = lane & 7;
        const int row_base = fg * (4 * FR) + fl;         // + j*4
        const int kk_base = ks * KW + q * 4;             // + s*32

        float2 acc[FR][NP];
        float tot[NV];
        uint32_t it = 0;
        uint32_t item_parity = 0;

        for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x, item_parity ^= 1) {
            const int64_t fb = item / p.ksplit;
            const int ksi = (int)(item % p.ksplit);
            const int64_t k0 = (int64_t)ksi * p.k_per_split;
            int64_t k1 = k0 + p.k_per_split;
            if (k1 > p.sig_size) k1 = p.sig_size;
            const int nchunks = (int)((k1 - k0 + KT - 1) / KT);

#pragma unroll
            for (int i = 0; i < NV; i++) tot[i] = 0.f;
#pragma unroll
            for (int j = 0; j < FR; j++)
#pragma unroll
                for (int pp = 0; pp < NP; pp++) acc[j][pp] = make_float2(0.f, 0.f);
            int since_flush = 0;

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
                xreduce_half<FR * 2 * NP>(v, r1, (q & 4) != 0, 4);
                xreduce_half<FR * NP>(r1, r2, (q & 2) != 0, 2);
                xreduce_half<FR * NP / 2>(r2, r3, (q & 1) != 0, 1);
#pragma unroll
                for (int i = 0; i < NV; i++) tot[i] += r3[i];
            };

            for (int c = 0; c < nchunks; c++, it++) {
                const int stage = it % S;
                mbar_wait(&full_bar[stage], (it / S) & 1);
                const uint8_t* d = smem + (size_t)stage * STAGE_BYTES;
                // this warp's 2*KW mask floats per pair row live inside one 256-float box
                const int mh = (2 * ks * KW) / 256;
                const float* mk = reinterpret_cast<const float*>(d + DATA_BYTES) +
                                  (size_t)mh * NPR * 256 + (size_t)mg * NP * 256 - mh * 256;
#pragma unroll
                for (int s = 0; s < SPC; s++) {
                    const int kk = kk_base + s * 32;
                    float4 dv[FR];
                    float4 sg = make_float4(0.f, 0.f, 0.f, 0.f);
                    if constexpr (K1In<TIN>::INTEGER) {
                        // u16: the frame sum is taken on the integer pipe (exact, and it keeps
                        // ~60 FADDs per step off the FP32 pipe that feeds the FFMA2s)
                        uint4 isum = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                        for (int j = 0; j < FR; j++) {
                            const uint4 mw = K1In<TIN>::load_magic(d, row_base + j * 4, kk, KT);
                            dv[j] = K1In<TIN>::magic_to_float(mw);
                            if (FG == 1 && do_sig) {
                                isum.x += mw.x; isum.y += mw.y; isum.z += mw.z; isum.w += mw.w;
                            }
                        }
                        if (FG == 1 && do_sig) {
                            const uint32_t off = (uint32_t)FR * K1In<TIN>::MAGIC;   // mod 2^32
                            isum.x -= off; isum.y -= off; isum.z -= off; isum.w -= off;
#pragma unroll
                            for (int o = 8; o <= 16; o <<= 1) {
                                isum.x += __shfl_xor_sync(0xffffffffu, isum.x, o);
                                isum.y += __shfl_xor_sync(0xffffffffu, isum.y, o);
                                isum.z += __shfl_xor_sync(0xffffffffu, isum.z, o);
                                isum.w += __shfl_xor_sync(0xffffffffu, isum.w, o);
                            }
                            sg = make_float4(__uint2float_rn(isum.x), __uint2float_rn(isum.y),
                                             __uint2float_rn(isum.z), __uint2float_rn(isum.w));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < FR; j++)
                            dv[j] = K1In<TIN>::load4(d, row_base + j * 4, kk, KT);
                        if (FG == 1 && do_sig) {
                            sg = dv[0];
#pragma unroll
                            for (int j = 1; j < FR; j++) {
                                sg.x += dv[j].x; sg.y += dv[j].y; sg.z += dv[j].z; sg.w += dv[j].w;
                            }
#pragma unroll
                            for (int o = 8; o <= 16; o <<= 1) {
                                sg.x += __shfl_xor_sync(0xffffffffu, sg.x, o);
                                sg.y += __shfl_xor_sync(0xffffffffu, sg.y, o);
                                sg.z += __shfl_xor_sync(0xffffffffu, sg.z, o);
                                sg.w += __shfl_xor_sync(0xffffffffu, sg.w, o);
                            }
                        }
                    }
                    if (FG == 1 && do_sig) {
                        // SumUDF: this warp is the only writer of its pixel range; the fl == 0
                        // lanes accumulate the 64-frame sums in shared memory
                        if (fl == 0) {
                            const int64_t kg = k0 + (int64_t)c * KT + kk;
                            if (kg < p.sig_size) {   // sig_size % 4 == 0 on this path
                                float4* dst = reinterpret_cast<float4*>(sig_s + kg);
                                float4 cur = *dst;
                                cur.x += sg.x; cur.y += sg.y; cur.z += sg.z; cur.w += sg.w;
                                *dst = cur;
                            }
                        }
                    }
#pragma unroll
                    for (int pp = 0; pp < NP; pp++) {
                        // group base = 2*(kk - 4q); pixels (4q, 4q+1) then (4q+2, 4q+3)
                        const float4 m01 = lds128(mk + pp * 256 + 2 * (kk - 4 * q) + 4 * q);
                        const float4 m23 = lds128(mk + pp * 256 + 2 * (kk - 4 * q) + 32 + 4 * q);
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
                if (++since_flush == FLUSH_EVERY) {
                    flush();
                    since_flush = 0;
                }
            }
            if (since_flush) flush();

            // lane (fl, q) holds frames j = q*(FR/8) + jj, columns mg*2NP .. ; combine the KS
            // pixel-split warps: warps with ks > 0 publish, the ks == 0 warp adds in fixed order
            // The exchange buffer is double-buffered by item parity and the barrier is split:
            // publishing warps only ARRIVE and move on to the next item, the collecting warp
            // SYNCs.  A publisher cannot lap the collector by two items because every pipeline
            // stage needs all eight warps (safe whenever an item has more chunks than stages).
            float* red = red_base + item_parity * (C::RED_ROWS * NV * 32);
            const int red_row = (cw / KS) * (KS - 1) + (ks - 1);
            const bool split_bar = KS > 1 && nchunks >= 2 * S;
            if (ks > 0) {
#pragma unroll
                for (int i = 0; i < NV; i++) red[(red_row * NV + i) * 32 + lane] = tot[i];
                if (split_bar) {
                    __threadfence_block();
                    asm volatile("bar.arrive %0, %1;" ::"r"(1 + (int)(cw / KS)), "r"(KS * 32)
                                 : "memory");
                } else {
                    named_bar_sync(1 + (int)(cw / KS), KS * 32);
                }
            } else if (KS > 1) {
                named_bar_sync(1 + (int)(cw / KS), KS * 32);
            }
            if (ks == 0) {
#pragma unroll
                for (int jj = 0; jj < FR / 8; jj++) {
                    const int j = q * (FR / 8) + jj;
                    const int64_t f = fb * K1_FB + fg * (4 * FR) + j * 4 + fl;
                    if (f < p.n_frames) {
#pragma unroll
                        for (int cidx = 0; cidx < 2 * NP; cidx++) {
                            const int i = jj * 2 * NP + cidx;
                            float sum = tot[i];
#pragma unroll
                            for (int s2 = 1; s2 < KS; s2++)
                                sum += red[(((cw / KS) * (KS - 1) + s2 - 1) * NV + i) * 32 + lane];
                            const int col = mg * 2 * NP + cidx;
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
            }
            if (KS > 1 && !split_bar) named_bar_sync(9 + (int)(cw / KS), KS * 32);
        }
    }
    if (do_sig) {
        // per-CTA partial row of the frame sum (reduced in fixed order by colsum_final_kernel)
        __syncthreads();
        float* dst = p.sig_part + (int64_t)blockIdx.x * p.sig_size;
        for (int64_t i = threadIdx.x; i < p.sig_size; i += K1_THREADS) dst[i] = sig_s[i];
    }
}

}  // namespace ltb
