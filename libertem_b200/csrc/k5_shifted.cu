// k5_shifted.cu -- masks shifted per frame by (dy, dx): descan-corrected virtual detectors.
//
// Replaces ApplyMasksEngine.process_frame_shifted (reference udf/masks.py:85-124), which the
// reference runs frame by frame in Python: frame pixel (y, x) meets mask pixel (y-dy, x-dx);
// pixels that do not overlap after the shift are dropped; no overlap at all gives 0.
// Every frame has its own offset, so the shared-mask-tile trick of K1 does not apply: one warp
// per frame walks the rows of the overlap rectangle with lanes along x (coalesced for frame
// and mask), 4 or 8 mask columns per pass, blocked accumulation (float32, or float64 for the
// reference's float64 dtype rule), shuffle tree.  Complex masks are interleaved (re, im) real rows.
#include "common.cuh"

namespace ltb {

// T: frame dtype, A: mask / accumulator / result dtype (float or double), NC: mask columns per
// pass over the overlap rectangle (the frame rows of a pass come from L1 / L2 after the first)
template <typename T, typename A, int NC>
__global__ void __launch_bounds__(256)
k5_shifted_kernel(const T* __restrict__ tile, int64_t n_frames, int sy, int sx, int64_t ld_tile,
                  const A* __restrict__ masks, int n_masks, int64_t ld_masks,
                  const int32_t* __restrict__ shifts, int per_frame, A* __restrict__ out,
                  int64_t ld_out, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t f = warp_global; f < n_frames; f += n_warps) {
        const int dy = shifts[per_frame ? 2 * f : 0];
        const int dx = shifts[per_frame ? 2 * f + 1 : 1];
        const int y0 = max(0, dy), y1 = min(sy, sy + dy);
        const int x0 = max(0, dx), x1 = min(sx, sx + dx);
        const T* frame = tile + f * ld_tile;
        for (int m0 = 0; m0 < n_masks; m0 += NC) {
            A tot[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) tot[c] = A(0);
            for (int y = y0; y < y1; y++) {
                A acc[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) acc[c] = A(0);
                const T* frow = frame + (int64_t)y * sx;
                const int64_t moff = (int64_t)(y - dy) * sx - dx;
                for (int x = x0 + lane; x < x1; x += 32) {
                    const A d = static_cast<A>(frow[x]);
#pragma unroll
                    for (int c = 0; c < NC; c++)
                        if (m0 + c < n_masks)
                            acc[c] = fma(d, masks[(int64_t)(m0 + c) * ld_masks + moff + x], acc[c]);
                }
#pragma unroll
                for (int c = 0; c < NC; c++) tot[c] += acc[c];
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                A v = tot[c];
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

template <typename T, typename A>
static int launch_shifted(const void* tile, int64_t F, int sy, int sx, int64_t ld,
                          const A* masks, int n_masks, int64_t ldm, const int32_t* shifts,
                          int per_frame, A* out, int64_t ldo, int accumulate, cudaStream_t st) {
    int64_t blocks = (F + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (n_masks > 4)
        k5_shifted_kernel<T, A, 8><<<(int)blocks, 256, 0, st>>>((const T*)tile, F, sy, sx, ld,
                                                                masks, n_masks, ldm, shifts,
                                                                per_frame, out, ldo, accumulate);
    else
        k5_shifted_kernel<T, A, 4><<<(int)blocks, 256, 0, st>>>((const T*)tile, F, sy, sx, ld,
                                                                masks, n_masks, ldm, shifts,
                                                                per_frame, out, ldo, accumulate);
    count_launch();
    set_last_kernel(5);
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

template <typename A>
static int shifted_dispatch(const void* tile, int tile_dtype, int64_t n_frames, int sig_y,
                            int sig_x, int64_t ld_tile, const A* masks, int n_masks,
                            int64_t ld_masks, const int32_t* shifts, int per_frame, A* out,
                            int64_t ld_out, int accumulate, void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_y > 0 && sig_x > 0 && n_masks >= 0, "masks_shifted: sizes");
    if (n_frames == 0 || n_masks == 0) return LTB_OK;
    LTB_REQUIRE(tile && masks && shifts && out, "masks_shifted: NULL pointer");
    LTB_REQUIRE(ld_tile >= (int64_t)sig_y * sig_x && ld_masks >= (int64_t)sig_y * sig_x &&
                    ld_out >= n_masks,
                "masks_shifted: leading dimension too small");
    cudaStream_t st = (cudaStream_t)stream;
#define LTB_K5_CASE(code, T)                                                                    \
    case code:                                                                                  \
        return launch_shifted<T, A>(tile, n_frames, sig_y, sig_x, ld_tile, masks, n_masks,      \
                                    ld_masks, shifts, per_frame, out, ld_out, accumulate, st);
    switch (tile_dtype) {
        LTB_K5_CASE(LTB_F32, float)
        LTB_K5_CASE(LTB_U16, uint16_t)
        LTB_K5_CASE(LTB_U8, uint8_t)
        LTB_K5_CASE(LTB_I16, int16_t)
        LTB_K5_CASE(LTB_I32, int32_t)
        LTB_K5_CASE(LTB_F64, double)
        default:
            set_error("masks_shifted: unsupported tile dtype %d", tile_dtype);
            return LTB_ERR_UNSUPPORTED;
    }
#undef LTB_K5_CASE
}

// ---- banded form (float32 masks / results): the kernel above re-reads every frame's shifted mask
// window from L2 -- M x 4 bytes per pixel, 0.09-0.20 of the HBM roofline (profiles/
// r2_k2_k5_timing.json).  Here the frames are visited in the order of their dy (`order`, sorted
// by the caller), so a chunk of consecutive frames needs almost the same mask rows for a given
// band of frame rows: a block keeps that ROW BAND of NC masks (+ the chunk's dy span) in shared
// memory and streams the chunk through it.  The masks cost one shared-memory read per pixel
// and column, the frames are read once per column group.  Work item = (chunk, row band, column
// group); band partial sums go to a workspace and are added in band order (deterministic).
// Shared memory per block is kept at <= 48 KiB so that 3-4 blocks share an SM (the first
// version held 200 KiB per block: one block of 8 warps per SM, latency-bound, slower than the
// kernel it replaced).
constexpr int K5B_THREADS = 512;
constexpr int K5B_CHUNK = 256;          // frames per work item (16 per warp)
constexpr size_t K5B_SMEM = 48 * 1024;

template <typename T, int NC>
__global__ void __launch_bounds__(K5B_THREADS)
k5_shifted_banded_kernel(const T* __restrict__ tile, int64_t n_frames, int sy, int sx,
                         int64_t ld_tile, const float* __restrict__ masks, int n_masks,
                         int64_t ld_masks, const int32_t* __restrict__ shifts, int per_frame,
                         const int32_t* __restrict__ order, int rows_per_band, int band_rows_max,
                         int n_bands, int n_cgroups,
                         float* __restrict__ part /* (n_bands, n_frames, n_masks) */) {
    extern __shared__ __align__(16) float k5_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_chunks = (n_frames + K5B_CHUNK - 1) / K5B_CHUNK;
    const int64_t n_items = n_chunks * n_bands * n_cgroups;
    const size_t cstride = (size_t)band_rows_max * sx;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        // column group innermost, bands next: the blocks that run at the same time read the
        // same frames (L2 reuse of the frame rows across column groups)
        const int cg = (int)(item % n_cgroups);
        const int band = (int)((item / n_cgroups) % n_bands);
        const int64_t chunk = item / ((int64_t)n_cgroups * n_bands);
        const int64_t i0 = chunk * K5B_CHUNK;
        const int64_t i1 = min(n_frames, i0 + K5B_CHUNK);
        const int m0 = cg * NC;
        const int y_lo = band * rows_per_band;
        const int y_hi = min(sy, y_lo + rows_per_band);
        // dy range of the chunk (frames are sorted by dy)
        const int dy_lo = per_frame ? shifts[2 * (int64_t)order[i0]] : shifts[0];
        const int dy_hi = per_frame ? shifts[2 * (int64_t)order[i1 - 1]] : shifts[0];
        const int r_lo = max(0, y_lo - dy_hi), r_hi = min(sy, y_hi - dy_lo);
        __syncthreads();                               // previous item's readers are done
        const int row_px = max(0, r_hi - r_lo) * sx;
        for (int c = 0; c < NC; c++) {
            float* dst = k5_sm + c * cstride;
            if (m0 + c < n_masks) {
                const float* src = masks + (int64_t)(m0 + c) * ld_masks + (int64_t)r_lo * sx;
                for (int i = threadIdx.x; i < row_px; i += K5B_THREADS) dst[i] = src[i];
            } else {
                for (int i = threadIdx.x; i < row_px; i += K5B_THREADS) dst[i] = 0.f;
            }
        }
        __syncthreads();
        for (int64_t i = i0 + warp; i < i1; i += K5B_THREADS / 32) {
            const int64_t f = order[i];
            const int dy = shifts[per_frame ? 2 * f : 0];
            const int dx = shifts[per_frame ? 2 * f + 1 : 1];
            const int y0 = max(y_lo, max(0, dy)), y1 = min(y_hi, min(sy, sy + dy));
            const int x0 = max(0, dx), x1 = min(sx, sx + dx);
            float tot[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) tot[c] = 0.f;
            const T* frame = tile + f * ld_tile;
            for (int y = y0; y < y1; y++) {
                float acc[NC];
#pragma unroll
                for (int c = 0; c < NC; c++) acc[c] = 0.f;
                const T* frow = frame + (int64_t)y * sx;
                const float* mrow = k5_sm + (size_t)(y - dy - r_lo) * sx - dx;
                for (int x = x0 + lane; x < x1; x += 32) {
                    const float d = static_cast<float>(frow[x]);
#pragma unroll
                    for (int c = 0; c < NC; c++) acc[c] = fmaf(d, mrow[c * cstride + x], acc[c]);
                }
#pragma unroll
                for (int c = 0; c < NC; c++) tot[c] += acc[c];
            }
#pragma unroll
            for (int c = 0; c < NC; c++) {
                float v = tot[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && m0 + c < n_masks)
                    part[((int64_t)band * n_frames + f) * n_masks + m0 + c] = v;
            }
        }
    }
}

__global__ void k5_band_reduce_kernel(const float* __restrict__ part, int n_bands, int64_t n_frames,
                                      int n_masks, float* __restrict__ out, int64_t ld_out,
                                      int accumulate) {
    const int64_t total = n_frames * n_masks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < n_bands; b++) s += part[(int64_t)b * total + i];
        float* dst = out + (i / n_masks) * ld_out + (i % n_masks);
        *dst = accumulate ? (*dst + s) : s;
    }
}

// band geometry for the largest dy span of a chunk: (mask columns per block, frame rows per band)
static bool k5_band_plan(int sy, int sx, int n_masks, int max_span, int* nc, int* rows) {
    for (int c : {8, 4}) {
        if (c == 8 && n_masks <= 4) continue;
        const int r = (int)(K5B_SMEM / ((size_t)c * sx * 4)) - max_span;
        if (r >= 2) {
            *nc = c;
            *rows = r > sy ? sy : r;
            return true;
        }
    }
    return false;
}

template <typename T>
static int launch_shifted_banded(const void* tile, int64_t F, int sy, int sx, int64_t ld,
                                 const float* masks, int n_masks, int64_t ldm,
                                 const int32_t* shifts, int per_frame, const int32_t* order,
                                 int max_span, float* out, int64_t ldo, int accumulate,
                                 float* part, cudaStream_t st) {
    int nc = 0, rows = 0;
    if (!k5_band_plan(sy, sx, n_masks, max_span, &nc, &rows)) return LTB_ERR_UNSUPPORTED;
    const int n_bands = (sy + rows - 1) / rows;
    const int n_cg = (n_masks + nc - 1) / nc;
    const int band_rows_max = rows + max_span;
    const size_t smem = (size_t)nc * band_rows_max * sx * sizeof(float);
    const int64_t n_items = ((F + K5B_CHUNK - 1) / K5B_CHUNK) * n_bands * n_cg;
    int64_t grid = (int64_t)sm_count() * 4;
    if (n_items < grid) grid = n_items;
    auto launch = [&](auto kern) -> int {
        LTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        kern<<<(int)grid, K5B_THREADS, smem, st>>>((const T*)tile, F, sy, sx, ld, masks, n_masks,
                                                   ldm, shifts, per_frame, order, rows,
                                                   band_rows_max, n_bands, n_cg, part);
        return LTB_OK;
    };
    int rc = nc == 8 ? launch(k5_shifted_banded_kernel<T, 8>)
                     : launch(k5_shifted_banded_kernel<T, 4>);
    if (rc != LTB_OK) return rc;
    const int64_t total = F * n_masks;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    k5_band_reduce_kernel<<<(int)blocks, 256, 0, st>>>(part, n_bands, F, n_masks, out, ldo,
                                                       accumulate);
    count_launch(2);
    set_last_kernel(50);
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

}  // namespace ltb

using namespace ltb;

extern "C" int ltb200_masks_shifted(const void* tile, int tile_dtype, int64_t n_frames, int sig_y,
                                    int sig_x, int64_t ld_tile, const float* masks, int n_masks,
                                    int64_t ld_masks, const int32_t* shifts, int per_frame,
                                    float* out, int64_t ld_out, int accumulate, void* stream) {
    return shifted_dispatch<float>(tile, tile_dtype, n_frames, sig_y, sig_x, ld_tile, masks,
                                   n_masks, ld_masks, shifts, per_frame, out, ld_out, accumulate,
                                   stream);
}

extern "C" int ltb200_masks_shifted_f64(const void* tile, int tile_dtype, int64_t n_frames,
                                        int sig_y, int sig_x, int64_t ld_tile,
                                        const double* masks, int n_masks, int64_t ld_masks,
                                        const int32_t* shifts, int per_frame, double* out,
                                        int64_t ld_out, int accumulate, void* stream) {
    return shifted_dispatch<double>(tile, tile_dtype, n_frames, sig_y, sig_x, ld_tile, masks,
                                    n_masks, ld_masks, shifts, per_frame, out, ld_out, accumulate,
                                    stream);
}

extern "C" size_t ltb200_masks_shifted_banded_workspace(int64_t n_frames, int sig_y, int sig_x,
                                                        int n_masks, int max_span) {
    int nc = 0, rows = 0;
    if (sig_y <= 0 || sig_x <= 0 || n_masks <= 0 || max_span < 0 ||
        !k5_band_plan(sig_y, sig_x, n_masks, max_span, &nc, &rows))
        return 0;
    const int n_bands = (sig_y + rows - 1) / rows;
    return (size_t)n_bands * (size_t)n_frames * (size_t)n_masks * sizeof(float);
}

extern "C" int ltb200_masks_shifted_banded(const void* tile, int tile_dtype, int64_t n_frames,
                                           int sig_y, int sig_x, int64_t ld_tile,
                                           const float* masks, int n_masks, int64_t ld_masks,
                                           const int32_t* shifts, int per_frame,
                                           const int32_t* order, int max_span, float* out,
                                           int64_t ld_out, int accumulate, void* workspace,
                                           size_t workspace_bytes, void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_y > 0 && sig_x > 0 && n_masks >= 0 && max_span >= 0,
                "masks_shifted_banded: sizes");
    if (n_frames == 0 || n_masks == 0) return LTB_OK;
    LTB_REQUIRE(tile && masks && shifts && order && out && workspace,
                "masks_shifted_banded: NULL pointer");
    LTB_REQUIRE(ld_tile >= (int64_t)sig_y * sig_x && ld_masks >= (int64_t)sig_y * sig_x &&
                    ld_out >= n_masks,
                "masks_shifted_banded: leading dimension too small");
    const size_t need = ltb200_masks_shifted_banded_workspace(n_frames, sig_y, sig_x, n_masks,
                                                              max_span);
    if (need == 0) {
        set_error("masks_shifted_banded: no band plan for %d x %d px, dy span %d (use "
                  "ltb200_masks_shifted)", sig_y, sig_x, max_span);
        return LTB_ERR_UNSUPPORTED;
    }
    LTB_REQUIRE(workspace_bytes >= need, "masks_shifted_banded: workspace of %zu B required", need);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)workspace;
#define LTB_K5B_CASE(code, T)                                                                   \
    case code:                                                                                  \
        return launch_shifted_banded<T>(tile, n_frames, sig_y, sig_x, ld_tile, masks, n_masks,  \
                                        ld_masks, shifts, per_frame, order, max_span, out,      \
                                        ld_out, accumulate, part, st);
    switch (tile_dtype) {
        LTB_K5B_CASE(LTB_F32, float)
        LTB_K5B_CASE(LTB_U16, uint16_t)
        LTB_K5B_CASE(LTB_U8, uint8_t)
        LTB_K5B_CASE(LTB_I16, int16_t)
        default:
            set_error("masks_shifted_banded: unsupported tile dtype %d", tile_dtype);
            return LTB_ERR_UNSUPPORTED;
    }
#undef LTB_K5B_CASE
}
