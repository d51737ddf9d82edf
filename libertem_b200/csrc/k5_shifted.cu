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
