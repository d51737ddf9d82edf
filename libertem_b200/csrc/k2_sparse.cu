// k2_sparse.cu -- K2: sparse masked reduction (gather + warp reduce).
//
// Replaces ApplyMasksEngine._process_flat_spsp -> rmatmul (reference udf/masks.py:68-69,
// common/numba/__init__.py:90-184).  Masks arrive as CSC over (sig_size, n_masks): per mask the
// ascending list of (pixel index, weight).  One warp per frame: lanes stride over a mask's
// non-zeros (ring/disk masks are runs of consecutive pixels, so the gathers coalesce), fp32 FMA
// per lane, warp-shuffle tree at the end.  Integer-valued inputs give exact sums (< 2^24).
#include "common.cuh"

namespace ltb {

template <typename T>
__global__ void __launch_bounds__(256)
k2_csc_kernel(const T* __restrict__ tile, int64_t n_frames, int64_t ld_tile,
              const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
              const float* __restrict__ values, int n_masks, float* __restrict__ out,
              int64_t ld_out, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t f = warp_global; f < n_frames; f += n_warps) {
        const T* row = tile + f * ld_tile;
        for (int m = 0; m < n_masks; m++) {
            const int32_t b = indptr[m], e = indptr[m + 1];
            float acc = 0.f;
            for (int32_t i = b + lane; i < e; i += 32)
                acc = fmaf(static_cast<float>(row[indices[i]]), values[i], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                float* dst = out + f * ld_out + m;
                *dst = accumulate ? (*dst + acc) : acc;
            }
        }
    }
}

template <typename T>
static int launch_csc(const void* tile, int64_t F, int64_t ld, const int32_t* indptr,
                      const int32_t* indices, const float* values, int n_masks, float* out,
                      int64_t ldo, int accumulate, cudaStream_t st) {
    int64_t blocks = (F + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k2_csc_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)tile, F, ld, indptr, indices, values,
                                                  n_masks, out, ldo, accumulate);
    count_launch();
    set_last_kernel(20);
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

}  // namespace ltb

using namespace ltb;

extern "C" int ltb200_masks_csc(const void* tile, int tile_dtype, int64_t n_frames,
                                int64_t sig_size, int64_t ld_tile, const int32_t* indptr,
                                const int32_t* indices, const float* values, int n_masks,
                                float* out, int64_t ld_out, int accumulate, void* stream) {
    LTB_REQUIRE(n_frames >= 0 && sig_size >= 0 && n_masks >= 0, "masks_csc: negative size");
    if (n_frames == 0 || n_masks == 0) return LTB_OK;
    LTB_REQUIRE(tile && indptr && out, "masks_csc: NULL pointer");
    LTB_REQUIRE(ld_tile >= sig_size && ld_out >= n_masks, "masks_csc: leading dimension too small");
    cudaStream_t st = (cudaStream_t)stream;
    switch (tile_dtype) {
        case LTB_F32: return launch_csc<float>(tile, n_frames, ld_tile, indptr, indices, values, n_masks, out, ld_out, accumulate, st);
        case LTB_U16: return launch_csc<uint16_t>(tile, n_frames, ld_tile, indptr, indices, values, n_masks, out, ld_out, accumulate, st);
        case LTB_U8: return launch_csc<uint8_t>(tile, n_frames, ld_tile, indptr, indices, values, n_masks, out, ld_out, accumulate, st);
        case LTB_I8: return launch_csc<int8_t>(tile, n_frames, ld_tile, indptr, indices, values, n_masks, out, ld_out, accumulate, st);
        case LTB_I16: return launch_csc<int16_t>(tile, n_frames, ld_tile, indptr, indices, values, n_masks, out, ld_out, accumulate, st);
        default:
            set_error("masks_csc: unsupported tile dtype %d", tile_dtype);
            return LTB_ERR_UNSUPPORTED;
    }
}
