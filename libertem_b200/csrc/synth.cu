// synth.cu -- counter-based synthetic 4D-STEM data on the device; bit-identical twin of
// oracle/synth.py (murmur3 finalizer over the 64-bit flat element index and a 32-bit seed).
#include "common.cuh"

namespace ltb {

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

__device__ __forceinline__ uint32_t hash_u32(uint64_t idx, uint32_t seed) {
    uint32_t lo = (uint32_t)(idx & 0xFFFFFFFFull), hi = (uint32_t)(idx >> 32);
    uint32_t h = fmix32(hi ^ seed);
    return fmix32(lo ^ h ^ 0x9E3779B9u);
}

// Poisson(3) inverse-CDF thresholds scaled to 2^32; same literals as oracle/synth.py
__constant__ uint32_t c_poisson3[20] = {
    213833830u,  855335321u,  1817587558u, 2779839795u, 3501528972u, 3934542479u, 4151049232u,
    4243837841u, 4278633569u, 4290232145u, 4293711718u, 4294660692u, 4294897936u, 4294952684u,
    4294964416u, 4294966763u, 4294967203u, 4294967280u, 4294967293u, 4294967295u};

__global__ void synth_f32_kernel(float* dst, uint64_t start, uint64_t count, uint32_t seed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride)
        dst[i] = (float)(hash_u32(start + i, seed) >> 8) * 5.9604644775390625e-08f;  // 2^-24
}

__global__ void synth_u16_kernel(uint16_t* dst, uint64_t start, uint64_t count, uint32_t seed) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) {
        uint32_t h = hash_u32(start + i, seed);
        // searchsorted(thresholds, h, side='right') == number of thresholds <= h
        int c = 0;
#pragma unroll
        for (int t = 0; t < 20; t++) c += (c_poisson3[t] <= h) ? 1 : 0;
        dst[i] = (uint16_t)c;
    }
}

}  // namespace ltb

extern "C" int ltb200_synth_fill(void* dst, int dtype, int64_t start, int64_t count,
                                 uint32_t seed, void* stream) {
    LTB_REQUIRE(dst != nullptr || count == 0, "synth_fill: dst is NULL");
    LTB_REQUIRE(start >= 0 && count >= 0, "synth_fill: negative start/count");
    if (count == 0) return LTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = ltb::sm_count() * 8;
    if (dtype == LTB_F32) {
        ltb::synth_f32_kernel<<<blocks, 256, 0, st>>>((float*)dst, (uint64_t)start,
                                                      (uint64_t)count, seed);
    } else if (dtype == LTB_U16) {
        ltb::synth_u16_kernel<<<blocks, 256, 0, st>>>((uint16_t*)dst, (uint64_t)start,
                                                      (uint64_t)count, seed);
    } else {
        ltb::set_error("synth_fill: unsupported dtype %d", dtype);
        return LTB_ERR_UNSUPPORTED;
    }
    ltb::count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}
