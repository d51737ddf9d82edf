// common.cuh -- error plumbing and sm_100a PTX wrappers shared by the ltb200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/ltb200.h"

namespace ltb {

// ---- host-side error state (thread local) ------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void set_last_kernel(int id);

#define LTB_CUDA_CHECK(expr)                                                          \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            ltb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                           __FILE__, __LINE__);                                       \
            return LTB_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

#define LTB_REQUIRE(cond, ...)                                                        \
    do {                                                                              \
        if (!(cond)) {                                                                \
            ltb::set_error(__VA_ARGS__);                                              \
            return LTB_ERR_ARG;                                                       \
        }                                                                             \
    } while (0)

int sm_count();   // of the current device (cached)

// ---- device-side PTX wrappers ------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// 2D tiled TMA load global -> shared, completion signalled on an mbarrier (UTMALDG in SASS)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int32_t c0,
                                            int32_t c1, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        ".L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float4 lds128(const float* p) {
    return *reinterpret_cast<const float4*>(p);
}

#endif  // __CUDACC__

// ---- host: TMA descriptor encoding through the driver entry point (no -lcuda needed) -----
int encode_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, size_t elem_bytes,
                   uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0,
                   uint32_t box1);
// same with a shared-memory swizzle mode (box0 * element size must not exceed the swizzle span)
int encode_tmap_2d_sw(CUtensorMap* out, const void* base, CUtensorMapDataType dt, uint64_t dim0,
                      uint64_t dim1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1,
                      CUtensorMapSwizzle swizzle);
// same with the L2 promotion granularity (the default above is 256 B: right for streams that
// walk a row, wasteful for boxes of 128-byte rows visited in another order)
int encode_tmap_2d_sw_promo(CUtensorMap* out, const void* base, CUtensorMapDataType dt,
                            uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0,
                            uint32_t box1, CUtensorMapSwizzle swizzle,
                            CUtensorMapL2promotion promotion);

// ---- K6 (tcgen05 dense masked reduction, k6_tensor.cu) used by the K1 dispatcher -----------
bool k6_shape_ok(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile);
size_t k6_workspace(int64_t n_frames, int64_t sig_size, int n_masks);
int k6_run(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
           const float* masks, int n_masks, int64_t ld_masks, float* out, int64_t ld_out,
           int accumulate, int chain, void* workspace, cudaStream_t st);
// uint16 tiles (<= 16 columns) with the frame sum (SumUDF) fused into the same pass
bool k6_u16_shape_ok(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
                     int n_masks);
size_t k6_u16_workspace(int64_t n_frames, int64_t sig_size, int n_masks, int with_sig_sum);
int k6_run_u16(const void* tile, int64_t n_frames, int64_t sig_size, int64_t ld_tile,
               const float* masks, int n_masks, int64_t ld_masks, float* out, int64_t ld_out,
               int accumulate, int chain, float* sig_sum, void* workspace, cudaStream_t st);

}  // namespace ltb
