// probe.cu -- read-only HBM streaming probes: the denominator of the frame-stream roofline.
//
// The masked-reduction kernels (K6 / K8 / K1) READ the frame stream once and write almost
// nothing, while MEASURED_PEAKS.json's hbm_gbs is a copy benchmark (read + write).  These probes
// measure what the memory system delivers to a kernel of the same shape that only reads:
//   mode 0: one persistent CTA per SM, a ring of 32 KiB shared-memory stages filled by bulk
//           TMA copies (cp.async.bulk global -> shared, mbarrier complete_tx), L2 evict_first,
//           no math at all -- the ingest half of K6 / K8 with everything else removed;
//   mode 1: the same grid reading with LDG.128 (ld.global.nc.L1::no_allocate) into registers,
//           XOR-reduced so the loads cannot be elided.
// Not on the product path: bench.py reports `roofline.peak_read_only` from it.
#include "common.cuh"

namespace ltb {

constexpr int PR_STAGE_BYTES = 32 * 1024;
constexpr int PR_STAGES = 6;

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__global__ void __launch_bounds__(32, 1)
probe_bulk_kernel(const uint8_t* __restrict__ buf, int64_t n_chunks) {
    extern __shared__ __align__(128) uint8_t psm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(psm + PR_STAGES * PR_STAGE_BYTES);
    if (threadIdx.x == 0) {
        for (int s = 0; s < PR_STAGES; s++) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    const uint64_t pol = l2_policy_evict_first();
    // chunk c belongs to CTA c % gridDim.x: neighbouring CTAs stream neighbouring 32 KiB chunks
    int64_t issued = blockIdx.x, waited = 0;
    uint32_t it = 0;
    for (int s = 0; s < PR_STAGES && issued < n_chunks; s++, issued += gridDim.x) {
        mbar_arrive_expect_tx(&full[s], PR_STAGE_BYTES);
        bulk_load_1d(psm + s * PR_STAGE_BYTES, buf + issued * PR_STAGE_BYTES, PR_STAGE_BYTES,
                     &full[s], pol);
    }
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x, it++, waited++) {
        const int s = (int)(it % PR_STAGES);
        mbar_wait(&full[s], (it / PR_STAGES) & 1);
        if (issued < n_chunks) {
            // the stage was never read through the generic proxy: it can be refilled at once
            mbar_arrive_expect_tx(&full[s], PR_STAGE_BYTES);
            bulk_load_1d(psm + s * PR_STAGE_BYTES, buf + issued * PR_STAGE_BYTES,
                         PR_STAGE_BYTES, &full[s], pol);
            issued += gridDim.x;
        }
    }
}

__global__ void __launch_bounds__(512, 2)
probe_ldg_kernel(const uint4* __restrict__ buf, int64_t n_vec, uint32_t* sink) {
    uint32_t acc = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n_vec; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                         : "l"(buf + i + j * stride));
#pragma unroll
        for (int j = 0; j < 8; j++) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    for (; i < n_vec; i += stride) {
        const uint4 v = buf[i];
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x9E3779B9u) *sink = acc;     // practically never: keeps the loads alive
}

}  // namespace ltb

using namespace ltb;

extern "C" int ltb200_probe_read(const void* buf, size_t bytes, int mode, void* sink,
                                 void* stream) {
    LTB_REQUIRE(buf != nullptr && (uintptr_t)buf % 16 == 0, "probe_read: 16 B aligned buffer");
    LTB_REQUIRE(mode == 0 || mode == 1, "probe_read: mode 0 (bulk TMA) or 1 (LDG.128)");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        const int64_t n_chunks = (int64_t)(bytes / PR_STAGE_BYTES);
        LTB_REQUIRE(n_chunks > 0, "probe_read: buffer smaller than one 32 KiB stage");
        const size_t smem = PR_STAGES * PR_STAGE_BYTES + 64;
        static thread_local int configured_dev = -1;
        int dev = 0;
        LTB_CUDA_CHECK(cudaGetDevice(&dev));
        if (configured_dev != dev) {
            LTB_CUDA_CHECK(cudaFuncSetAttribute(probe_bulk_kernel,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
            configured_dev = dev;
        }
        int grid = sm_count();
        if (n_chunks < grid) grid = (int)n_chunks;
        probe_bulk_kernel<<<grid, 32, smem, st>>>((const uint8_t*)buf, n_chunks);
    } else {
        LTB_REQUIRE(sink != nullptr, "probe_read: mode 1 needs a 4-byte device sink");
        probe_ldg_kernel<<<sm_count() * 2, 512, 0, st>>>((const uint4*)buf,
                                                         (int64_t)(bytes / 16), (uint32_t*)sink);
    }
    count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}
