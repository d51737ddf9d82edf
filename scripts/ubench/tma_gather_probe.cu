// Micro-benchmark: rate of small 2D TMA boxes [128 rows x W floats] at scattered column offsets
// (the "gather by TMA" candidate for K7's ring gather) on L2-resident data.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_gather_probe tma_gather_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1),
                 "r"(smem_u32(bar)) : "memory");
}

constexpr int STAGES = 4;
constexpr int STAGE_BYTES = 32768;     // 128 rows x 64 floats per stage

__global__ void __launch_bounds__(32, 1)
probe(const __grid_constant__ CUtensorMap tm, const int* __restrict__ cols, int n_cols, int W,
      int n_stages, int n_fb) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + STAGES * STAGE_BYTES);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    const int boxes = 64 / W;                  // boxes per stage
    const uint32_t box_bytes = 128u * W * 4u;
    int issued = 0;
    auto issue = [&](int st) {
        const int s = st % STAGES;
        mbar_expect(&full[s], STAGE_BYTES);
        const int fb = (blockIdx.x + st) % n_fb;
        for (int b = 0; b < boxes; b++) {
            const int c = cols[(uint32_t)(blockIdx.x * 7919 + st * boxes + b) % n_cols];
            tma_2d(sm + s * STAGE_BYTES + b * box_bytes, &tm, c, fb * 128, &full[s]);
        }
    };
    for (; issued < STAGES && issued < n_stages; issued++) issue(issued);
    for (int st = 0; st < n_stages; st++) {
        mbar_wait(&full[st % STAGES], (st / STAGES) & 1);
        if (issued < n_stages) issue(issued++);
    }
}

int main(int argc, char** argv) {
    const int K = 65536;          // pixels per frame row
    const int F = 128 * 2;        // 2 frame blocks: 64 MiB, L2-resident
    float* d;
    cudaMalloc(&d, (size_t)F * K * 4);
    cudaMemset(d, 0, (size_t)F * K * 4);
    const int n_cols = 1 << 16;
    int* dcols;
    cudaMalloc(&dcols, n_cols * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * STAGE_BYTES + 64);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    for (int mode = 0; mode < 2; mode++)          // 0: scattered columns, 1: consecutive columns
    for (int W : {4, 8, 16, 32, 64}) {
        std::vector<int> h(n_cols);
        uint32_t x = 12345;
        for (int i = 0; i < n_cols; i++) {
            x = x * 1664525u + 1013904223u;
            h[i] = mode == 0 ? (int)((x >> 8) % (K / W)) * W : (i * W) % K;
        }
        cudaMemcpy(dcols, h.data(), n_cols * 4, cudaMemcpyHostToDevice);
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)F};
        cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        cuuint32_t box[2] = {(cuuint32_t)W, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int n_stages = 4000;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe<<<148, 32, STAGES * STAGE_BYTES + 64>>>(tm, dcols, n_cols, W, 200, 2);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        probe<<<148, 32, STAGES * STAGE_BYTES + 64>>>(tm, dcols, n_cols, W, n_stages, 2);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 148.0 * n_stages * STAGE_BYTES;
        const double rows = 148.0 * n_stages * (64 / W) * 128;
        printf("%s W=%2d (%3d B rows): %.3f ms  %.1f GB/s useful  %.2f us/stage/SM  %.3f rows/ns/SM  (%s)\n",
               mode ? "consecutive" : "scattered  ", W, W * 4, ms, bytes / ms / 1e6, ms * 1e3 / n_stages,
               rows / 148 / (ms * 1e6), cudaGetErrorString(err));
    }
    return 0;
}
