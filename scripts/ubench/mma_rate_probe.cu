// Micro-benchmark: issue rate of tcgen05.mma.kind::tf32 (M 128, K 8, A from TMEM, B from shared
// memory, K-major 128-byte swizzle) as a function of N -- what does one small MMA really cost?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_rate_probe mma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b),
                 "r"(idesc), "r"(acc) : "memory");
}

template <int N, int NACC, int NISS = 1>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    uint8_t* base = sm + ((1024 - (smem_u32(sm) & 1023)) & 1023);
    for (int i = threadIdx.x; i < 256 * 128 / 4; i += 128) reinterpret_cast<float*>(base)[i] = 1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NISS));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tmem_slot;
    // zero the A operand columns (448..463)
    {
        const uint32_t a = t + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + 448;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(a + 8), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // NISS issuing threads (lane 0 of warps 0 .. NISS-1), each with its own accumulators
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < NISS) {
        const int wi = threadIdx.x >> 5;
        const uint64_t b = desc_k_sw128(smem_u32(base));
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 12; u++) {
                const int acc_i = (i * 12 + u) % NACC;            // rotate accumulators
                mma(t + (wi * NACC + acc_i) * N, t + 448 + (u & 1) * 8, b + (uint64_t)((u & 3) * 2), idesc_tf32(N), 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0 && wi == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(t) : "memory");
}

template <int N, int NACC, int NISS = 1>
void run(const char* tag, long long* d_out) {
    const int iters = 20000;
    cudaFuncSetAttribute(probe<N, NACC, NISS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128 + 2048);
    for (int rep = 0; rep < 2; rep++) probe<N, NACC, NISS><<<148, 128, 256 * 128 + 2048>>>(iters, d_out);
    long long c;
    cudaError_t e = cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-28s N=%3d accumulators=%d issuers=%d: %6.1f cycles per MMA (all issuers)  (floor 128*N/256 = %d)  %s\n", tag, N, NACC, NISS,
           (double)c / (iters * 12.0 * NISS), N / 2, cudaGetErrorString(e));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 8);
    run<16, 4>("tf32 M128 K8 A=TMEM", d_out);
    run<32, 4>("tf32 M128 K8 A=TMEM", d_out);
    run<64, 1>("tf32 M128 K8 A=TMEM same D", d_out);
    run<64, 4>("tf32 M128 K8 A=TMEM", d_out);
    run<112, 2>("tf32 M128 K8 A=TMEM", d_out);
    run<128, 2>("tf32 M128 K8 A=TMEM", d_out);
    run<256, 1>("tf32 M128 K8 A=TMEM", d_out);
    run<64, 1, 2>("tf32 M128 K8 A=TMEM", d_out);
    run<64, 1, 4>("tf32 M128 K8 A=TMEM", d_out);
    run<32, 2, 4>("tf32 M128 K8 A=TMEM", d_out);
    run<112, 1, 2>("tf32 M128 K8 A=TMEM", d_out);
    return 0;
}
