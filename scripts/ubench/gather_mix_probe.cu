// Micro-benchmark 2 for K7's ring gather (L2-resident data, 148 CTAs):
//   A. LDG.128 into registers + STS.128 (K7 lane mapping: 16 quads x 2 rows per warp instr),
//      PW producer warps, B loads in flight per thread
//   B. LDGSTS (cp.async 16 B) for some of the quads + TMA boxes [128 rows x 16 B] for the
//      others, concurrently: do the two paths add up?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_mix_probe gather_mix_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint4 ldg16(const void* src) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t dst, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

constexpr int K = 65536, STAGE_BYTES = 32768, NSTAGE = 4;
__constant__ int NFB = 2;     // frame blocks of 128 rows in the buffer (2: L2 resident)

// A: PW*32 threads; a stage = 128 rows x 16 quads = 2048 copies
template <int PW, int B>
__global__ void __launch_bounds__(PW * 32, 1)
probe_ldg(const float* __restrict__ data, const int* __restrict__ list, int n_list, int n_stages, long long* clocks) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const int pt = threadIdx.x;
    const uint32_t sbase = smem_u32(sm);
    constexpr int PER = 2048 / (PW * 32);       // copies per thread per stage
    constexpr int RSTEP = PW * 2;               // rows per step (16 quads x RSTEP rows per CTA step)
    long long t0 = clock64();
    for (int st = 0; st < n_stages; st++) {
        const uint32_t dst0 = sbase + (st % NSTAGE) * STAGE_BYTES + pt * 16;
        const int fb = (blockIdx.x * 7 + st) % NFB;
        const int li = (blockIdx.x * 977 + st * 16) % (n_list - 64);
        const int px = list[li + (pt & 15)];
        const float* src = data + (size_t)(fb * 128 + (pt >> 4)) * K + px;
#pragma unroll
        for (int j0 = 0; j0 < PER; j0 += B) {
            uint4 v[B];
#pragma unroll
            for (int j = 0; j < B; j++) v[j] = ldg16(src + (size_t)(j0 + j) * RSTEP * K);
#pragma unroll
            for (int j = 0; j < B; j++) sts16(dst0 + (j0 + j) * (PW * 512), v[j]);
        }
    }
    if (pt == 0) clocks[blockIdx.x] = clock64() - t0;
}

// B: warps 0..3 LDGSTS for quads [0, NL) of each stage, warp 4 lane 0 TMA boxes for the others
template <int LT>
__global__ void __launch_bounds__(LT + 32, 1)
probe_mix(const __grid_constant__ CUtensorMap tm, const float* __restrict__ data, const int* __restrict__ list,
          int n_list, int n_stages, int n_ldgsts_quads, long long* clocks) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + NSTAGE * STAGE_BYTES);
    const int pt = threadIdx.x;
    const uint32_t sbase = smem_u32(sm);
    if (pt == 0) {
        for (int s = 0; s < NSTAGE; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64();
    const int NL = n_ldgsts_quads;              // quads per stage through LDGSTS (0..16)
    if (pt < LT) {
        if (NL > 0) {
            const int per_row = NL;
            for (int st = 0; st < n_stages; st++) {
                const uint32_t dst0 = sbase + (st % NSTAGE) * STAGE_BYTES;
                const int fb = (blockIdx.x * 7 + st) % NFB;
                const int li = (blockIdx.x * 977 + st * 16) % (n_list - 64);
                for (int c = pt; c < 128 * per_row; c += LT) {
                    const int q = c % per_row, row = c / per_row;
                    cp16(dst0 + c * 16, data + (size_t)(fb * 128 + row) * K + list[li + q]);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 3;" ::: "memory");
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
    } else if (pt == LT && NL < 16) {
        const int nb = 16 - NL;
        int issued = 0;
        auto issue = [&](int st) {
            const int s = st % NSTAGE;
            mbar_expect(&full[s], nb * 2048);
            const int fb = (blockIdx.x * 7 + st) % NFB;
            const int li = (blockIdx.x * 977 + st * 16) % (n_list - 64);
            for (int b = 0; b < nb; b++)
                tma_2d(sbase + s * STAGE_BYTES + (NL + b) * 2048, &tm, list[li + NL + b], fb * 128, &full[s]);
        };
        for (; issued < NSTAGE && issued < n_stages; issued++) issue(issued);
        for (int st = 0; st < n_stages; st++) {
            mbar_wait(&full[st % NSTAGE], (st / NSTAGE) & 1);
            if (issued < n_stages) issue(issued++);
        }
    }
    __syncthreads();
    if (pt == 0) clocks[blockIdx.x] = clock64() - t0;
}

int main(int argc, char** argv) {
    const int nfb = argc > 1 ? atoi(argv[1]) : 2;
    const int ROWS = nfb * 128;
    cudaMemcpyToSymbol(NFB, &nfb, sizeof(int));
    printf("buffer: %d frame blocks = %.0f MiB\n", nfb, (double)ROWS * K * 4 / 1048576);
    float* d;
    cudaMalloc(&d, (size_t)ROWS * K * 4);
    cudaMemset(d, 0, (size_t)ROWS * K * 4);
    const int n_list = 1 << 16;
    std::vector<int> ring(n_list);
    uint32_t x = 777;
    int pos = 0, left = 0;
    for (int i = 0; i < n_list; i++) {
        x = x * 1664525u + 1013904223u;
        if (left == 0) { pos = (int)((x >> 8) % (K / 4 - 8)) * 4; left = 2 + (x >> 28) % 6; }
        ring[i] = pos; pos += 4; left--;
    }
    int* dring; cudaMalloc(&dring, n_list * 4);
    cudaMemcpy(dring, ring.data(), n_list * 4, cudaMemcpyHostToDevice);
    long long* dclk; cudaMalloc(&dclk, 148 * 8);
    const int n_stages = 3000;
    const int smem = NSTAGE * STAGE_BYTES + 64;
    auto report = [&](const char* name, float ms) {
        long long clk[148]; cudaMemcpy(clk, dclk, sizeof(clk), cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < 148; i++) mean += clk[i]; mean /= 148;
        const double quads = (double)n_stages * 2048;
        printf("%-46s: %.3f ms  %.2f us/stage  %.3f quads/clk/SM  %.2f quads/ns/SM (%.0f MHz)  %s\n", name, ms,
               ms * 1e3 / n_stages, quads / mean, quads / (ms * 1e6), mean / (ms * 1e3), cudaGetErrorString(cudaGetLastError()));
    };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN_LDG(PW, B) { auto k = probe_ldg<PW, B>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    for (int r = 0; r < 2; r++) { cudaEventRecord(e0); k<<<148, PW * 32, smem>>>(d, dring, n_list, n_stages, dclk); cudaEventRecord(e1); cudaDeviceSynchronize(); } \
    float ms; cudaEventElapsedTime(&ms, e0, e1); char nm[64]; snprintf(nm, 64, "LDG.128+STS.128, %d warps, %d loads in flight", PW, B); report(nm, ms); }
    RUN_LDG(4, 8) RUN_LDG(4, 16) RUN_LDG(8, 8) RUN_LDG(8, 4) RUN_LDG(16, 4) RUN_LDG(16, 2)
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)ROWS};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {4, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
#define RUN_MIX(LT, nl) { auto k = probe_mix<LT>; cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k<<<148, LT + 32, smem>>>(tm, d, dring, n_list, n_stages, nl, dclk); cudaEventRecord(e1); cudaDeviceSynchronize(); } \
        float ms; cudaEventElapsedTime(&ms, e0, e1); char nm[80]; snprintf(nm, 80, "%d LDGSTS threads: %d quads LDGSTS + %d TMA", LT, nl, 16 - nl); report(nm, ms); }
    RUN_MIX(128, 16) RUN_MIX(256, 16) RUN_MIX(512, 16) RUN_MIX(128, 8) RUN_MIX(128, 0)
    return 0;
}
