// Micro-benchmark: 16-byte cp.async (LDGSTS.128) gather rate per SM as a function of the
// lane -> address mapping, on L2-resident data.  Question behind it: what bounds K7's ring
// gather -- copies per clock, distinct 128-byte lines per warp instruction, or sectors?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ldgsts_gather_probe ldgsts_gather_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp16_ca(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int K = 65536;            // floats per row
constexpr int ROWS = 256;           // 64 MiB, L2 resident
constexpr int STAGE_BYTES = 32768;
constexpr int NSTAGE = 4;

// mode: 0 = K7 mapping (16 quads x 2 rows per warp instruction, quads from `list`)
//       1 = 32 consecutive quads of one row          2 = one quad, 32 rows
//       3 = 8 consecutive quads (one 128 B line) x 4 rows
//       4 = 4 consecutive quads (64 B) x 8 rows
template <bool CA>
__global__ void __launch_bounds__(128, 1)
probe(const float* __restrict__ data, const int* __restrict__ list, int n_list, int mode,
      int n_stages, long long* clocks) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const int pt = threadIdx.x;
    const int lane = pt & 31, warp = pt >> 5;
    const uint32_t sbase = smem_u32(sm);
    long long t0 = clock64();
    for (int st = 0; st < n_stages; st++) {
        const uint32_t dst0 = sbase + (st % NSTAGE) * STAGE_BYTES + pt * 16;
        const int fb = (blockIdx.x + st) & 1;
        // a stage = 128 rows x 16 quads; per thread 16 copies
#pragma unroll 4
        for (int j = 0; j < 16; j++) {
            int row, quad_px;
            const int li = (blockIdx.x * 977 + st * 16) % (n_list - 64);
            if (mode == 0) {
                const int q = pt & 15;                 // 16 quads of the stage
                row = (pt >> 4) + 8 * j;               // 8 rows per step, 16 steps
                quad_px = list[li + q];
            } else if (mode == 1) {
                row = warp * 32 + 2 * j + (lane >> 4);  // 16 quads x 2 rows, quads contiguous
                quad_px = list[li] + 4 * (lane & 15);
            } else if (mode == 2) {
                row = warp * 32 + lane;
                quad_px = list[li + j];
            } else if (mode == 3) {
                row = warp * 32 + 4 * (j & 7) + (lane >> 3);
                quad_px = (list[li + (j >> 3)] & ~31) + 4 * (lane & 7);
            } else {
                row = warp * 32 + 8 * (j & 3) + (lane >> 2);
                quad_px = (list[li + (j >> 2)] & ~15) + 4 * (lane & 3);
            }
            const float* src = data + (size_t)(fb * 128 + row) * K + quad_px;
            if (CA) cp16_ca(dst0 + j * 2048, src); else cp16(dst0 + j * 2048, src);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 3;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (pt == 0) clocks[blockIdx.x] = clock64() - t0;
}

int main() {
    float* d;
    cudaMalloc(&d, (size_t)ROWS * K * 4);
    cudaMemset(d, 0, (size_t)ROWS * K * 4);
    // ring-like quad list: runs of 2..7 quads at scattered positions; and an isolated-quad list
    const int n_list = 1 << 16;
    std::vector<int> ring(n_list), iso(n_list);
    uint32_t x = 777;
    int pos = 0, left = 0;
    for (int i = 0; i < n_list; i++) {
        x = x * 1664525u + 1013904223u;
        if (left == 0) { pos = (int)((x >> 8) % (K / 4 - 8)) * 4; left = 2 + (x >> 28) % 6; }
        ring[i] = pos; pos += 4; left--;
        x = x * 1664525u + 1013904223u;
        iso[i] = (int)((x >> 8) % (K / 4 - 64)) * 4;
    }
    int *dring, *diso;
    cudaMalloc(&dring, n_list * 4); cudaMalloc(&diso, n_list * 4);
    cudaMemcpy(dring, ring.data(), n_list * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(diso, iso.data(), n_list * 4, cudaMemcpyHostToDevice);
    long long* dclk; cudaMalloc(&dclk, 148 * 8);
    cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NSTAGE * STAGE_BYTES);
    cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NSTAGE * STAGE_BYTES);
    const char* names[] = {"K7 mapping, ring-like runs", "K7 mapping, isolated quads", "32 consecutive quads / row",
                           "1 quad x 32 rows", "8 consecutive quads x 4 rows", "4 consecutive quads x 8 rows"};
    for (int ca = 0; ca < 2; ca++)
    for (int v = 0; v < 6; v++) {
        const int mode = v == 0 ? 0 : v == 1 ? 0 : v - 1;
        const int* lst = v == 1 ? diso : dring;
        const int n_stages = 3000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (ca) probe<true><<<148, 128, NSTAGE * STAGE_BYTES>>>(d, lst, n_list, mode, n_stages, dclk);
            else probe<false><<<148, 128, NSTAGE * STAGE_BYTES>>>(d, lst, n_list, mode, n_stages, dclk);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long clk[148]; cudaMemcpy(clk, dclk, sizeof(clk), cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < 148; i++) mean += clk[i]; mean /= 148;
        const double quads = (double)n_stages * 2048;
        printf("%s %-30s: %.3f ms  %.2f us/stage  %.3f quads/clk/SM  %.2f quads/ns/SM  (%.0f MHz)  %.0f GB/s useful  %s\n",
               ca ? ".ca" : ".cg", names[v], ms, ms * 1e3 / n_stages, quads / mean, quads / (ms * 1e6),
               mean / (ms * 1e3), 148.0 * quads * 16 / (ms * 1e6), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
