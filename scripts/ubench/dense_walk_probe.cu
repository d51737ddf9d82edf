// Micro-benchmark for the dense-walk form of K7: what does the memory system deliver when a
// persistent CTA per SM streams [32 px x 128 frames] TMA boxes (128-byte rows, 1 MiB apart) of
// its own frame block in (a) row-major order, (b) ring order (boxes sorted by the innermost ring
// they touch, ties row-major), (c) random order?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dense_walk_probe dense_walk_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
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
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                       uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                 "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}

constexpr int SIG = 512 * 512, BOX_PX = 32, FB = 128, BOX_BYTES = BOX_PX * FB * 4, NSTAGE = 10;

__global__ void __launch_bounds__(32, 1)
walk_kernel(const __grid_constant__ CUtensorMap tm, const int* __restrict__ box_px, int n_boxes, int n_fb) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + NSTAGE * BOX_BYTES);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint32_t it_issue = 0, it_wait = 0;
    const long total = (long)((n_fb - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * n_boxes;
    long issued = 0;
    int fb_i = blockIdx.x, b_i = 0;
    auto issue = [&]() {
        const int s = it_issue % NSTAGE;
        mbar_expect(&full[s], BOX_BYTES);
        tma_2d(smem_u32(sm + s * BOX_BYTES), &tm, box_px[b_i], fb_i * FB, &full[s], pol);
        it_issue++;
        issued++;
        if (++b_i == n_boxes) { b_i = 0; fb_i += gridDim.x; }
    };
    for (int s = 0; s < NSTAGE && issued < total; s++) issue();
    for (long w = 0; w < total; w++) {
        const int s = it_wait % NSTAGE;
        mbar_wait(&full[s], (it_wait / NSTAGE) & 1);
        it_wait++;
        if (issued < total) issue();
    }
}

int main(int argc, char** argv) {
    const int n_fb = argc > 1 ? atoi(argv[1]) : 148;
    const long n_frames = (long)n_fb * FB;
    float* buf;
    cudaMalloc(&buf, n_frames * SIG * 4);
    cudaMemset(buf, 0, n_frames * SIG * 4);
    const int n_boxes = SIG / BOX_PX;
    std::vector<int> seq(n_boxes), ring(n_boxes), rnd(n_boxes), ring_run(n_boxes);
    std::vector<int> rmin(n_boxes);
    const double bw = 364.0 / 32;
    for (int b = 0; b < n_boxes; b++) {
        seq[b] = b * BOX_PX;
        const int y = b / 16, x0 = (b % 16) * 32;
        double m = 1e9;
        for (int x = x0; x < x0 + 32; x++) m = std::min(m, std::hypot(x - 256.0, y - 256.0));
        rmin[b] = (int)(m / bw);
    }
    ring = seq;
    std::stable_sort(ring.begin(), ring.end(), [&](int a, int b) { return rmin[a / 32] < rmin[b / 32]; });
    rnd = seq;
    srand(1);
    for (int i = n_boxes - 1; i > 0; i--) std::swap(rnd[i], rnd[rand() % (i + 1)]);
    // ring order but x-adjacent pairs kept together (key of the pair = min of the two)
    {
        std::vector<int> pairs(n_boxes / 2);
        for (int i = 0; i < n_boxes / 2; i++) pairs[i] = i;
        std::stable_sort(pairs.begin(), pairs.end(), [&](int a, int b) {
            return std::min(rmin[2 * a], rmin[2 * a + 1]) < std::min(rmin[2 * b], rmin[2 * b + 1]);
        });
        for (int i = 0; i < n_boxes / 2; i++) {
            ring_run[2 * i] = pairs[i] * 64;
            ring_run[2 * i + 1] = pairs[i] * 64 + 32;
        }
    }
    CUtensorMap tm;
    {
        cuInit(0);
        cuuint64_t dims[2] = {(cuuint64_t)SIG, (cuuint64_t)n_frames};
        cuuint64_t strides[1] = {(cuuint64_t)SIG * 4};
        cuuint32_t box[2] = {BOX_PX, FB};
        cuuint32_t es[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, es,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", (int)r); return 1; }
    }
    int* d_list;
    cudaMalloc(&d_list, n_boxes * 4);
    const size_t smem = NSTAGE * BOX_BYTES + 256;
    cudaFuncSetAttribute(walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[4] = {"row-major", "ring order", "ring order, x-pairs", "random"};
    std::vector<int>* lists[4] = {&seq, &ring, &ring_run, &rnd};
    for (int rep = 0; rep < 2; rep++)
        for (int k = 0; k < 4; k++) {
            cudaMemcpy(d_list, lists[k]->data(), n_boxes * 4, cudaMemcpyHostToDevice);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0);
            walk_kernel<<<148, 32, smem>>>(tm, d_list, n_boxes, n_fb);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%-22s: %8.3f ms  %7.1f GB/s  (%s)\n", names[k], ms, n_frames * SIG * 4.0 / ms / 1e6,
                   cudaGetErrorString(err));
        }
    return 0;
}
