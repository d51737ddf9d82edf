// api.cu -- error state, device info, TMA descriptor encoding for libltb200.so
#include "common.cuh"
#include <cstring>
#include <mutex>

namespace ltb {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;
static thread_local int g_last_kernel = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }
void set_last_kernel(int id) { g_last_kernel = id; }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int encode_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, size_t elem_bytes,
                   uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0,
                   uint32_t box1) {
    (void)elem_bytes;
    return encode_tmap_2d_sw(out, base, dt, dim0, dim1, stride1_bytes, box0, box1,
                             CU_TENSOR_MAP_SWIZZLE_NONE);
}

int encode_tmap_2d_sw(CUtensorMap* out, const void* base, CUtensorMapDataType dt, uint64_t dim0,
                      uint64_t dim1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1,
                      CUtensorMapSwizzle swizzle) {
    return encode_tmap_2d_sw_promo(out, base, dt, dim0, dim1, stride1_bytes, box0, box1, swizzle,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

int encode_tmap_2d_sw_promo(CUtensorMap* out, const void* base, CUtensorMapDataType dt,
                            uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0,
                            uint32_t box1, CUtensorMapSwizzle swizzle,
                            CUtensorMapL2promotion promotion) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable");
        return LTB_ERR_CUDA;
    }
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    promotion, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu, stride %llu, "
                  "box %u x %u)",
                  (int)r, (unsigned long long)dim0, (unsigned long long)dim1,
                  (unsigned long long)stride1_bytes, box0, box1);
        return LTB_ERR_CUDA;
    }
    return LTB_OK;
}

}  // namespace ltb

extern "C" {

int ltb200_abi_version(void) { return LTB200_ABI_VERSION; }
const char* ltb200_last_error(void) { return ltb::g_err; }
int ltb200_last_kernel(void) { return ltb::g_last_kernel; }
int64_t ltb200_launch_count(int reset) {
    int64_t v = ltb::g_launches;
    if (reset) ltb::g_launches = 0;
    return v;
}

int ltb200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                       int64_t* smem_optin_bytes) {
    int v = 0;
    LTB_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    if (sm_count) *sm_count = v;
    LTB_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
    if (cc_major) *cc_major = v;
    LTB_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
    if (cc_minor) *cc_minor = v;
    LTB_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (smem_optin_bytes) *smem_optin_bytes = v;
    return LTB_OK;
}

}  // extern "C"
