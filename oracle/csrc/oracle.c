/* C restatement of the reference's numba dense x sparse kernel -- TEST INFRASTRUCTURE ONLY.
 *
 * Restates /root/reference/src/libertem/common/numba/__init__.py:169-184 (_rmatmul_csr):
 * loop over sparse rows k ascending; copy column k of the dense left matrix into a row
 * buffer; for every non-zero (k, m, v) of that row: res_t[m, f] += rowbuf[f] * v for all f.
 * Arithmetic: float32 product, float32 accumulate (numba fastmath may contract to FMA; for
 * the integer-valued inputs the bit-exact parity tests use, both are exact).
 *
 * Never linked into the product library (libertem_b200/csrc); loaded only by oracle/udf_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>

void oracle_rmatmul_csr_f32(const float* left, int64_t F, int64_t K,
                            const float* data, const int32_t* indices, const int32_t* indptr,
                            float* res_t /* (n_cols, F) zero-initialised */)
{
    float* rowbuf = (float*)malloc(sizeof(float) * (size_t)F);
    for (int64_t k = 0; k < K; k++) {
        int32_t off = indptr[k], items = indptr[k + 1] - off;
        if (items <= 0) continue;
        for (int64_t f = 0; f < F; f++) rowbuf[f] = left[f * K + k];
        for (int32_t i = 0; i < items; i++) {
            int32_t col = indices[off + i];
            float v = data[off + i];
            float* dst = res_t + (int64_t)col * F;
            for (int64_t f = 0; f < F; f++) {
                float tmp = rowbuf[f] * v;
                dst[f] += tmp;
            }
        }
    }
    free(rowbuf);
}

/* float32 dense x complex64 sparse (RadialFourierAnalysis masks): interleaved (re, im). */
void oracle_rmatmul_csr_c64(const float* left, int64_t F, int64_t K,
                            const float* data /* 2*nnz */, const int32_t* indices,
                            const int32_t* indptr, float* res_t /* (n_cols, F, 2) */)
{
    float* rowbuf = (float*)malloc(sizeof(float) * (size_t)F);
    for (int64_t k = 0; k < K; k++) {
        int32_t off = indptr[k], items = indptr[k + 1] - off;
        if (items <= 0) continue;
        for (int64_t f = 0; f < F; f++) rowbuf[f] = left[f * K + k];
        for (int32_t i = 0; i < items; i++) {
            int32_t col = indices[off + i];
            float vr = data[2 * (off + i)], vi = data[2 * (off + i) + 1];
            float* dst = res_t + (int64_t)col * F * 2;
            for (int64_t f = 0; f < F; f++) {
                dst[2 * f] += rowbuf[f] * vr;
                dst[2 * f + 1] += rowbuf[f] * vi;
            }
        }
    }
    free(rowbuf);
}
