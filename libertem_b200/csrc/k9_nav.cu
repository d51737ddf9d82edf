// k9_nav.cu -- K9: nav-space post-processing of the centre-of-mass moments on the device
//
// Replaces the numpy pipeline of CoMUDF.get_results (reference src/libertem/udf/com.py:650-717)
// and the 360 x 2 curl sweep of guess_corrections (com.py:145-295) for result buffers that
// already live in HBM (the gathered (n_frames, 3) slab of [m00, m10, m01]):
//   center_shifts (com.py:100-107)  -> apply_correction (2x2 float64 matrix, com.py:110-127)
//   -> regression (mean / least-squares plane over the valid scan positions, com.py:600-648)
//   -> magnitude / divergence / curl with np.gradient stencils (com.py:130-142).
// Arithmetic follows the reference's dtypes: the shifts and raw_com are float32 (divide,
// subtract the reference point); everything after the float64 rotation matrix is float64 and is
// returned as float64, like the reference's result arrays.  Sums (regression normal equations,
// gradient Gram matrix) are reduced in a fixed order (per-block partials, then one block).
//
// The rotation / flip sweep needs no 720 passes: curl(T f) is linear in the four gradient
// fields d(y,x)/d(axis 0,1), so its RMS for ANY 2x2 matrix T is a quadratic form in the 4x4 Gram
// matrix of those fields -- ltb200_com_gradient_gram computes the 10 distinct sums once and
// the host evaluates the 720 candidates in closed form.
#include "common.cuh"

#include <cmath>

namespace ltb {

constexpr int NAV_THREADS = 256;
constexpr int NAV_MAX_BLOCKS = 296;
constexpr int NAV_NSUM = 12;     // 1, y, x, yy, xy, xx, fy, fy*y, fy*x, fx, fx*y, fx*x

template <int NS>
__device__ __forceinline__ void block_partials(double (&v)[NS], double* out /* [gridDim.x][NS] */) {
    __shared__ double red[NAV_THREADS / 32][NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NS; i++) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) red[warp][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double s = 0.0;
        for (int w = 0; w < NAV_THREADS / 32; w++) s += red[w][threadIdx.x];
        out[(size_t)blockIdx.x * NS + threadIdx.x] = s;
    }
}

// raw moments -> float32 shifts / com, float64 corrected field, regression sums
__global__ void __launch_bounds__(NAV_THREADS)
com_field_kernel(const float* __restrict__ raw, int64_t ld_raw, const int32_t* __restrict__ row_of_nav,
                 const uint8_t* __restrict__ valid, int ny, int nx, float cy, float cx, double t00,
                 double t01, double t10, double t11, float* __restrict__ raw_shifts,
                 float* __restrict__ raw_com, double* __restrict__ field, double* __restrict__ partials) {
    const int64_t n = (int64_t)ny * nx;
    double s[NAV_NSUM];
#pragma unroll
    for (int i = 0; i < NAV_NSUM; i++) s[i] = 0.0;
    const float nan = __int_as_float(0x7fc00000);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = row_of_nav ? (int64_t)row_of_nav[i] : i;
        float sy = nan, sx = nan, comy = nan, comx = nan;
        double fy = nan, fx = nan;
        if (row >= 0) {
            const float m00 = raw[row * ld_raw], m10 = raw[row * ld_raw + 1],
                        m01 = raw[row * ld_raw + 2];
            // center_shifts: divide where the sum is non-zero, else the reference point
            const float qy = m00 != 0.f ? __fdiv_rn(m10, m00) : cy;
            const float qx = m00 != 0.f ? __fdiv_rn(m01, m00) : cx;
            sy = __fsub_rn(qy, cy);
            sx = __fsub_rn(qx, cx);
            comy = __fadd_rn(sy, cy);
            comx = __fadd_rn(sx, cx);
            fy = t00 * (double)sy + t01 * (double)sx;
            fx = t10 * (double)sy + t11 * (double)sx;
        }
        raw_shifts[2 * i] = sy;
        raw_shifts[2 * i + 1] = sx;
        raw_com[2 * i] = comy;
        raw_com[2 * i + 1] = comx;
        field[2 * i] = fy;
        field[2 * i + 1] = fx;
        if (valid ? valid[i] != 0 : row >= 0) {
            const double y = (double)(i / nx), x = (double)(i % nx);
            s[0] += 1.0; s[1] += y; s[2] += x; s[3] += y * y; s[4] += x * y; s[5] += x * x;
            s[6] += fy; s[7] += fy * y; s[8] += fy * x;
            s[9] += fx; s[10] += fx * y; s[11] += fx * x;
        }
    }
    block_partials<NAV_NSUM>(s, partials);
}

// one thread: reduce the partials in block order and solve for the regression (3, 2)
// mode: -1 none, 0 subtract mean, 1 least-squares plane c0 + c1 y + c2 x, 2 given coefficients
__global__ void com_regression_kernel(const double* __restrict__ partials, int n_blocks, int mode,
                                      double* __restrict__ regression /* (3,2) in/out */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s[NAV_NSUM];
    for (int i = 0; i < NAV_NSUM; i++) s[i] = 0.0;
    for (int b = 0; b < n_blocks; b++)
        for (int i = 0; i < NAV_NSUM; i++) s[i] += partials[(size_t)b * NAV_NSUM + i];
    double c[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    if (mode == 0 && s[0] > 0) {
        c[0][0] = s[6] / s[0];
        c[0][1] = s[9] / s[0];
    } else if (mode == 1 && s[0] > 0) {
        // normal equations of the plane fit, solved with the coordinates centred on their means
        // (keeps the 3x3 system well conditioned); equivalent to np.linalg.lstsq on [1, y, x]
        const double n = s[0], my = s[1] / n, mx = s[2] / n;
        const double syy = s[3] - n * my * my, sxy = s[4] - n * mx * my, sxx = s[5] - n * mx * mx;
        const double det = syy * sxx - sxy * sxy;
        for (int k = 0; k < 2; k++) {
            const double f = s[6 + 3 * k], fy = s[7 + 3 * k], fx = s[8 + 3 * k];
            const double mf = f / n;
            const double cy_ = fy - my * f, cx_ = fx - mx * f;      // centred cross moments
            double b1 = 0.0, b2 = 0.0;
            if (fabs(det) > 0.0) {
                b1 = (cy_ * sxx - cx_ * sxy) / det;
                b2 = (cx_ * syy - cy_ * sxy) / det;
            } else if (syy > 0.0) {
                b1 = cy_ / syy;              // all valid positions in one scan row / column:
            } else if (sxx > 0.0) {          // minimum-norm solution like lstsq
                b2 = cx_ / sxx;
            }
            c[1][k] = b1;
            c[2][k] = b2;
            c[0][k] = mf - b1 * my - b2 * mx;
        }
    } else if (mode == 2) {
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < 2; k++) c[i][k] = regression[2 * i + k];
    }
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 2; k++) regression[2 * i + k] = c[i][k];
}

// field -= plane (valid positions), store field / components / magnitude
__global__ void __launch_bounds__(NAV_THREADS)
com_apply_regression_kernel(double* __restrict__ field, const int32_t* __restrict__ row_of_nav,
                            const uint8_t* __restrict__ valid, int ny, int nx,
                            const double* __restrict__ regression, double* __restrict__ field_out,
                            double* __restrict__ field_y, double* __restrict__ field_x,
                            double* __restrict__ magnitude) {
    const int64_t n = (int64_t)ny * nx;
    double c[6];
#pragma unroll
    for (int i = 0; i < 6; i++) c[i] = regression[i];
    // np.allclose(result[1:], 0) / np.allclose(regression[0], 0): atol 1e-8 (com.py:634-676)
    const bool has_lin = fabs(c[2]) > 1e-8 || fabs(c[3]) > 1e-8 || fabs(c[4]) > 1e-8 ||
                         fabs(c[5]) > 1e-8;
    const bool has_mean = fabs(c[0]) > 1e-8 || fabs(c[1]) > 1e-8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double fy = field[2 * i], fx = field[2 * i + 1];
        const bool ok = valid ? valid[i] != 0 : (row_of_nav ? row_of_nav[i] >= 0 : true);
        if (ok) {
            if (has_lin) {
                const double y = (double)(i / nx), x = (double)(i % nx);
                fy -= c[0] + y * c[2] + x * c[4];
                fx -= c[1] + y * c[3] + x * c[5];
            } else if (has_mean) {
                fy -= c[0];
                fx -= c[1];
            }
            field[2 * i] = fy;
            field[2 * i + 1] = fx;
        }
        field_out[2 * i] = fy;
        field_out[2 * i + 1] = fx;
        field_y[i] = fy;
        field_x[i] = fx;
        magnitude[i] = sqrt(fy * fy + fx * fx);
    }
}

// np.gradient along one axis: central differences inside, one-sided first order at the edges
__device__ __forceinline__ double grad_axis(const double* f, int64_t i, int64_t stride, int pos,
                                            int len) {
    if (pos == 0) return f[i + stride] - f[i];
    if (pos == len - 1) return f[i] - f[i - stride];
    return (f[i + stride] - f[i - stride]) * 0.5;
}

__global__ void __launch_bounds__(NAV_THREADS)
com_div_curl_kernel(const double* __restrict__ field, int ny, int nx, double* __restrict__ divergence,
                    double* __restrict__ curl) {
    const int64_t n = (int64_t)ny * nx;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i % nx);
        const double* fy = field;          // interleaved (y, x): stride 2 doubles per position
        const double* fx = field + 1;
        const double dyy = grad_axis(fy, 2 * i, 2 * (int64_t)nx, y, ny);
        const double dyx = grad_axis(fy, 2 * i, 2, x, nx);
        const double dxy = grad_axis(fx, 2 * i, 2 * (int64_t)nx, y, ny);
        const double dxx = grad_axis(fx, 2 * i, 2, x, nx);
        divergence[i] = dyy + dxx;     // d(y)/d(axis 0) + d(x)/d(axis 1)
        curl[i] = dyx - dxy;           // d(y)/d(axis 1) - d(x)/d(axis 0)
    }
}

// Gram matrix of the four gradient fields g = (dy/d0, dy/d1, dx/d0, dx/d1) over the window
// rows [r0, r1) x columns [c0, c1), plus the sums of y and x and the count: 10 + 4 + 3 sums
constexpr int GRAM_NS = 17;
__global__ void __launch_bounds__(NAV_THREADS)
com_gradient_gram_kernel(const float* __restrict__ yc, const float* __restrict__ xc, int ny, int nx,
                         int r0, int r1, int c0, int c1, double* __restrict__ partials) {
    double s[GRAM_NS];
#pragma unroll
    for (int i = 0; i < GRAM_NS; i++) s[i] = 0.0;
    const int64_t wn = (int64_t)(r1 - r0) * (c1 - c0);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < wn;
         j += (int64_t)gridDim.x * blockDim.x) {
        const int y = r0 + (int)(j / (c1 - c0)), x = c0 + (int)(j % (c1 - c0));
        const int64_t i = (int64_t)y * nx + x;
        auto grad = [&](const float* f, int64_t stride, int pos, int len) -> double {
            if (pos == 0) return (double)f[i + stride] - (double)f[i];
            if (pos == len - 1) return (double)f[i] - (double)f[i - stride];
            return ((double)f[i + stride] - (double)f[i - stride]) * 0.5;
        };
        const double g[4] = {grad(yc, nx, y, ny), grad(yc, 1, x, nx), grad(xc, nx, y, ny),
                             grad(xc, 1, x, nx)};
        int k = 0;
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = a; b < 4; b++) s[k++] += g[a] * g[b];
#pragma unroll
        for (int a = 0; a < 4; a++) s[10 + a] += g[a];
        s[14] += (double)yc[i];
        s[15] += (double)xc[i];
        s[16] += 1.0;
    }
    block_partials<GRAM_NS>(s, partials);
}

__global__ void nav_reduce_partials_kernel(const double* __restrict__ partials, int n_blocks, int ns,
                                           double* __restrict__ out) {
    const int i = threadIdx.x;
    if (i >= ns) return;
    double s = 0.0;
    for (int b = 0; b < n_blocks; b++) s += partials[(size_t)b * ns + i];
    out[i] = s;
}

// divergence of T (y, x) over the window: min / max (pass 0) or a 5-bin histogram over
// [-range, range] like np.histogram (pass 1; the last bin is closed on the right)
__global__ void __launch_bounds__(NAV_THREADS)
com_div_stats_kernel(const float* __restrict__ yc, const float* __restrict__ xc, int ny, int nx,
                     int r0, int r1, int c0, int c1, double t00, double t01, double t10,
                     double t11, int pass, double range, double* __restrict__ minmax,
                     unsigned long long* __restrict__ hist) {
    const int64_t wn = (int64_t)(r1 - r0) * (c1 - c0);
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < wn;
         j += (int64_t)gridDim.x * blockDim.x) {
        const int y = r0 + (int)(j / (c1 - c0)), x = c0 + (int)(j % (c1 - c0));
        const int64_t i = (int64_t)y * nx + x;
        auto grad = [&](const float* f, int64_t stride, int pos, int len) -> double {
            if (pos == 0) return (double)f[i + stride] - (double)f[i];
            if (pos == len - 1) return (double)f[i] - (double)f[i - stride];
            return ((double)f[i + stride] - (double)f[i - stride]) * 0.5;
        };
        // div(T f) = d(t00 y + t01 x)/d0 + d(t10 y + t11 x)/d1
        const double d = t00 * grad(yc, nx, y, ny) + t01 * grad(xc, nx, y, ny) +
                         t10 * grad(yc, 1, x, nx) + t11 * grad(xc, 1, x, nx);
        if (pass == 0) {
            lo = fmin(lo, d);
            hi = fmax(hi, d);
        } else if (d >= -range && d <= range && range > 0) {
            int b = (int)floor((d + range) / (2.0 * range) * 5.0);
            if (b > 4) b = 4;
            if (b < 0) b = 0;
            atomicAdd(&hist[b], 1ull);
        }
    }
    if (pass == 0) {
        // order-independent reductions: atomics on the bit patterns of non-negative offsets are
        // awkward for doubles; a per-block store + host reduce over <= 296 blocks is enough
        __shared__ double slo[NAV_THREADS], shi[NAV_THREADS];
        slo[threadIdx.x] = lo;
        shi[threadIdx.x] = hi;
        __syncthreads();
        for (int o = NAV_THREADS / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                slo[threadIdx.x] = fmin(slo[threadIdx.x], slo[threadIdx.x + o]);
                shi[threadIdx.x] = fmax(shi[threadIdx.x], shi[threadIdx.x + o]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            minmax[2 * blockIdx.x] = slo[0];
            minmax[2 * blockIdx.x + 1] = shi[0];
        }
    }
}

static int nav_blocks(int64_t n) {
    int64_t b = (n + NAV_THREADS - 1) / NAV_THREADS;
    if (b > NAV_MAX_BLOCKS) b = NAV_MAX_BLOCKS;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace ltb

using namespace ltb;

extern "C" size_t ltb200_com_workspace(int ny, int nx) {
    // float64 field (ny*nx, 2) + per-block partial sums / min-max
    return (size_t)ny * nx * 2 * sizeof(double) + (size_t)NAV_MAX_BLOCKS * GRAM_NS * sizeof(double) +
           256;
}

extern "C" int ltb200_com_postprocess(const float* raw, int64_t ld_raw, const int32_t* row_of_nav,
                                      const uint8_t* valid, int ny, int nx, double cy, double cx,
                                      const double* transform, int regression_mode,
                                      double* regression, float* raw_shifts, float* raw_com,
                                      double* field, double* field_y, double* field_x,
                                      double* magnitude, double* divergence, double* curl,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    LTB_REQUIRE(ny >= 2 && nx >= 2, "com_postprocess: np.gradient needs at least 2 scan positions "
                                    "per axis, got %d x %d", ny, nx);
    LTB_REQUIRE(raw && transform && regression && raw_shifts && raw_com && field && field_y &&
                    field_x && magnitude && divergence && curl && workspace,
                "com_postprocess: NULL pointer");
    LTB_REQUIRE(ld_raw >= 3, "com_postprocess: raw rows hold (m00, m10, m01)");
    LTB_REQUIRE(regression_mode >= -1 && regression_mode <= 2, "com_postprocess: bad mode");
    LTB_REQUIRE(workspace_bytes >= ltb200_com_workspace(ny, nx), "com_postprocess: workspace");
    LTB_REQUIRE((uintptr_t)workspace % 8 == 0, "com_postprocess: workspace alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)ny * nx;
    double* fld = (double*)workspace;
    double* partials = fld + 2 * n;
    const int blocks = nav_blocks(n);
    com_field_kernel<<<blocks, NAV_THREADS, 0, st>>>(raw, ld_raw, row_of_nav, valid, ny, nx,
                                                     (float)cy, (float)cx, transform[0],
                                                     transform[1], transform[2], transform[3],
                                                     raw_shifts, raw_com, fld, partials);
    com_regression_kernel<<<1, 32, 0, st>>>(partials, blocks, regression_mode, regression);
    com_apply_regression_kernel<<<blocks, NAV_THREADS, 0, st>>>(fld, row_of_nav, valid, ny, nx,
                                                                regression, field, field_y,
                                                                field_x, magnitude);
    com_div_curl_kernel<<<blocks, NAV_THREADS, 0, st>>>(fld, ny, nx, divergence, curl);
    count_launch(4);
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

extern "C" int ltb200_com_gradient_gram(const float* y_centers, const float* x_centers, int ny,
                                        int nx, int r0, int r1, int c0, int c1, double* sums17,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    LTB_REQUIRE(ny >= 2 && nx >= 2, "com_gradient_gram: needs at least 2 positions per axis");
    LTB_REQUIRE(0 <= r0 && r0 < r1 && r1 <= ny && 0 <= c0 && c0 < c1 && c1 <= nx,
                "com_gradient_gram: empty or out-of-range window");
    LTB_REQUIRE(y_centers && x_centers && sums17 && workspace, "com_gradient_gram: NULL pointer");
    LTB_REQUIRE(workspace_bytes >= (size_t)NAV_MAX_BLOCKS * GRAM_NS * sizeof(double),
                "com_gradient_gram: workspace");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = nav_blocks((int64_t)(r1 - r0) * (c1 - c0));
    com_gradient_gram_kernel<<<blocks, NAV_THREADS, 0, st>>>(y_centers, x_centers, ny, nx, r0, r1,
                                                             c0, c1, (double*)workspace);
    nav_reduce_partials_kernel<<<1, 32, 0, st>>>((const double*)workspace, blocks, GRAM_NS, sums17);
    count_launch(2);
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}

extern "C" int ltb200_com_divergence_stats(const float* y_centers, const float* x_centers, int ny,
                                           int nx, int r0, int r1, int c0, int c1,
                                           const double* transform, int pass, double range,
                                           double* minmax_blocks, int* n_blocks,
                                           unsigned long long* hist5, void* stream) {
    LTB_REQUIRE(ny >= 2 && nx >= 2, "com_divergence_stats: needs at least 2 positions per axis");
    LTB_REQUIRE(0 <= r0 && r0 < r1 && r1 <= ny && 0 <= c0 && c0 < c1 && c1 <= nx,
                "com_divergence_stats: empty or out-of-range window");
    LTB_REQUIRE(y_centers && x_centers && transform && n_blocks, "com_divergence_stats: NULL");
    LTB_REQUIRE(pass == 0 ? minmax_blocks != nullptr : hist5 != nullptr,
                "com_divergence_stats: output of the pass missing");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = nav_blocks((int64_t)(r1 - r0) * (c1 - c0));
    *n_blocks = blocks;
    com_div_stats_kernel<<<blocks, NAV_THREADS, 0, st>>>(y_centers, x_centers, ny, nx, r0, r1, c0,
                                                         c1, transform[0], transform[1],
                                                         transform[2], transform[3], pass, range,
                                                         minmax_blocks, hist5);
    count_launch();
    LTB_CUDA_CHECK(cudaGetLastError());
    return LTB_OK;
}
