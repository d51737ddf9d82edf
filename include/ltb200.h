/* ltb200.h -- C ABI of the B200-native masked-reduction engine (libltb200.so).
 *
 * Drop-in boundary for LiberTEM's ApplyMasksUDF / CoMUDF / SumUDF / SumSigUDF hot path.
 * Every entry point replaces one seam of the reference (paths relative to the LiberTEM
 * source tree, src/libertem/...):
 *
 *   ltb200_masks_dense      <- ApplyMasksEngine.process_flat / _process_flat_{torch,standard}
 *                              (udf/masks.py:31-83) + the `+=` of ApplyMasksUDF.process_tile
 *                              (udf/masks.py:383-392); CoMUDF.process_tile (udf/com.py:577-582)
 *                              is the same call with the 3 CoM mask rows appended; SumSigUDF
 *                              (udf/sumsigudf.py:28-38) is an all-ones mask row.
 *   ltb200_masks_dense_f64  <- the same seam when np.result_type(input, mask) is float64
 *                              (udf/masks.py:360-368 dtype rule).
 *   ltb200_masks_shifted    <- ApplyMasksEngine.process_frame_shifted (udf/masks.py:85-124).
 *   ltb200_masks_csc        <- ApplyMasksEngine._process_flat_spsp -> rmatmul
 *                              (udf/masks.py:68-69, common/numba/__init__.py:90-184).
 *                              (`sig_sum` of ltb200_masks_dense fuses SumUDF, udf/sum.py:44-49,
 *                              into the same pass; uint16 tiles are ingested natively.)
 *   ltb200_group_masks(_tc) <- ApplyMasksUDF with radial_mask_factory masks
 *                              (analysis/radialfourier.py:106-146,184-194); _tc = tensor cores.
 *   ltb200_masks_dense_tc   <- the process_flat seam again, explicitly on the tensor cores
 *                              (_tc_u16: uint16 tiles; ltb200_masks_dense_i8: uint16 tiles x
 *                              int8 masks, exact, on the int8 tensor cores).
 *   ltb200_synth_fill       <- test/bench data source standing in for MemoryDataSet contents
 *                              (io/dataset/memory.py:202-452); twin of oracle/synth.py.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers owned by the caller (torch tensors on the host
 *     side); the library allocates nothing persistent.  `workspace` is caller-provided
 *     device scratch of at least ltb200_*_workspace(...) bytes (may be NULL when that is 0).
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous w.r.t. the host.
 *   - Return value: 0 on success, negative LTB_ERR_* otherwise; ltb200_last_error() gives a
 *     thread-local message.  Nothing throws, nothing falls back to the CPU.
 *   - Matrices are row-major; `ld_*` are leading dimensions in ELEMENTS.
 */
#ifndef LTB200_H
#define LTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTB200_ABI_VERSION 1

#if defined(__GNUC__)
#define LTB_API __attribute__((visibility("default")))
#else
#define LTB_API
#endif

/* element types of input tiles (numpy dtype of the dataset / tile) */
enum ltb200_dtype {
    LTB_F32 = 0,
    LTB_U16 = 1,
    LTB_U8 = 2,
    LTB_I16 = 3,
    LTB_F64 = 4,
    LTB_I32 = 5,
    LTB_U32 = 6,
    LTB_I64 = 7,
    LTB_U64 = 8,
    LTB_I8 = 9
};

enum ltb200_error {
    LTB_OK = 0,
    LTB_ERR_ARG = -1,         /* invalid argument (shape, alignment, NULL pointer) */
    LTB_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed */
    LTB_ERR_UNSUPPORTED = -3, /* valid request this build has no kernel for */
    LTB_ERR_WORKSPACE = -4    /* workspace too small */
};

LTB_API int ltb200_abi_version(void);
LTB_API const char* ltb200_last_error(void);

/* sm count / compute capability / opt-in shared memory of `device` */
LTB_API int ltb200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                       int64_t* smem_optin_bytes);

/* ---------------------------------------------------------------------------------------
 * Dense masked reduction (K1):  out[f, m] (+)= sum_k tile[f, k] * masks[m, k]
 *   tile   : (n_frames, sig_size) of `tile_dtype`, leading dimension ld_tile
 *   masks  : (n_masks, sig_size) float32 -- the physical layout of the reference's
 *            F-ordered (sig_size, n_masks) mask matrix (common/container.py:86-91)
 *   out    : (n_frames, n_masks) float32, leading dimension ld_out
 *   accumulate: 0 -> out = result, 1 -> out += result (partial sig tiles)
 *   sig_sum: optional (sig_size,) float32, sig_sum[k] += sum_f tile[f, k] (SumUDF), or NULL
 * float32 arithmetic (FFMA), blocked accumulation (chains <= 512 terms).
 * ------------------------------------------------------------------------------------- */
LTB_API size_t ltb200_masks_dense_workspace(int64_t n_frames, int64_t sig_size, int n_masks,
                                    int with_sig_sum);
LTB_API int ltb200_masks_dense(const void* tile, int tile_dtype, int64_t n_frames, int64_t sig_size,
                       int64_t ld_tile, const float* masks, int n_masks, int64_t ld_masks,
                       float* out, int64_t ld_out, int accumulate, float* sig_sum,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Same contraction in float64 (int32/int64/float64 inputs or float64 masks). */
LTB_API int ltb200_masks_dense_f64(const void* tile, int tile_dtype, int64_t n_frames,
                           int64_t sig_size, int64_t ld_tile, const double* masks, int n_masks,
                           int64_t ld_masks, double* out, int64_t ld_out, int accumulate,
                           void* stream);

/* ---------------------------------------------------------------------------------------
 * The same dense contraction on the tensor cores (K6): tcgen05.mma kind::tf32 with the
 * split-TF32 scheme (hi/lo parts of tile and masks, float32 accumulation in TMEM cut into
 * chains of `chain` x 32 pixels that are summed in float32 registers; chain <= 0 -> default).
 * 9-16 and 25-32 columns per pass use three of the four hi/lo products (lo(tile) x lo(mask),
 * below 2^-21 |x||m| per term, is dropped); 25-32 columns run the 20-warp form with separate
 * accumulator-drain warps.  Environment switches for A/B runs, read per call: LTB200_K6_CHAIN,
 * LTB200_K6_THREE=0, LTB200_K6_DW=0|1|2, LTB200_K6_ISSUERS=1|2 (csrc/k6_tensor.cu).
 * float32 tiles, sig_size % 4 == 0, 16-byte aligned rows.  ltb200_masks_dense routes wide
 * float32 stacks here by itself (see ltb200_set_k1_variant); this entry point is the
 * explicit form used by the parity tests.  Returns LTB_ERR_UNSUPPORTED for shapes the TMA /
 * UMMA layouts cannot take.
 * ------------------------------------------------------------------------------------- */
LTB_API size_t ltb200_masks_dense_tc_workspace(int64_t n_frames, int64_t sig_size, int n_masks);
LTB_API int ltb200_masks_dense_tc(const float* tile, int64_t n_frames, int64_t sig_size,
                                  int64_t ld_tile, const float* masks, int n_masks,
                                  int64_t ld_masks, float* out, int64_t ld_out, int accumulate,
                                  int chain, void* workspace, size_t workspace_bytes,
                                  void* stream);

/* uint16 tiles on the same kernel (the dtype conversion of the reference's tile decode,
 * io/dataset/base/backend.py:69-117, happens in registers: hi = top 11 bits, lo = the rest,
 * both exact TF32 numbers, so integer data x binary masks is bit-exact), 1..16 columns,
 * sig_size % 8 == 0 and >= 256.  `sig_sum` (nullable, (sig_size) float32) fuses SumUDF
 * (udf/sum.py:44-49) into the pass: extra warps sum every staged tile over its frames as exact
 * 64-bit integers; sig_sum[k] += that sum, rounded once. */
LTB_API size_t ltb200_masks_dense_tc_u16_workspace(int64_t n_frames, int64_t sig_size,
                                                   int n_masks, int with_sig_sum);
LTB_API int ltb200_masks_dense_tc_u16(const uint16_t* tile, int64_t n_frames, int64_t sig_size,
                                      int64_t ld_tile, const float* masks, int n_masks,
                                      int64_t ld_masks, float* out, int64_t ld_out,
                                      int accumulate, int chain, float* sig_sum, void* workspace,
                                      size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Integer fast path (K8): uint16 (LTB_U16) or uint8 (LTB_U8) tiles x int8 masks (binary virtual-detector masks, all-ones
 * for SumSigUDF, small integer weights) on the int8 tensor cores (tcgen05.mma kind::i8).  For
 * integer data and integer masks the reference's float32 sums (udf/masks.py:59-77) are exact,
 * so exact int32 accumulation reproduces them bit for bit; the bytes of the TMA-staged tile
 * are the MMA operand as they land in shared memory -- no per-pixel instruction runs.
 * out = float32 of the exact integer result.  `sig_sum` (nullable) fuses SumUDF
 * (udf/sum.py:44-49) as a second MMA over the same stage.  1..32 columns, 16-byte aligned rows,
 * sig_size >= 256 (uint16) / 512 (uint8); signals beyond 65536 pixels are K-split so that the
 * int32 accumulators stay exact (<= 4 Mi pixels); LTB_ERR_UNSUPPORTED otherwise.
 * ------------------------------------------------------------------------------------- */
LTB_API size_t ltb200_masks_dense_i8_workspace(int tile_dtype, int64_t n_frames, int64_t sig_size,
                                               int n_masks, int with_sig_sum);
LTB_API int ltb200_masks_dense_i8(const void* tile, int tile_dtype, int64_t n_frames,
                                  int64_t sig_size, int64_t ld_tile, const int8_t* masks,
                                  int n_masks, int64_t ld_masks, float* out, int64_t ld_out,
                                  int accumulate, float* sig_sum, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* tuning / test knob: which dense kernel ltb200_masks_dense uses for float32 tiles.
 * 0 = auto (default), 1 = FFMA2 even/odd-pixel accumulator pairs, 2 = FFMA2 mask-pair
 * accumulators, 3 = tcgen05 tensor-core kernel (K6) whenever the shape allows */
LTB_API int ltb200_set_k1_variant(int variant);

/* which kernel the last ltb200_masks_dense call on this thread selected:
 * 1 = TMA-staged kernel (even/odd tile), 3 = TMA-staged kernel (mask-pair tile),
 * 2 = generic kernel, 6 = tcgen05 tensor-core kernel; and for the other entry points:
 * 8 = int8 tensor-core kernel (K8), 20 = sparse CSC kernel (K2), 4 = group-sparse FFMA2
 * kernel (K4), 5 / 50 = shifted-mask kernel (K5, warp-per-frame / banded), 7 / 70 / 71 = group-sparse tensor-core kernel (K7)
 * with the ring-major / quad-banded / mirror-symmetric plan
 * (diagnostics / tests; the nav-space kernels K9 do not change it) */
LTB_API int ltb200_last_kernel(void);
/* number of kernel launches issued by this library on this thread since the last reset */
LTB_API int64_t ltb200_launch_count(int reset);

/* ---------------------------------------------------------------------------------------
 * Sparse masked reduction (K2): masks given as CSC over (sig_size, n_masks), i.e. for each
 * mask m the entries [indptr[m], indptr[m+1]) of (indices = pixel k ascending, values).
 * out[f, m] (+)= sum_i tile[f, indices[i]] * values[i]
 * ------------------------------------------------------------------------------------- */
LTB_API int ltb200_masks_csc(const void* tile, int tile_dtype, int64_t n_frames, int64_t sig_size,
                     int64_t ld_tile, const int32_t* indptr, const int32_t* indices,
                     const float* values, int n_masks, float* out, int64_t ld_out,
                     int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------
 * Shifted masks (K5): per frame f the masks are displaced by (dy, dx) = shifts[f] (or shifts[0]
 * when per_frame == 0): out[f, m] (+)= sum over the overlap of tile[f, y, x] * masks[m, y-dy, x-dx]
 * (ApplyMasksEngine.process_frame_shifted, udf/masks.py:85-124).  2D signals, float32 masks.
 * ------------------------------------------------------------------------------------- */
LTB_API int ltb200_masks_shifted(const void* tile, int tile_dtype, int64_t n_frames, int sig_y,
                                 int sig_x, int64_t ld_tile, const float* masks, int n_masks,
                                 int64_t ld_masks, const int32_t* shifts, int per_frame,
                                 float* out, int64_t ld_out, int accumulate, void* stream);
/* Banded form for float32 results (the default when it applies).  The frames are visited in
 * the order `order` (device, n_frames int32: a permutation that sorts the frames by dy; the
 * identity for per_frame == 0), in chunks of 256: a block keeps the row band of 4 or 8 masks a
 * chunk needs -- the band's frame rows widened by the chunk's dy span -- in <= 48 KiB of shared
 * memory and streams the chunk through it.  max_span: the largest (dy_last - dy_first) over the
 * chunks of 256 consecutive entries of `order` (the caller sorted, so it knows).  Band partial
 * sums go to `workspace` (ltb200_masks_shifted_banded_workspace bytes; 0 = no band plan for
 * this geometry, use ltb200_masks_shifted) and are added in band order (deterministic). */
LTB_API size_t ltb200_masks_shifted_banded_workspace(int64_t n_frames, int sig_y, int sig_x,
                                                     int n_masks, int max_span);
LTB_API int ltb200_masks_shifted_banded(const void* tile, int tile_dtype, int64_t n_frames,
                                        int sig_y, int sig_x, int64_t ld_tile, const float* masks,
                                        int n_masks, int64_t ld_masks, const int32_t* shifts,
                                        int per_frame, const int32_t* order, int max_span,
                                        float* out, int64_t ld_out, int accumulate,
                                        void* workspace, size_t workspace_bytes, void* stream);
/* float64 masks / accumulation / result: the reference's dtype rule for float64 masks or frames
 * (result_type(input, mask), udf/masks.py:360-368); tile dtypes f32, f64, u8, u16, i16, i32 */
LTB_API int ltb200_masks_shifted_f64(const void* tile, int tile_dtype, int64_t n_frames, int sig_y,
                                     int sig_x, int64_t ld_tile, const double* masks, int n_masks,
                                     int64_t ld_masks, const int32_t* shifts, int per_frame,
                                     double* out, int64_t ld_out, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------
 * Group-sparse masked reduction (K4) -- RadialFourierAnalysis: masks come in groups (rings)
 * whose members (orders) share one pixel support.  Per group g: entries
 * [group_off[g], group_off[g+1]) (multiples of 128, zero-weight padded) with pixel index
 * entry_px[e]; table_packed holds 28 "pair rows" x (2 * n_entries) floats: row p = the
 * (re, im) / (col 2p, col 2p+1) weights of every entry, in the bank-conflict-free order
 * documented in libertem_b200/group_masks.py.  out[f, (g*n_pairs + p)*2 + {0,1}] (+)= sums,
 * i.e. a complex64 (n_frames, n_groups*n_pairs) matrix.  float32 tiles only.
 * group_off is passed twice: host copy (validated) and device copy (read by the kernel).
 * ------------------------------------------------------------------------------------- */
LTB_API size_t ltb200_group_masks_workspace(void);
LTB_API int ltb200_group_masks(const void* tile, int tile_dtype, int64_t n_frames,
                               int64_t sig_size, int64_t ld_tile, const int32_t* entry_px,
                               const float* table_packed, const int32_t* group_off_host,
                               const int32_t* group_off_dev, int n_groups, int n_pairs,
                               float* out, int64_t ld_out, int accumulate, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * The same group-sparse contraction on the tensor cores (K7): tcgen05.mma kind::tf32 with the
 * split-TF32 scheme of K6, accumulators in TMEM.  Same entry list / offsets as K4 (offsets
 * multiples of 64).  table_split is (N, n_entries) float32 row-major with
 * N = ltb200_group_masks_tc_columns(n_pairs): for the real column r = 2*pair + {0 re, 1 im},
 * h = r / (N/4), j = r % (N/4): row h*(N/2) + j holds hi(weight), row h*(N/2) + N/4 + j holds
 * lo(weight) (hi = weight rounded to TF32, lo = weight - hi rounded to TF32; unused rows zero).
 * chain: 32-entry sub-tiles per TMEM accumulation chain (<= 0 -> default 2).
 * ------------------------------------------------------------------------------------- */
LTB_API int ltb200_group_masks_tc_columns(int n_pairs);
LTB_API int ltb200_group_masks_tc(const float* tile, int64_t n_frames, int64_t sig_size,
                                  int64_t ld_tile, const int32_t* entry_px,
                                  const float* table_split, const int32_t* group_off_host,
                                  const int32_t* group_off_dev, int n_groups, int n_pairs,
                                  float* out, int64_t ld_out, int accumulate, int chain,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* Quad / banded plan (the default for 16-byte aligned frame rows): entry_px lists QUADS -- the
 * first pixel (a multiple of 4) of 4 consecutive pixels that are gathered with one 16-byte
 * copy per frame; the weight table has 4 entries per quad, zero for pixels outside the ring.
 * The n_groups = n_bands x n_rings groups are (pixel band, ring) pairs, band-major (group
 * b * n_rings + r = the quads of ring r inside band b of the flattened signal; n_bands >= 1).
 * Work is scheduled band by band over a few frame blocks at a time, so the 128-byte lines that
 * several rings share are fetched from DRAM once and served from L2 (a ring crosses an image
 * row in runs of ~18 pixels; ring-by-ring gathering moves 2.2-2.7x the bytes it uses).  For
 * n_bands > 1 the band partial sums go to `workspace` (ltb200_group_masks_tc_workspace bytes)
 * and are added in fixed band order into out (n_frames, >= n_rings * n_pairs * 2). */
LTB_API size_t ltb200_group_masks_tc_workspace(int64_t n_frames, int n_groups, int n_pairs,
                                               int n_bands);
LTB_API int ltb200_group_masks_tc_banded(const float* tile, int64_t n_frames, int64_t sig_size,
                                         int64_t ld_tile, const int32_t* entry_px,
                                         const float* table_split,
                                         const int32_t* group_off_host,
                                         const int32_t* group_off_dev, int n_groups, int n_pairs,
                                         int n_bands, float* out, int64_t ld_out, int accumulate,
                                         int chain, void* workspace, size_t workspace_bytes,
                                         void* stream);

/* Dense-walk plan (K10, csrc/k10_walk.cu; the default for RadialFourierAnalysis when
 * sig_size % 32 == 0 and the frame rows are 16-byte aligned): the same contraction as
 * ltb200_group_masks_tc (reference analysis/radialfourier.py:106-146 masks through
 * udf/masks.py:59-77), but every pixel crosses the L2 -> SM fabric once inside a dense TMA box
 * [128 frames x 32 px]; the boxes are visited sorted by the first group they touch and the
 * groups are separated by the weights (an op = 8 consecutive pixels x one group touching
 * them).  The groups of even / odd id form two independent pipelines (index k = 0 / 1 below).
 * All lists come from libertem_b200/walk_plan.py (build_walk), are the same for every block of
 * 128 frames and are final: the kernel only shifts and masks their words.
 *   boxes   (n_visits) uint32: first pixel of the box (multiple of 32) | 4-bit mask of the
 *           8-pixel slices in use; a multiple of 4 visits per segment;
 *   ops_c   uint32 word lists of the four MMA issuers c = group id % 4 (issuers c and c + 2
 *           belong to pipeline c % 2), in walk order, a multiple of 4 words per segment; a
 *           word is an op of the issuer or the marker of a box without ops of the issuer:
 *           bits 0-2 accumulator buffer, 3 / 4 first / last op of an accumulation chain, 5 / 6
 *           first / last word of the issuer in its box, 7 no tensor work (marker, padding),
 *           8-9 slice of the box, 10 mbarrier parity of the issuer's wait for the drain of the
 *           buffer (first op), 11-12 A stage of the box, 13 mbarrier parity of the A stage;
 *   events_k uint32 per accumulation chain of pipeline k: bits 0-2 buffer, 3 register slot,
 *           4 last chain of the group in this segment, 5 mbarrier parity, 6-7 the issuer that
 *           uses the buffer next, 8.. group id;
 *   table_c (n_stages_c, 112, 32) float32: the weight blocks of the ops of issuer c in list
 *           order, 4 per stage (markers take no slot; a segment starts on a new stage), as the
 *           byte image of a shared-memory stage: rows [hi(r) | lo(r)], r = 2 * pair + {0 re,
 *           1 im} < 56, the 8 weights of op j at floats [8 j, 8 j + 8) of a row, 16-byte
 *           chunks XOR-swizzled with (row & 7);
 *   seg_off_host (11, n_segments + 1) int32 on the HOST: visit, word_0..3, table-stage_0..3,
 *           event_0..1 offsets of the segments (independent work items; a group receives sums
 *           from <= 2 segments).
 * out (n_frames, >= n_groups * n_pairs * 2) float32 = complex64 (n_groups * n_pairs), written
 * (accumulate = 0) or added to (accumulate = 1, staged through the workspace). */
LTB_API size_t ltb200_group_masks_walk_workspace(int64_t n_frames, int n_groups, int n_pairs,
                                                 int accumulate);
LTB_API int ltb200_group_masks_walk(const float* tile, int64_t n_frames, int64_t sig_size,
                                    int64_t ld_tile, const uint32_t* boxes, const uint32_t* ops0,
                                    const uint32_t* ops1, const uint32_t* ops2,
                                    const uint32_t* ops3, const uint32_t* events0,
                                    const uint32_t* events1, const float* table0,
                                    const float* table1, const float* table2,
                                    const float* table3, const int32_t* seg_off_host,
                                    int n_segments, int n_groups, int n_pairs, float* out,
                                    int64_t ld_out, int accumulate, void* workspace,
                                    size_t workspace_bytes, void* stream);

/* Mirror-symmetric plan (opt-in: validated on B200, 3 % faster than the banded plan on cfg4 --
 * the kernel is gather-bound, not bound by what the symmetry saves): for stacks whose masks obey m(sy - y, x) = conj(m(y, x)) -- the
 * radial_mask_factory stacks about the default centre -- a stage holds 32 ORBITS: 8 quads of
 * rows y < sy/2 followed by the 8 mirrored quads (same columns, row sy - y).  The kernel forms
 * I(p) + I(p') and I(p) - I(p'); the sums meet the real parts of the weights, the differences
 * the imaginary parts: two MMAs of 64 columns per pixel PAIR instead of two of 112 per pixel,
 * and a weight table of 128 rows per orbit instead of 112 per pixel.  table_sym is
 * (128, n_entries / 2) float32: rows [0, 32) hi(Re w_c), [32, 64) lo(Re w_c), [64, 96)
 * hi(Im w_c), [96, 128) lo(Im w_c) for complex column c, w = the orbit-averaged weight of the
 * upper pixel.  Same group / band / workspace conventions as ltb200_group_masks_tc_banded; the
 * rows without a mirror partner (0 and sy/2) go through that entry point with accumulate = 1. */
LTB_API int ltb200_group_masks_tc_sym(const float* tile, int64_t n_frames, int64_t sig_size,
                                      int64_t ld_tile, const int32_t* entry_px,
                                      const float* table_sym, const int32_t* group_off_host,
                                      const int32_t* group_off_dev, int n_groups, int n_pairs,
                                      int n_bands, float* out, int64_t ld_out, int accumulate,
                                      int chain, void* workspace, size_t workspace_bytes,
                                      void* stream);

/* ---------------------------------------------------------------------------------------
 * Synthetic data (twin of oracle/synth.py): fills dst[0..count) with value(start + i).
 *   LTB_F32: uniform [0,1) (24-bit);  LTB_U16: Poisson(3) counts.
 * ------------------------------------------------------------------------------------- */
LTB_API int ltb200_synth_fill(void* dst, int dtype, int64_t start, int64_t count, uint32_t seed,
                      void* stream);

/* ---------------------------------------------------------------------------------------
 * Nav-space post-processing of the centre-of-mass moments (K9) -- CoMUDF.get_results
 * (udf/com.py:650-717) on the device.  raw: (rows, ld_raw >= 3) float32 [m00, m10, m01] per
 * scan position; row_of_nav (nullable, ny*nx int32): row of raw for every scan position, -1
 * outside the roi (outputs NaN there); valid (nullable, ny*nx uint8): positions that enter the
 * regression (default: row >= 0).  center_shifts (com.py:100-107) in float32, then the 2x2
 * float64 `transform` (apply_correction, com.py:110-127), the regression (mode -1 none, 0
 * subtract the mean, 1 subtract the least-squares plane c0 + c1 y + c2 x -- both over the valid
 * positions --, 2 subtract the given plane; `regression` is the (3, 2) float64 device array,
 * read in mode 2 and written otherwise; com.py:600-648), magnitude / divergence / curl with
 * np.gradient stencils (com.py:130-142) in float64 -- the reference's result arrays for these are
 * float64 too.  Outputs cover the full scan grid: raw_shifts / raw_com (float32) and field
 * (float64) are (ny*nx, 2) as (y, x), the others (ny*nx) float64.
 * ny, nx >= 2 (np.gradient).  Sums are reduced in a fixed order (deterministic).
 * ------------------------------------------------------------------------------------- */
LTB_API size_t ltb200_com_workspace(int ny, int nx);
LTB_API int ltb200_com_postprocess(const float* raw, int64_t ld_raw, const int32_t* row_of_nav,
                                   const uint8_t* valid, int ny, int nx, double cy, double cx,
                                   const double* transform /* host, 4 */, int regression_mode,
                                   double* regression /* device, 6 */, float* raw_shifts,
                                   float* raw_com, double* field, double* field_y,
                                   double* field_x, double* magnitude, double* divergence,
                                   double* curl,
                                   void* workspace, size_t workspace_bytes, void* stream);
/* guess_corrections (udf/com.py:145-295) without 720 passes: the curl of a linearly transformed
 * field is linear in the gradient fields g = (dy/d0, dy/d1, dx/d0, dx/d1) of the (ny, nx) float32
 * maps y_centers / x_centers, so its RMS for any 2x2 matrix is a quadratic form in their Gram
 * matrix.  sums17 (device, float64): the 10 upper-triangular Gram entries (row-major), the 4
 * sums of g, sum(y), sum(x) and the count, over the window rows [r0, r1) x columns [c0, c1)
 * (gradients are taken on the full grid, like np.gradient before slicing).
 * workspace: >= 296 * 17 * 8 bytes. */
LTB_API int ltb200_com_gradient_gram(const float* y_centers, const float* x_centers, int ny,
                                     int nx, int r0, int r1, int c0, int c1, double* sums17,
                                     void* workspace, size_t workspace_bytes, void* stream);
/* divergence of transform . (y, x) over the same window: pass 0 -> per-block (min, max) pairs in
 * minmax_blocks (device, 2 * 296 float64; *n_blocks pairs are valid), pass 1 -> counts of the 5
 * equal bins over [-range, range] (np.histogram(bins=5, range=...)) added to hist5 (device). */
LTB_API int ltb200_com_divergence_stats(const float* y_centers, const float* x_centers, int ny,
                                        int nx, int r0, int r1, int c0, int c1,
                                        const double* transform /* host, 4 */, int pass,
                                        double range, double* minmax_blocks, int* n_blocks,
                                        unsigned long long* hist5, void* stream);

/* ---------------------------------------------------------------------------------------
 * Read-only HBM streaming probe (measurement only, not on the product path): reads `bytes`
 * bytes of `buf` once with the grid shape of the masked-reduction kernels (one persistent CTA
 * per SM) and does nothing with them.  mode 0: bulk TMA copies into a shared-memory ring
 * (the ingest half of K6 / K8, no math); mode 1: LDG.128 into registers (`sink`: 4-byte device
 * scratch).  bench.py times it to report `roofline.peak_read_only` next to the copy-benchmark
 * peak of MEASURED_PEAKS.json.
 * ------------------------------------------------------------------------------------- */
LTB_API int ltb200_probe_read(const void* buf, size_t bytes, int mode, void* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LTB200_H */
