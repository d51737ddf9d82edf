"""CPU restatement of the reference UDF hot path (ApplyMasks / CoM / Sum / SumSig).

Test infrastructure (see oracle/__init__.py): the checker, never the product.
All ``file:line`` citations are relative to /root/reference/.

Pinned against the unmodified reference by tests/golden/*.npz
(tests/test_oracle_golden.py).
"""
import ctypes
import os

import numpy as np

from . import masks_gen

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------
# partitioning / tiling  (src/libertem/io/dataset/base/partition.py:66-99,
#                         src/libertem/io/dataset/base/tiling.py:205-243)
# --------------------------------------------------------------------------------------

def partition_boundaries(num_frames, num_partitions):
    """``np.linspace(0, N, P+1, dtype=int)`` -- partition.py:72-88."""
    num_partitions = min(num_partitions, num_frames)
    b = np.linspace(0, num_frames, num=max(2, num_partitions + 1), endpoint=True, dtype=int)
    return [(int(a), int(c)) for a, c in zip(b[:-1], b[1:])]


def iter_tiles(part_start, part_stop, sig_shape, tileshape=None):
    """Yield ``(f0, f1, sig_slices)`` in the reference's order: depth blocks outer,
    sig slices inner (tiling.py:224-240 depth loop, :87-131 slice loop).  ``tileshape=None``
    is the float32 in-memory case: one tile = whole partition, full frames."""
    if tileshape is None:
        yield part_start, part_stop, tuple(slice(0, s) for s in sig_shape)
        return
    depth = tileshape[0]
    tsig = tuple(tileshape[1:])
    sig_slices = []
    # TilingScheme.make_for_shape: slices over the sig dims, row-major (tiling_scheme.py:60-117)
    ranges = [range(0, s, t) for s, t in zip(sig_shape, tsig)]
    for idx in np.ndindex(*[len(r) for r in ranges]):
        sl = tuple(
            slice(r[i], min(r[i] + t, s)) for r, i, t, s in zip(ranges, idx, tsig, sig_shape)
        )
        sig_slices.append(sl)
    for f0 in range(part_start, part_stop, depth):
        f1 = min(f0 + depth, part_stop)
        for sl in sig_slices:
            yield f0, f1, sl


# --------------------------------------------------------------------------------------
# dtype rules
# --------------------------------------------------------------------------------------

def input_dtype(dataset_dtype, preferred=(np.float32,)):
    """udf/base.py:106-123 ``_get_dtype``: result_type over each UDF's preferred dtype."""
    t = np.dtype(dataset_dtype)
    for p in preferred:
        t = np.result_type(p, t)
    return t


# --------------------------------------------------------------------------------------
# mask stack  (src/libertem/common/container.py:260-314, 74-94)
# --------------------------------------------------------------------------------------

def compute_mask_stack(mask_factories):
    """Call the factories and ``np.concatenate`` to ``(M, sy, sx)`` -- container.py:260-314
    (dense branch; sparse masks are densified here, the CSR view is built per sig slice)."""
    if callable(mask_factories):
        raw = mask_factories()
        raw = raw.toarray() if hasattr(raw, 'toarray') else np.asarray(raw)
        return raw
    slices = []
    for f in mask_factories:
        m = f()
        m = m.toarray() if hasattr(m, 'toarray') else np.asarray(m)
        slices.append(m.reshape((1,) + m.shape))
    return np.concatenate(slices)


def masks_for_sig_slice(stack, sig_slice, dtype):
    """``m[sig].reshape(M, -1).T.astype(dtype)`` -- container.py:74-94 (dense slicer).
    Returns shape ``(K_tile, M)`` (F-ordered, like the reference)."""
    m = stack[(slice(None),) + tuple(sig_slice)]
    m = m.reshape((stack.shape[0], -1)).T
    return m.astype(dtype)


# --------------------------------------------------------------------------------------
# dense-sparse product  (src/libertem/common/numba/__init__.py:90-184)
# --------------------------------------------------------------------------------------

_lib = None


def _oracle_lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, '_build', 'liboracle.so')
        if not os.path.exists(path):
            raise RuntimeError(
                f'{path} missing: run `make -C oracle` (or __graft_entry__.build())'
            )
        _lib = ctypes.CDLL(path)
    return _lib


def rmatmul_csr(left_dense, data, indices, indptr, n_cols):
    """``_rmatmul_csr`` loop order (numba/__init__.py:169-184): for each sparse row k
    (ascending), for each nnz (k, m, v): ``res_t[m, f] += left[f, k] * v`` for all f.
    float32 product and float32 accumulation; returns ``res_t.T.copy()``."""
    left_dense = np.ascontiguousarray(left_dense)
    F, K = left_dense.shape
    dt = np.result_type(left_dense.dtype, data.dtype)
    if dt == np.float32 and left_dense.dtype == np.float32:
        lib = _oracle_lib()
        res_t = np.zeros((n_cols, F), dtype=np.float32)
        data = np.ascontiguousarray(data, dtype=np.float32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        lib.oracle_rmatmul_csr_f32(
            left_dense.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(F), ctypes.c_int64(K),
            data.ctypes.data_as(ctypes.c_void_p), indices.ctypes.data_as(ctypes.c_void_p),
            indptr.ctypes.data_as(ctypes.c_void_p), res_t.ctypes.data_as(ctypes.c_void_p),
        )
        return res_t.T.copy()
    if dt == np.complex64 and left_dense.dtype == np.float32:
        lib = _oracle_lib()
        res_t = np.zeros((n_cols, F), dtype=np.complex64)
        data = np.ascontiguousarray(data, dtype=np.complex64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        lib.oracle_rmatmul_csr_c64(
            left_dense.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(F), ctypes.c_int64(K),
            data.ctypes.data_as(ctypes.c_void_p), indices.ctypes.data_as(ctypes.c_void_p),
            indptr.ctypes.data_as(ctypes.c_void_p), res_t.ctypes.data_as(ctypes.c_void_p),
        )
        return res_t.T.copy()
    # generic (slow) fallback, same loop order
    res_t = np.zeros((n_cols, F), dtype=dt)
    for k in range(len(indptr) - 1):
        for i in range(indptr[k], indptr[k + 1]):
            res_t[indices[i]] += (left_dense[:, k] * data[i]).astype(dt)
    return res_t.T.copy()


def rmatmul(left_dense, right_sparse):
    """numba/__init__.py:90-151 dispatch for a scipy CSR/CSC right-hand side."""
    import scipy.sparse as sp
    if left_dense.shape[1] != right_sparse.shape[0]:
        raise ValueError('Shape mismatch')
    if isinstance(right_sparse, sp.csr_matrix):
        return rmatmul_csr(left_dense, right_sparse.data, right_sparse.indices,
                           right_sparse.indptr, right_sparse.shape[1])
    if isinstance(right_sparse, sp.csc_matrix):
        # _rmatmul_csc (:153-166): for each column m, each nnz (k, v): res_t[m, f] += left[f,k]*v
        dt = np.result_type(left_dense.dtype, right_sparse.dtype)
        res_t = np.zeros((right_sparse.shape[1], left_dense.shape[0]), dtype=dt)
        ip, ix, dv = right_sparse.indptr, right_sparse.indices, right_sparse.data
        for m in range(right_sparse.shape[1]):
            for i in range(ip[m], ip[m + 1]):
                res_t[m] += (left_dense[:, ix[i]] * dv[i]).astype(dt)
        return res_t.T.copy()
    raise ValueError(type(right_sparse))


# --------------------------------------------------------------------------------------
# ApplyMasksUDF  (src/libertem/udf/masks.py:12-125, 353-392)
# --------------------------------------------------------------------------------------

def process_flat(flat_tile, masks_t, use_torch=True):
    """``flat_tile @ masks`` with masks ``(K, M)`` -- masks.py:59-66 (torch.mm when the
    reference would pick torch: float input, same dtype, dense) else :76-77 (numpy ``@``)."""
    if (use_torch and flat_tile.dtype.kind == 'f' and flat_tile.dtype == masks_t.dtype):
        try:
            import torch
            return torch.mm(torch.from_numpy(np.ascontiguousarray(flat_tile)),
                            torch.from_numpy(masks_t)).numpy()
        except ImportError:  # pragma: no cover
            pass
    return flat_tile @ masks_t


def apply_masks(data, mask_stack, sig_dims=2, num_partitions=1, tileshape=None,
                mask_dtype=None, use_sparse=False, preferred_dtype=np.float32,
                use_torch=True, roi=None):
    """ApplyMasksUDF over a whole dataset through the reference's partition/tile loops.

    data: array ``nav + sig``; mask_stack: ``(M,) + sig``.  Returns the flat
    ``(n_frames_in_roi, M)`` 'intensity' buffer (udf/masks.py:360-392); result dtype
    ``result_type(input_dtype, mask_dtype)`` (:362).
    """
    import scipy.sparse as sp
    sig_shape = data.shape[-sig_dims:]
    flat = data.reshape((-1,) + tuple(sig_shape))
    n = flat.shape[0]
    in_dtype = input_dtype(data.dtype, (preferred_dtype,))
    if mask_dtype is None:
        mask_dtype = mask_stack.dtype
    res_dtype = np.result_type(in_dtype, mask_dtype)
    M = mask_stack.shape[0]
    out = np.zeros((n, M), dtype=res_dtype)
    keep = np.ones(n, dtype=bool) if roi is None else np.asarray(roi).reshape(-1).astype(bool)
    for p0, p1 in partition_boundaries(n, num_partitions):
        for f0, f1, sl in iter_tiles(p0, p1, sig_shape, tileshape):
            sel = np.arange(f0, f1)[keep[f0:f1]]
            if len(sel) == 0:
                continue
            tile = flat[(sel,) + tuple(sl)].astype(in_dtype)
            flat_tile = np.ascontiguousarray(tile.reshape((tile.shape[0], -1)))
            if use_sparse:
                m = mask_stack[(slice(None),) + tuple(sl)].reshape((M, -1)).T
                # container.py:53-64: CSR from COO coords, canonical format
                csr = sp.csr_matrix(m.astype(mask_dtype))
                part = rmatmul(flat_tile, csr)
            else:
                part = process_flat(flat_tile, masks_for_sig_slice(mask_stack, sl, mask_dtype),
                                    use_torch=use_torch)
            out[sel] += part
    return out[keep]


# --------------------------------------------------------------------------------------
# SumUDF / SumSigUDF  (src/libertem/udf/sum.py:6-58, src/libertem/udf/sumsigudf.py:6-39)
# --------------------------------------------------------------------------------------

def sum_udf(data, sig_dims=2, num_partitions=1, tileshape=None, dtype=np.float32, roi=None):
    """per-partition ``intensity += np.sum(tile, axis=0)`` then ``merge: dest += src``."""
    sig_shape = data.shape[-sig_dims:]
    flat = data.reshape((-1,) + tuple(sig_shape))
    n = flat.shape[0]
    in_dtype = input_dtype(data.dtype, (dtype,))
    keep = np.ones(n, dtype=bool) if roi is None else np.asarray(roi).reshape(-1).astype(bool)
    dest = np.zeros(sig_shape, dtype=in_dtype)
    for p0, p1 in partition_boundaries(n, num_partitions):
        part = np.zeros(sig_shape, dtype=in_dtype)
        for f0, f1, sl in iter_tiles(p0, p1, sig_shape, tileshape):
            sel = np.arange(f0, f1)[keep[f0:f1]]
            if len(sel) == 0:
                continue
            tile = flat[(sel,) + tuple(sl)].astype(in_dtype)
            part[tuple(sl)] += np.sum(tile, axis=0)
        dest += part
    return dest


def sumsig_udf(data, sig_dims=2, num_partitions=1, tileshape=None, roi=None):
    """``intensity[f] += np.sum(tile.reshape(n, -1), axis=1)``; dtype result_type(input, f32)."""
    sig_shape = data.shape[-sig_dims:]
    flat = data.reshape((-1,) + tuple(sig_shape))
    n = flat.shape[0]
    in_dtype = input_dtype(data.dtype, (np.float32,))
    res_dtype = np.result_type(in_dtype, np.float32)
    keep = np.ones(n, dtype=bool) if roi is None else np.asarray(roi).reshape(-1).astype(bool)
    out = np.zeros(n, dtype=res_dtype)
    for p0, p1 in partition_boundaries(n, num_partitions):
        for f0, f1, sl in iter_tiles(p0, p1, sig_shape, tileshape):
            sel = np.arange(f0, f1)[keep[f0:f1]]
            if len(sel) == 0:
                continue
            tile = flat[(sel,) + tuple(sl)].astype(in_dtype)
            out[sel] += np.sum(tile.reshape((tile.shape[0], -1)), axis=1)
    return out[keep]


# --------------------------------------------------------------------------------------
# CoMUDF  (src/libertem/udf/com.py)
# --------------------------------------------------------------------------------------

def com_mask_stack(sig_shape, cy, cx, r=float('inf'), ri=0.):
    """[D, gradient_y*D, gradient_x*D] as float32 -- com.py:47-97, 534-575."""
    sy, sx = sig_shape
    if ri is None or np.isclose(ri, 0.):
        base = masks_gen.circular(centerX=cx, centerY=cy, imageSizeX=sx, imageSizeY=sy, radius=r)
    else:
        base = masks_gen.ring(centerX=cx, centerY=cy, imageSizeX=sx, imageSizeY=sy,
                              radius=r, radius_inner=ri)
    gy = masks_gen.gradient_y(imageSizeX=sx, imageSizeY=sy) * base
    gx = masks_gen.gradient_x(imageSizeX=sx, imageSizeY=sy) * base
    return np.stack([np.asarray(base), gy, gx]).astype(np.float32)


def center_shifts(img_sum, img_y, img_x, ref_y, ref_x):
    """com.py:100-107."""
    x_centers = np.divide(img_x, img_sum, where=img_sum != 0)
    y_centers = np.divide(img_y, img_sum, where=img_sum != 0)
    x_centers[img_sum == 0] = ref_x
    y_centers[img_sum == 0] = ref_y
    x_centers -= ref_x
    y_centers -= ref_y
    return (y_centers, x_centers)


def rotate_deg(degrees):
    """corrections/coordinates.py:11-27."""
    radians = np.pi * degrees / 180
    return np.array([
        (np.cos(radians), np.sin(radians)),
        (-np.sin(radians), np.cos(radians)),
    ])


def flip_y_matrix():
    """corrections/coordinates.py:30-37."""
    return np.array([(-1, 0), (0, 1)])


def apply_correction(y_centers, x_centers, scan_rotation, flip_y, forward=True):
    """com.py:110-127."""
    shape = y_centers.shape
    transform = flip_y_matrix() if flip_y else np.eye(2)
    transform = rotate_deg(scan_rotation) @ transform
    y_centers = y_centers.reshape(-1)
    x_centers = x_centers.reshape(-1)
    if not forward:
        transform = np.linalg.inv(transform)
    y_t, x_t = transform @ (y_centers, x_centers)
    return (y_t.reshape(shape), x_t.reshape(shape))


def divergence(y_centers, x_centers):
    """com.py:130-131."""
    return np.gradient(y_centers, axis=0) + np.gradient(x_centers, axis=1)


def curl_2d(y_centers, x_centers):
    """com.py:134-138."""
    return np.gradient(y_centers, axis=1) - np.gradient(x_centers, axis=0)


def magnitude(y_centers, x_centers):
    """com.py:141-142."""
    return np.sqrt(y_centers ** 2 + x_centers ** 2)


def com_get_results(raw_mask_result, nav_shape, cy, cx, scan_rotation=0., flip_y=False,
                    regression=-1, roi=None):
    """CoMUDF.get_results (com.py:650-717) + get_regression (:600-648).

    raw_mask_result: flat ``(n_roi_frames, 3)``; returns dict of flat nav buffers
    (compressed by roi like the reference's BufferWrapper.raw_data) + 'regression'.
    """
    nav_shape = tuple(nav_shape)
    n = int(np.prod(nav_shape))
    dtype = raw_mask_result.dtype
    if roi is None:
        full = raw_mask_result.reshape(nav_shape + (3,))
        valid_mask = np.ones(nav_shape, dtype=bool)
    else:
        roi = np.asarray(roi).reshape(nav_shape).astype(bool)
        full = np.full(nav_shape + (3,), np.nan, dtype=dtype)  # buffers.py:470-505
        full[roi] = raw_mask_result
        valid_mask = roi
    raw_shifts = center_shifts(img_sum=full[..., 0], img_y=full[..., 1], img_x=full[..., 2],
                               ref_y=cy, ref_x=cx)
    raw_com = (raw_shifts[0].copy() + cy, raw_shifts[1].copy() + cx)
    field = apply_correction(raw_shifts[0], raw_shifts[1], scan_rotation, flip_y)
    raw_shifts = np.moveaxis(np.array(raw_shifts), 0, -1)
    raw_com = np.moveaxis(np.array(raw_com), 0, -1)
    field = np.moveaxis(np.array(field), 0, -1)

    reg = np.zeros((3, 2))
    inp = None

    def get_inp():
        a = np.ones(field.shape[:-1] + (3,))
        y, x = np.ogrid[:field.shape[0], :field.shape[1]]
        a[..., 1] = y
        a[..., 2] = x
        return a

    if isinstance(regression, (int, np.integer)):
        if regression == 0:
            reg[0] = np.mean(field[valid_mask], axis=0)
        elif regression == 1:
            inp = get_inp()
            reg[:] = np.linalg.lstsq(inp[valid_mask], field[valid_mask], rcond=None)[0]
        elif regression != -1:
            raise ValueError(regression)
    else:
        reg[:] = np.array(regression)
    has_lin = not np.allclose(reg[1:], 0)
    if has_lin and inp is None:
        inp = get_inp()
    if not has_lin:
        inp = None
    if inp is not None:
        field[valid_mask] -= inp[valid_mask] @ reg
    elif not np.allclose(reg[0], 0):
        field[valid_mask] -= reg[0]

    results = {
        'raw_shifts': raw_shifts, 'raw_com': raw_com, 'field': field,
        'field_y': field[..., 0], 'field_x': field[..., 1],
        'magnitude': magnitude(field[..., 0], field[..., 1]),
        'divergence': divergence(field[..., 0], field[..., 1]),
        'curl': curl_2d(field[..., 0], field[..., 1]),
    }
    out = {}
    for k, v in results.items():
        # com.py:707-716: roi-compress or flatten; dtype is whatever numpy produced
        # (float64 after the float64 2x2 transform), only the dtype *kind* is checked
        out[k] = v[roi] if roi is not None else v.reshape((n, -1))
    out['regression'] = reg.astype(np.float64)
    return out


def com_udf(data, cy=None, cx=None, r=float('inf'), ri=0., scan_rotation=0., flip_y=False,
            regression=-1, num_partitions=1, tileshape=None, roi=None):
    """Full CoMUDF: 3-mask engine pass (com.py:577-582) + get_results."""
    sy, sx = data.shape[-2:]
    nav_shape = data.shape[:-2]
    if cy is None:
        cy = sy // 2          # com.py:512-520
    if cx is None:
        cx = sx // 2
    stack = com_mask_stack((sy, sx), cy, cx, r, ri)
    raw = apply_masks(data, stack, num_partitions=num_partitions, tileshape=tileshape,
                      mask_dtype=np.float32, use_torch=True, roi=roi)
    res = com_get_results(raw, nav_shape, cy, cx, scan_rotation, flip_y, regression, roi)
    res['raw_mask_result'] = raw
    return res


# --------------------------------------------------------------------------------------
# RadialFourierAnalysis  (src/libertem/analysis/radialfourier.py:106-146, 184-194, 316-354)
# --------------------------------------------------------------------------------------

def radial_fourier_params(sig_shape, cx=None, cy=None, ri=0, ro=None, n_bins=1, max_order=24):
    """get_parameters defaults -- radialfourier.py:316-354."""
    sy, sx = sig_shape
    cx = sx / 2 if cx is None else cx
    cy = sy / 2 if cy is None else cy
    if ro is None:
        ro = masks_gen.bounding_radius(cx, cy, sx, sy)
    return dict(cx=cx, cy=cy, ri=ri, ro=ro, n_bins=n_bins, max_order=max_order)


def radial_mask_stack(sig_shape, cx, cy, ri, ro, n_bins, max_order):
    """``ring_b(r) * exp(i*o*phi)`` as complex64 ``(n_bins*(max_order+1), sy, sx)``, bin-major
    -- radialfourier.py:106-146 (dense branch; identical values to the sparse branch)."""
    sy, sx = sig_shape
    dtype = np.complex64
    rings = masks_gen.radial_bins(centerX=cx, centerY=cy, imageSizeX=sx, imageSizeY=sy,
                                  radius=ro, radius_inner=ri, n_bins=n_bins, dtype=dtype)
    orders = np.arange(max_order + 1, dtype=dtype)
    _, phi = masks_gen.polar_map(centerX=cx, centerY=cy, imageSizeX=sx, imageSizeY=sy)
    modulator = np.exp(phi.astype(dtype) * orders[:, np.newaxis, np.newaxis] * 1j)
    ring_stack = rings[:, np.newaxis, ...] * modulator
    return ring_stack.reshape((-1, sy, sx))


def radial_fourier(data, num_partitions=1, use_sparse=True, **params):
    """raw_results ``(n_bins, max_order+1, *nav)`` complex64 -- radialfourier.py:184-194."""
    sig_shape = data.shape[-2:]
    nav_shape = data.shape[:-2]
    p = radial_fourier_params(sig_shape, **params)
    stack = radial_mask_stack(sig_shape, **p)
    if use_sparse:
        import scipy.sparse as sp
        flat = data.reshape((-1, sig_shape[0] * sig_shape[1])).astype(np.float32)
        csr = sp.csr_matrix(stack.reshape((stack.shape[0], -1)).T)
        inten = np.zeros((flat.shape[0], stack.shape[0]), dtype=np.complex64)
        for p0, p1 in partition_boundaries(flat.shape[0], num_partitions):
            inten[p0:p1] = rmatmul(flat[p0:p1], csr)
    else:
        inten = apply_masks(data, stack, num_partitions=num_partitions,
                            mask_dtype=np.complex64, use_torch=False)
    res = inten.reshape((int(np.prod(nav_shape)), -1)).T
    return res.reshape((p['n_bins'], p['max_order'] + 1) + tuple(nav_shape))
