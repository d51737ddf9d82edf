"""Tensor-level wrappers over the C ABI (device pointers come from torch CUDA tensors).

This is the seam of the reference's ``ApplyMasksEngine.process_flat``
(src/libertem/udf/masks.py:31-83): two arrays in, ``(frames, masks)`` array out.
"""
import numpy as np
import torch

from . import _lib
from ._lib import get_lib, check

_TORCH_DTYPES = {
    torch.float32: _lib.LTB_F32, torch.uint16: _lib.LTB_U16, torch.uint8: _lib.LTB_U8,
    torch.int16: _lib.LTB_I16, torch.float64: _lib.LTB_F64, torch.int32: _lib.LTB_I32,
    torch.uint32: _lib.LTB_U32, torch.int64: _lib.LTB_I64, torch.uint64: _lib.LTB_U64,
    torch.int8: _lib.LTB_I8,
}

_workspaces = {}

#: when set to a list, masks_dense appends (start_event, end_event, n_frames, sig_size,
#: itemsize) around its launches (bench.py uses this to time the dominant kernel live)
EVENT_LOG = None


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _workspace(device, nbytes):
    ws = _workspaces.get(device)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[device] = ws
    return ws


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.LTB200Error(f'{name} must be a CUDA tensor (no CPU fallback)')


def masks_dense(tile, masks, out=None, accumulate=False, sig_sum=None):
    """``out[f, m] (+)= sum_k tile[f, k] * masks[m, k]`` on the GPU.

    tile: (F, K) CUDA tensor (row stride arbitrary, unit inner stride); masks: (M, K) float32
    (or float64 -> float64 path).  Returns ``out`` (F, M).
    """
    lib = get_lib()
    _require_cuda(tile, 'tile')
    _require_cuda(masks, 'masks')
    if tile.dim() != 2 or masks.dim() != 2 or tile.shape[1] != masks.shape[1]:
        raise ValueError(f'shape mismatch: tile {tuple(tile.shape)} masks {tuple(masks.shape)}')
    if tile.dtype not in _TORCH_DTYPES:
        raise TypeError(f'unsupported tile dtype {tile.dtype}')
    if tile.shape[1] > 0 and tile.stride(1) != 1:
        tile = tile.contiguous()
    if masks.shape[1] > 0 and masks.stride(1) != 1:
        masks = masks.contiguous()
    F, K = tile.shape
    M = masks.shape[0]
    f64 = masks.dtype == torch.float64
    res_dtype = torch.float64 if f64 else torch.float32
    if not f64 and masks.dtype != torch.float32:
        raise TypeError(f'masks must be float32 or float64, got {masks.dtype}')
    if out is None:
        out = torch.zeros((F, M), dtype=res_dtype, device=tile.device)
        accumulate = False
    else:
        _require_cuda(out, 'out')
        if out.shape != (F, M) or out.dtype != res_dtype or (M > 1 and out.stride(1) != 1):
            raise ValueError('out must be (F, M) of the result dtype with unit inner stride')
    ld_tile = tile.stride(0) if F > 1 else max(K, 1)
    ld_masks = masks.stride(0) if M > 1 else max(K, 1)
    ld_out = out.stride(0) if F > 1 else max(M, 1)
    with torch.cuda.device(tile.device):
        st = _stream_ptr(tile.device)
        if f64:
            if sig_sum is not None:
                raise _lib.LTB200Error('sig_sum is only fused on the float32 path')
            check(lib.ltb200_masks_dense_f64(
                tile.data_ptr(), _TORCH_DTYPES[tile.dtype], F, K, ld_tile, masks.data_ptr(), M,
                ld_masks, out.data_ptr(), ld_out, int(bool(accumulate)), st))
            return out
        need = lib.ltb200_masks_dense_workspace(F, K, M, int(sig_sum is not None))
        ws = _workspace(tile.device, need) if need else None
        if sig_sum is not None:
            _require_cuda(sig_sum, 'sig_sum')
            if sig_sum.dtype != torch.float32 or sig_sum.numel() != K or not sig_sum.is_contiguous():
                raise ValueError('sig_sum must be a contiguous float32 tensor of sig_size')
        log = EVENT_LOG
        if log is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(tile.device))
        check(lib.ltb200_masks_dense(
            tile.data_ptr(), _TORCH_DTYPES[tile.dtype], F, K, ld_tile,
            masks.data_ptr(), M, ld_masks, out.data_ptr(), ld_out, int(bool(accumulate)),
            sig_sum.data_ptr() if sig_sum is not None else None,
            ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0, st))
        if log is not None:
            e1.record(torch.cuda.current_stream(tile.device))
            log.append((e0, e1, F, K, tile.element_size()))
    return out


def masks_dense_tc(tile, masks, out=None, accumulate=False, chain=0):
    """The dense contraction on the tensor cores (K6, split-TF32 tcgen05): float32 tile (F, K)
    and masks (M, K) -> out (F, M).  Explicit form of what ``masks_dense`` selects for wide
    stacks; raises for shapes the kernel does not take."""
    lib = get_lib()
    _require_cuda(tile, 'tile')
    _require_cuda(masks, 'masks')
    if tile.dtype != torch.float32 or masks.dtype != torch.float32:
        raise TypeError('masks_dense_tc takes float32 tiles and masks')
    if tile.dim() != 2 or masks.dim() != 2 or tile.shape[1] != masks.shape[1]:
        raise ValueError(f'shape mismatch: tile {tuple(tile.shape)} masks {tuple(masks.shape)}')
    if tile.stride(1) != 1:
        tile = tile.contiguous()
    if masks.stride(1) != 1:
        masks = masks.contiguous()
    F, K = tile.shape
    M = masks.shape[0]
    if out is None:
        out = torch.zeros((F, M), dtype=torch.float32, device=tile.device)
        accumulate = False
    ld_tile = tile.stride(0) if F > 1 else max(K, 1)
    ld_masks = masks.stride(0) if M > 1 else max(K, 1)
    ld_out = out.stride(0) if F > 1 else max(M, 1)
    with torch.cuda.device(tile.device):
        need = lib.ltb200_masks_dense_tc_workspace(F, K, M)
        ws = _workspace(tile.device, need)
        check(lib.ltb200_masks_dense_tc(
            tile.data_ptr(), F, K, ld_tile, masks.data_ptr(), M, ld_masks, out.data_ptr(),
            ld_out, int(bool(accumulate)), int(chain), ws.data_ptr(), ws.numel(),
            _stream_ptr(tile.device)))
    return out


def masks_dense_tc_u16(tile, masks, out=None, accumulate=False, chain=0, sig_sum=None):
    """uint16 form of ``masks_dense_tc`` (1..16 columns) with the optional fused frame sum:
    ``sig_sum`` (K,) float32 += sum over the frames of the tile (SumUDF)."""
    lib = get_lib()
    _require_cuda(tile, 'tile')
    _require_cuda(masks, 'masks')
    if tile.dtype != torch.uint16 or masks.dtype != torch.float32:
        raise TypeError('masks_dense_tc_u16 takes uint16 tiles and float32 masks')
    if tile.dim() != 2 or masks.dim() != 2 or tile.shape[1] != masks.shape[1]:
        raise ValueError(f'shape mismatch: tile {tuple(tile.shape)} masks {tuple(masks.shape)}')
    if tile.stride(1) != 1:
        tile = tile.contiguous()
    if masks.stride(1) != 1:
        masks = masks.contiguous()
    F, K = tile.shape
    M = masks.shape[0]
    if out is None:
        out = torch.zeros((F, M), dtype=torch.float32, device=tile.device)
        accumulate = False
    if sig_sum is not None:
        _require_cuda(sig_sum, 'sig_sum')
        if sig_sum.dtype != torch.float32 or sig_sum.numel() != K or not sig_sum.is_contiguous():
            raise ValueError('sig_sum must be a contiguous float32 tensor of sig_size elements')
    ld_tile = tile.stride(0) if F > 1 else max(K, 1)
    ld_masks = masks.stride(0) if M > 1 else max(K, 1)
    ld_out = out.stride(0) if F > 1 else max(M, 1)
    with torch.cuda.device(tile.device):
        need = lib.ltb200_masks_dense_tc_u16_workspace(F, K, M, int(sig_sum is not None))
        ws = _workspace(tile.device, need)
        check(lib.ltb200_masks_dense_tc_u16(
            tile.data_ptr(), F, K, ld_tile, masks.data_ptr(), M, ld_masks, out.data_ptr(),
            ld_out, int(bool(accumulate)), int(chain),
            sig_sum.data_ptr() if sig_sum is not None else None, ws.data_ptr(), ws.numel(),
            _stream_ptr(tile.device)))
    return out


def masks_dense_i8(tile, masks, out=None, accumulate=False, sig_sum=None):
    """Integer fast path (K8, int8 tensor cores): uint16 / uint8 tile (F, K) x int8 masks (M, K),
    M <= 16 -> float32 out (F, M) of the exact integer sums; ``sig_sum`` (K,) float32 +=
    the exact frame sum of the tile (SumUDF)."""
    lib = get_lib()
    _require_cuda(tile, 'tile')
    _require_cuda(masks, 'masks')
    if tile.dtype not in (torch.uint16, torch.uint8) or masks.dtype != torch.int8:
        raise TypeError('masks_dense_i8 takes uint16 / uint8 tiles and int8 masks')
    tdt = _TORCH_DTYPES[tile.dtype]
    if tile.dim() != 2 or masks.dim() != 2 or tile.shape[1] != masks.shape[1]:
        raise ValueError(f'shape mismatch: tile {tuple(tile.shape)} masks {tuple(masks.shape)}')
    if tile.stride(1) != 1:
        tile = tile.contiguous()
    if masks.stride(1) != 1:
        masks = masks.contiguous()
    F, K = tile.shape
    M = masks.shape[0]
    if out is None:
        out = torch.zeros((F, M), dtype=torch.float32, device=tile.device)
        accumulate = False
    if sig_sum is not None:
        _require_cuda(sig_sum, 'sig_sum')
        if sig_sum.dtype != torch.float32 or sig_sum.numel() != K or not sig_sum.is_contiguous():
            raise ValueError('sig_sum must be a contiguous float32 tensor of sig_size elements')
    ld_tile = tile.stride(0) if F > 1 else max(K, 1)
    ld_masks = masks.stride(0) if M > 1 else max(K, 1)
    ld_out = out.stride(0) if F > 1 else max(M, 1)
    with torch.cuda.device(tile.device):
        need = lib.ltb200_masks_dense_i8_workspace(tdt, F, K, M, int(sig_sum is not None))
        ws = _workspace(tile.device, max(need, 256))
        check(lib.ltb200_masks_dense_i8(
            tile.data_ptr(), tdt, F, K, ld_tile, masks.data_ptr(), M, ld_masks, out.data_ptr(),
            ld_out, int(bool(accumulate)), sig_sum.data_ptr() if sig_sum is not None else None,
            ws.data_ptr(), ws.numel(), _stream_ptr(tile.device)))
    return out


def synth_fill(shape, dtype, seed, device, start=0):
    """Device twin of oracle.synth.dataset: counter-based synthetic data."""
    lib = get_lib()
    tdt = {np.dtype('float32'): torch.float32, np.dtype('uint16'): torch.uint16}[np.dtype(dtype)]
    out = torch.empty(tuple(shape), dtype=tdt, device=device)
    with torch.cuda.device(out.device):
        check(lib.ltb200_synth_fill(out.data_ptr(), _TORCH_DTYPES[tdt], int(start), out.numel(),
                                    int(seed) & 0xFFFFFFFF, _stream_ptr(out.device)))
    return out


def last_kernel():
    return get_lib().ltb200_last_kernel()


def launch_count(reset=False):
    return get_lib().ltb200_launch_count(int(reset))


def masks_csc(tile, indptr, indices, values, n_masks, out=None, accumulate=False):
    """Sparse masks as CSC over (sig_size, n_masks): ``out[f, m] (+)= sum_i tile[f, idx_i] * v_i``."""
    lib = get_lib()
    _require_cuda(tile, 'tile')
    if tile.dim() != 2:
        raise ValueError('tile must be 2D (frames, sig_size)')
    if tile.shape[1] > 0 and tile.stride(1) != 1:
        tile = tile.contiguous()
    F, K = tile.shape
    if out is None:
        out = torch.zeros((F, n_masks), dtype=torch.float32, device=tile.device)
        accumulate = False
    ld_tile = tile.stride(0) if F > 1 else max(K, 1)
    ld_out = out.stride(0) if F > 1 else max(n_masks, 1)
    with torch.cuda.device(tile.device):
        check(lib.ltb200_masks_csc(
            tile.data_ptr(), _TORCH_DTYPES[tile.dtype], F, K, ld_tile, indptr.data_ptr(),
            indices.data_ptr(), values.data_ptr(), int(n_masks), out.data_ptr(), ld_out,
            int(bool(accumulate)), _stream_ptr(tile.device)))
    return out


def set_k1_variant(variant):
    """0 auto, 1 FFMA2 even/odd-pixel tile, 2 FFMA2 mask-pair tile, 3 tcgen05 kernel (K6)"""
    check(get_lib().ltb200_set_k1_variant(int(variant)))


def masks_shifted(tile, masks, shifts, out=None, accumulate=False, banded=None):
    """per-frame shifted masks: tile (F, sy, sx), masks (M, sy*sx) float32 or float64 (-> float64
    accumulation and result), shifts int32 (F, 2) or (1, 2) (dy, dx) -> out (F, M).
    ``banded``: None = automatic (float32 results, >= 64 frames, host-side shifts: the frames
    are sorted by dy and streamed through a shared-memory band of the masks), False = the
    warp-per-frame kernel."""
    lib = get_lib()
    _require_cuda(tile, 'tile')
    _require_cuda(masks, 'masks')
    F, sy, sx = tile.shape
    tile = tile.contiguous()
    masks = masks.contiguous()
    M = masks.shape[0]
    if masks.dtype not in (torch.float32, torch.float64):
        raise TypeError(f'masks must be float32 or float64, got {masks.dtype}')
    f64 = masks.dtype == torch.float64
    host_shifts = None
    if isinstance(shifts, np.ndarray):
        host_shifts = np.ascontiguousarray(shifts, dtype=np.int32).reshape(-1, 2)
        shifts = torch.from_numpy(host_shifts)
    elif not shifts.is_cuda:
        host_shifts = shifts.to(torch.int32).contiguous().numpy().reshape(-1, 2)
    shifts = shifts.to(device=tile.device, dtype=torch.int32).contiguous()
    per_frame = int(shifts.shape[0] != 1)
    if per_frame and shifts.shape[0] != F:
        raise ValueError('need one (dy, dx) per frame or a single pair')
    if out is None:
        out = torch.zeros((F, M), dtype=masks.dtype, device=tile.device)
        accumulate = False
    elif out.dtype != masks.dtype or (M > 1 and out.stride(1) != 1):
        raise ValueError('out must have the dtype of the masks and unit inner stride')
    ld_out = out.stride(0) if F > 1 else max(M, 1)
    want_banded = (banded is not False and not f64 and host_shifts is not None and F >= 64
                   and M > 0 and tile.dtype in (torch.float32, torch.uint16, torch.uint8,
                                                torch.int16))
    if want_banded:
        # sort the frames by dy: a chunk of 256 consecutive frames then spans few mask rows
        if per_frame:
            order = np.argsort(host_shifts[:, 0], kind='stable').astype(np.int32)
            dys = host_shifts[order, 0]
            starts = np.arange(0, F, 256)
            ends = np.minimum(starts + 256, F) - 1
            max_span = int((dys[ends] - dys[starts]).max())
        else:
            order = np.arange(F, dtype=np.int32)
            max_span = 0
        need = lib.ltb200_masks_shifted_banded_workspace(F, sy, sx, M, max_span)
        if need:
            ws = _workspace(tile.device, need)
            order_t = torch.from_numpy(order).to(tile.device)
            with torch.cuda.device(tile.device):
                check(lib.ltb200_masks_shifted_banded(
                    tile.data_ptr(), _TORCH_DTYPES[tile.dtype], F, sy, sx, sy * sx,
                    masks.data_ptr(), M, masks.shape[1], shifts.data_ptr(), per_frame,
                    order_t.data_ptr(), max_span, out.data_ptr(), ld_out,
                    int(bool(accumulate)), ws.data_ptr(), ws.numel(),
                    _stream_ptr(tile.device)))
            return out
    fn = lib.ltb200_masks_shifted_f64 if f64 else lib.ltb200_masks_shifted
    with torch.cuda.device(tile.device):
        check(fn(tile.data_ptr(), _TORCH_DTYPES[tile.dtype], F, sy, sx, sy * sx, masks.data_ptr(),
                 M, masks.shape[1], shifts.data_ptr(), per_frame, out.data_ptr(), ld_out,
                 int(bool(accumulate)), _stream_ptr(tile.device)))
    return out


def probe_read(buf, mode=0):
    """read-only streaming probe over a CUDA tensor (measurement only): mode 0 bulk-TMA ingest
    into shared memory, mode 1 LDG.128 (include/ltb200.h: ltb200_probe_read)"""
    _require_cuda(buf, 'buf')
    lib = get_lib()
    sink = _workspace(buf.device, 4)
    with torch.cuda.device(buf.device):
        check(lib.ltb200_probe_read(buf.data_ptr(), buf.numel() * buf.element_size(), int(mode),
                                    sink.data_ptr(), _stream_ptr(buf.device)))
