"""ApplyMasksUDF on the B200 engine.

Same constructor, buffers, dtype rules and tile protocol as the reference
(src/libertem/udf/masks.py:12-404); the contraction itself runs in the CUDA kernels behind
the C ABI (K1 dense / K2 sparse), never on the CPU.
"""
import numpy as np
import torch

from .base import UDF, UDFException
from ..common.buffers import AuxBufferWrapper, torch_dtype
from ..common.container import MaskContainer
from .. import engine


def as_device_tile(tile, device):
    """numpy / CPU tensors are staged to the device (the reference's BACKEND_CUDA hands numpy
    tiles to a CUDA worker, common/udf.py:45); CUDA tensors pass through."""
    if isinstance(tile, np.ndarray):
        if tile.dtype == np.uint16:
            t = torch.from_numpy(np.ascontiguousarray(tile).view(np.int16)).to(device)
            return t.view(torch.uint16)
        if tile.dtype == np.uint32:
            t = torch.from_numpy(np.ascontiguousarray(tile).view(np.int32)).to(device)
            return t.view(torch.uint32)
        return torch.from_numpy(np.ascontiguousarray(tile)).to(device)
    if not tile.is_cuda:
        return tile.to(device)
    return tile


class ApplyMasksEngine:
    """``process_tile(tile) -> (frames, masks)``: udf/masks.py:12-125."""

    def __init__(self, masks: MaskContainer, meta, use_torch=True):
        self.masks = masks
        self.meta = meta
        self.result_dtype = np.result_type(meta.input_dtype, masks.dtype)
        # complex frames (reference udf/masks.py:360-368: result_type(input, mask) is complex;
        # tests/analysis/test_analysis_masks.py:151-212): the tile is read ONCE as its
        # interleaved (re, im) float view (F, 2K) against mask rows expanded over that view --
        # the same dense kernels, no separate real / imaginary planes
        self.complex_input = np.dtype(meta.input_dtype).kind == 'c'
        self.sparse = bool(masks.use_sparse) and self.result_dtype == np.float32
        # dtype the kernels accumulate in
        if self.result_dtype in (np.float32, np.complex64):
            self.compute = np.float32
        else:
            self.compute = np.float64

    def group_plan(self, sig_slice=None):
        """K4 plan when the masks are complex64, numerous and group-sparse (all members of a
        run of consecutive masks share one support -- radial Fourier rings), else None.
        Built once per MaskContainer (shared by the partition copies of the UDF)."""
        if self.result_dtype != np.complex64 or len(self.masks) <= 12 or self.complex_input:
            return None
        sl = self.meta.sig_slice if sig_slice is None else sig_slice
        return self.masks.get_group_plan(sl, self.device)

    @property
    def device(self):
        return self.meta.device if self.meta.device is not None else torch.device('cuda')

    def _mask_dtype_for_device(self):
        if self.result_dtype.kind == 'c':
            return np.complex64 if self.compute == np.float32 else np.complex128
        return self.compute

    def n_real_columns(self):
        n = len(self.masks)
        return 2 * n if self.result_dtype.kind == 'c' else n

    def dense_rows(self, sig_slice=None):
        """device mask rows ``(R, K_tile)`` in the compute dtype (complex -> re/im rows); for
        complex frames ``(2M, 2 K_tile)`` rows over the interleaved (re, im) view of the tile"""
        sl = self.meta.sig_slice if sig_slice is None else sig_slice
        if self.complex_input:
            return self.masks.get_device_dense_for_complex(sl, self.device, self.compute)
        return self.masks.get_device_dense(sl, self.device, self._mask_dtype_for_device())

    @staticmethod
    def _float_view(flat_tile):
        """(F, K) complex tile -> (F, 2K) float view [re0, im0, re1, im1, ...] (no copy)"""
        if flat_tile.is_complex():
            if flat_tile.shape[1] > 0 and flat_tile.stride(1) != 1:
                flat_tile = flat_tile.contiguous()
            return torch.view_as_real(flat_tile).reshape(flat_tile.shape[0], -1)
        return flat_tile

    def process_flat(self, flat_tile, out=None, accumulate=False, sig_slice=None):
        """the seam of masks.py:31-83: ``flat_tile (F, K) -> (F, R)`` real columns"""
        if flat_tile.dtype == torch.float32:
            plan = self.group_plan(sig_slice)
            if plan is not None:
                from .. import group_masks as gm
                cout = None
                if out is not None:
                    cout = torch.view_as_complex(out.reshape(out.shape[0], -1, 2))
                res = gm.group_masks(flat_tile, plan, out=cout, accumulate=accumulate)
                return torch.view_as_real(res).reshape(res.shape[0], -1)
        if self.sparse and flat_tile.dtype in (torch.float32, torch.uint16, torch.uint8,
                                               torch.int16, torch.int8):
            sl = self.meta.sig_slice if sig_slice is None else sig_slice
            indptr, indices, values = self.masks.get_device_csc(sl, self.device)
            return engine.masks_csc(flat_tile, indptr, indices, values, len(self.masks),
                                    out=out, accumulate=accumulate)
        rows = self.dense_rows(sig_slice)
        if self.complex_input:
            if not flat_tile.is_complex():
                flat_tile = flat_tile.to(torch.complex64 if self.compute == np.float32
                                         else torch.complex128)
            flat_tile = self._float_view(flat_tile)
        if self.compute == np.float64 and flat_tile.dtype in (torch.uint16, torch.uint8):
            flat_tile = flat_tile.to(torch.int32)
        return engine.masks_dense(flat_tile, rows, out=out, accumulate=accumulate)

    def as_result(self, real_cols):
        """(F, R) real columns -> (F, M) in the declared result dtype"""
        if self.result_dtype.kind == 'c':
            return torch.view_as_complex(real_cols.reshape(real_cols.shape[0], -1, 2))
        want = torch_dtype(self.result_dtype)
        return real_cols if real_cols.dtype == want else real_cols.to(want)

    def process_tile(self, tile):
        tile = as_device_tile(tile, self.device)
        flat = tile.reshape(tile.shape[0], -1)
        return self.as_result(self.process_flat(flat))

    def process_frame_shifted(self, frame, shifts):
        """masks shifted by (dy, dx) relative to the frame; non-overlapping parts are dropped
        (udf/masks.py:85-124)."""
        sig_shape = tuple(self.meta.dataset_shape.sig)
        n = len(self.masks)
        sig = self.meta.sig_slice
        shifts = tuple(int(s) for s in shifts)
        left = sig.intersection_with(sig.shift_by(shifts))
        right = sig.intersection_with(sig.shift_by(tuple(-s for s in shifts)))
        if left.is_null():
            return torch.zeros((n,), dtype=torch_dtype(self.result_dtype), device=self.device)
        frame = as_device_tile(frame, self.device).reshape(sig_shape)
        data = left.get(frame).reshape(1, -1)
        rows = self.dense_rows()
        if self.complex_input:
            # rows run over the interleaved (re, im) view: (R, sy, sx, 2)
            rows = rows.reshape((rows.shape[0],) + tuple(sig.shape) + (2,))
            sub = right.get(rows[..., 0], sig_only=True)
            sub = torch.stack([sub, right.get(rows[..., 1], sig_only=True)], dim=-1)
            sub = sub.reshape(rows.shape[0], -1).contiguous()
            data = self._float_view(data.contiguous().to(
                torch.complex64 if self.compute == np.float32 else torch.complex128))
        else:
            rows = rows.reshape((rows.shape[0],) + tuple(sig.shape))
            sub = right.get(rows, sig_only=True).reshape(rows.shape[0], -1).contiguous()
        res = engine.masks_dense(data.contiguous(), sub)
        return self.as_result(res).reshape((n,))


class ApplyMasksUDF(UDF):
    """Apply masks to frames; result buffer ``intensity`` of shape ``(*nav, len(masks))``.

    Parameters as in the reference (udf/masks.py:255-256): ``mask_factories, use_torch=True,
    use_sparse=None, mask_count=None, mask_dtype=None, preferred_dtype=None, backends=None,
    shifts=None``.  ``use_torch`` and ``backends`` are accepted for drop-in compatibility;
    there is one backend here (CUDA).
    """

    _slab_buffer = 'intensity'     # float32 nav buffer the fused dense kernel may write directly

    def __init__(self, mask_factories, use_torch=True, use_sparse=None, mask_count=None,
                 mask_dtype=None, preferred_dtype=None, backends=None, shifts=None, **kwargs):
        _backends = backends
        if backends is None:
            backends = self.BACKEND_ALL
        backends = tuple(b for b in backends if b in ('cuda', 'cupy', 'numpy'))
        if len(backends) == 0:
            raise ValueError(f'No compatible backend found in {_backends}')
        if shifts is not None:
            if isinstance(use_sparse, str) and use_sparse.startswith('scipy.sparse'):
                raise ValueError(f'Sparse backend {use_sparse} not supported for shifts, '
                                 'use sparse.pydata instead.')
            if not isinstance(shifts, AuxBufferWrapper):
                shifts = np.asarray(shifts)
        self._mask_container = None
        super().__init__(mask_factories=mask_factories, use_torch=use_torch,
                         use_sparse=use_sparse, mask_count=mask_count, mask_dtype=mask_dtype,
                         preferred_dtype=preferred_dtype, backends=backends, shifts=shifts,
                         **kwargs)

    def copy_for_partition(self):
        # partition instances of one process share the (lazily computed, cached) mask
        # container: masks are built and uploaded once per run, not once per partition
        new = super().copy_for_partition()
        new._mask_container = self.masks
        return new

    def get_preferred_input_dtype(self):
        if self.params.preferred_dtype is None:
            return super().get_preferred_input_dtype()
        return self.params.preferred_dtype

    def get_mask_dtype(self):
        if self.params.mask_dtype is None:
            return self.masks.dtype
        return self.params.mask_dtype

    def get_mask_count(self):
        if self.params.mask_count is None:
            return len(self.masks)
        return self.params.mask_count

    @property
    def masks(self):
        if self._mask_container is None:
            self._mask_container = self._make_mask_container()
        return self._mask_container

    def _make_mask_container(self):
        p = self.params
        default_sparse = 'scipy.sparse' if p.shifts is None else 'sparse.pydata'
        return MaskContainer(p.mask_factories, dtype=p.mask_dtype, use_sparse=p.use_sparse,
                             count=p.mask_count, backend='cuda', default_sparse=default_sparse)

    def get_task_data(self):
        return {'engine': ApplyMasksEngine(self.masks, self.meta, self.params.use_torch)}

    def get_result_buffers(self):
        dtype = np.result_type(self.meta.input_dtype, self.get_mask_dtype())
        return {'intensity': self.buffer(kind='nav', extra_shape=(self.get_mask_count(),),
                                         dtype=dtype, where='device')}

    def get_backends(self):
        return self.params.backends

    def get_method(self):
        return 'frame' if self.params.get('shifts') is not None else 'tile'

    def process_tile(self, tile):
        eng = self.task_data['engine']
        view = self.results.intensity
        tile = as_device_tile(tile, eng.device)
        flat = tile.reshape(tile.shape[0], -1)
        if view.dtype == torch.float32 and view.is_cuda:
            # results.intensity[:] += ... fused into the kernel's store (accumulate)
            eng.process_flat(flat, out=view, accumulate=True)
        else:
            view[:] += self.forbuf(eng.process_tile(tile), view)

    def process_tile_shifted(self, tile, shifts):
        """all frames of a full-frame tile with their own (dy, dx) in ONE launch (K5) instead
        of the reference's frame-by-frame loop (udf/masks.py:85-124): float32 and float64
        results, complex masks as interleaved (re, im) rows; returns False if this case needs
        the loop (complex frames, sub-frame tiles)"""
        eng = self.task_data['engine']
        view = self.results.intensity
        sig = tuple(self.meta.dataset_shape.sig)
        if (eng.complex_input or len(sig) != 2 or tuple(tile.shape[1:]) != sig
                or not view.is_cuda or eng.result_dtype not in (
                    np.float32, np.float64, np.complex64, np.complex128)):
            return False
        f64 = eng.compute == np.float64
        ok = (torch.float32, torch.float64, torch.uint16, torch.uint8, torch.int16, torch.int32)
        if tile.dtype not in ok:
            return False
        rows = eng.dense_rows()          # complex masks: (2M, K) interleaved (re, im) rows
        if rows.dtype != (torch.float64 if f64 else torch.float32):
            return False
        out = view
        if view.is_complex():
            out = torch.view_as_real(view).reshape(view.shape[0], -1)
        if out.dtype != rows.dtype or (out.shape[1] > 1 and out.stride(1) != 1):
            return False
        sh = np.ascontiguousarray(shifts, dtype=np.int32).reshape(-1, 2)
        engine.masks_shifted(tile, rows, sh, out=out, accumulate=True)
        return True

    def process_frame(self, frame):
        shifts = self._current_shift
        view = self.results.intensity
        view[:] += self.forbuf(self.task_data['engine'].process_frame_shifted(frame, shifts),
                               view)

    # -- fused-pass contribution (see libertem_b200/runner.py) ----------------------------------
    def _fused_spec(self):
        eng = self.task_data['engine']
        if self.params.get('shifts') is not None:
            return None
        if eng.sparse:
            return {'kind': 'csc', 'buffer': 'intensity', 'engine': eng}
        if eng.compute != np.float32:
            return None
        if eng.group_plan() is not None:
            return {'kind': 'own_pass', 'buffer': 'intensity', 'engine': eng}
        return {'kind': 'dense', 'buffer': 'intensity', 'engine': eng,
                'columns': eng.n_real_columns()}
