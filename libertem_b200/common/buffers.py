"""Result buffers: the ownership conventions of the reference ``BufferWrapper``
(src/libertem/common/buffers.py:326-946), backed by torch tensors.

* ``kind='nav'``: one entry per frame, stored flat as ``(n_frames_or_roi_count, *extra_shape)``
  (``raw_data``), exposed nav-shaped through ``.data`` with NaN (float) / 0 (int) outside an ROI
  (buffers.py:470-505).
* ``kind='sig'``: one entry per detector pixel; ``kind='single'``: just ``extra_shape``.
* ``where='device'`` buffers live in HBM (the kernels write them directly); everything is
  allocated by the runtime, a UDF only writes through the views it is handed.
* ``use``: None (public), 'private' (never returned), 'result_only' (only from get_results).
"""
import numpy as np
import torch

from .shape import Shape

_NP2TORCH = {
    np.dtype('float32'): torch.float32, np.dtype('float64'): torch.float64,
    np.dtype('complex64'): torch.complex64, np.dtype('complex128'): torch.complex128,
    np.dtype('int8'): torch.int8, np.dtype('uint8'): torch.uint8, np.dtype('int16'): torch.int16,
    np.dtype('uint16'): torch.uint16, np.dtype('int32'): torch.int32,
    np.dtype('uint32'): torch.uint32, np.dtype('int64'): torch.int64,
    np.dtype('uint64'): torch.uint64, np.dtype('bool'): torch.bool,
}


def torch_dtype(dt):
    return _NP2TORCH[np.dtype(dt)]


def to_numpy(t):
    """device/host tensor -> numpy (the reference's ``export()`` D2H, buffers.py:901-907)."""
    if isinstance(t, np.ndarray):
        return t
    t = t.detach()
    if t.is_cuda:
        t = t.cpu()
    if t.dtype in (torch.uint16, torch.uint32, torch.uint64):
        signed = {torch.uint16: torch.int16, torch.uint32: torch.int32,
                  torch.uint64: torch.int64}[t.dtype]
        unsigned = {torch.uint16: np.uint16, torch.uint32: np.uint32,
                    torch.uint64: np.uint64}[t.dtype]
        return t.view(signed).numpy().view(unsigned)
    return t.numpy()


class BufferWrapper:
    def __init__(self, kind, extra_shape=(), dtype='float32', where=None, use=None):
        if kind not in ('nav', 'sig', 'single'):
            raise ValueError("kind must be one of 'nav', 'sig', 'single'")
        if use not in (None, 'private', 'result_only'):
            raise ValueError("use must be None, 'private' or 'result_only'")
        self._kind = kind
        self._extra_shape = tuple(int(e) for e in extra_shape)
        self._dtype = np.dtype(dtype)
        self._where = where
        self._use = use
        self._data = None           # torch.Tensor or np.ndarray, flat layout
        self._ds_shape = None
        self._roi = None
        self._n_frames = None       # frames covered by this buffer (dataset or partition)

    kind = property(lambda self: self._kind)
    extra_shape = property(lambda self: self._extra_shape)
    dtype = property(lambda self: self._dtype)
    where = property(lambda self: self._where)
    use = property(lambda self: self._use)
    roi = property(lambda self: self._roi)

    def new_like(self):
        return BufferWrapper(self._kind, self._extra_shape, self._dtype, self._where, self._use)

    # -- shapes ---------------------------------------------------------------------------
    def set_shape_ds(self, dataset_shape, roi=None):
        self._ds_shape = dataset_shape
        self._roi = None if roi is None else np.asarray(roi).reshape(-1).astype(bool)
        n = dataset_shape.nav.size
        self._n_frames = int(self._roi.sum()) if self._roi is not None else n

    def set_shape_partition(self, dataset_shape, n_frames_in_partition, roi=None):
        self._ds_shape = dataset_shape
        self._roi = None
        self._n_frames = int(n_frames_in_partition)

    def _flat_shape(self):
        if self._kind == 'nav':
            return (self._n_frames,) + self._extra_shape
        if self._kind == 'sig':
            return (self._ds_shape.sig.size,) + self._extra_shape
        return self._extra_shape if self._extra_shape else (1,)

    def allocate(self, device=None):
        """zeros; on ``device`` when ``where='device'`` else on the host."""
        shape = self._flat_shape()
        if self._where == 'device' and device is not None:
            self._data = torch.zeros(shape, dtype=torch_dtype(self._dtype), device=device)
        else:
            self._data = torch.zeros(shape, dtype=torch_dtype(self._dtype))
        return self

    def has_data(self):
        return self._data is not None

    def replace_array(self, arr):
        self._data = arr

    # -- access ---------------------------------------------------------------------------
    @property
    def tensor(self):
        """flat storage (torch tensor; device for where='device')"""
        return self._data

    def rows(self, start, stop):
        """view of nav rows [start, stop) -- what a tile/partition writes into
        (buffers.py:792-821)"""
        if self._kind != 'nav':
            return self._data
        return self._data[start:stop]

    @property
    def raw_data(self):
        """flat numpy array (nav compressed by the roi)"""
        arr = to_numpy(self._data) if self._data is not None else None
        if arr is not None and self._kind == 'single' and not self._extra_shape:
            return arr
        return arr

    @property
    def data(self):
        """user-facing array: nav-/sig-shaped, NaN-(or zero-)filled outside the roi"""
        arr = self.raw_data
        if self._kind == 'nav':
            nav = tuple(self._ds_shape.nav)
            if self._roi is None:
                return arr.reshape(nav + self._extra_shape)
            fill = np.nan if arr.dtype.kind in 'fc' else 0
            full = np.full((len(self._roi),) + self._extra_shape, fill, dtype=arr.dtype)
            full[self._roi] = arr
            return full.reshape(nav + self._extra_shape)
        if self._kind == 'sig':
            return arr.reshape(tuple(self._ds_shape.sig) + self._extra_shape)
        return arr.reshape(self._extra_shape if self._extra_shape else (1,))

    def __array__(self, dtype=None, copy=None):
        d = self.data
        return d if dtype is None else d.astype(dtype)

    def __repr__(self):
        return (f'<BufferWrapper kind={self._kind} dtype={self._dtype} '
                f'extra_shape={self._extra_shape}>')


class AuxBufferWrapper(BufferWrapper):
    """Per-frame auxiliary input (e.g. ApplyMasksUDF ``shifts``), buffers.py:995-1048."""

    def set_buffer(self, data):
        arr = np.asarray(data, dtype=self._dtype)
        self._aux = arr

    def for_frames(self, ds_shape, roi=None):
        n = ds_shape.nav.size
        arr = self._aux.reshape((n,) + self._extra_shape)
        if roi is not None:
            arr = arr[np.asarray(roi).reshape(-1).astype(bool)]
        return arr


def check_cast(fromvar, tovar):
    """safe-cast check of the default merge (reference udf/base.py:1768-1771)."""
    f = np.dtype(str(fromvar.dtype).replace('torch.', '')) if not isinstance(
        fromvar, np.ndarray) else fromvar.dtype
    t = np.dtype(str(tovar.dtype).replace('torch.', '')) if not isinstance(
        tovar, np.ndarray) else tovar.dtype
    if not np.can_cast(f, t, casting='safe'):
        raise TypeError(f'Unsafe automatic casting from {f} to {t}')


def reshaped_view(shape_like: Shape):
    return tuple(shape_like)
