"""UDF plugin protocol: the drop-in boundary of the reference's ``libertem.udf.base.UDF``
(src/libertem/udf/base.py:1270-1732) for the masked-reduction hot path.

Same life cycle and method names: ``__init__(**kwargs)`` (kwargs re-instantiate the UDF per
partition, base.py:1316-1325), ``get_result_buffers`` / ``get_task_data`` / ``get_backends`` /
``get_preferred_input_dtype`` / ``process_tile`` / ``postprocess`` / ``merge`` /
``get_results``; buffers are declared with ``self.buffer(...)`` and written through
``self.results.<name>`` views that the runtime sets up per tile.  Tiles are CUDA tensors
(``BACKEND_CUDA`` device class): there is no CPU execution path.
"""
import numpy as np

from ..common.buffers import BufferWrapper, AuxBufferWrapper, check_cast
from ..common.shape import Shape
from ..common.slice import Slice


class UDFException(Exception):
    pass


class UDFParams:
    """kwargs of the UDF, attribute- and item-accessible (base.py:1304-1311)."""

    def __init__(self, kwargs):
        self._kwargs = dict(kwargs)

    def __getattr__(self, k):
        if k.startswith('_'):
            raise AttributeError(k)
        try:
            return self._kwargs[k]
        except KeyError:
            raise AttributeError(k)

    def __getitem__(self, k):
        return self._kwargs[k]

    def get(self, k, default=None):
        return self._kwargs.get(k, default)

    def __contains__(self, k):
        return k in self._kwargs


class UDFData:
    """Named buffers with per-tile views (base.py:480-700)."""

    def __init__(self, buffers):
        object.__setattr__(self, '_buffers', dict(buffers))
        object.__setattr__(self, '_views', {})

    def __getattr__(self, k):
        if k.startswith('_'):
            raise AttributeError(k)
        views = object.__getattribute__(self, '_views')
        if k in views:
            return views[k]
        bufs = object.__getattribute__(self, '_buffers')
        if k in bufs:
            return bufs[k].tensor
        raise AttributeError(k)

    def __setattr__(self, k, v):
        raise AttributeError('assign into the buffer view instead: results.%s[:] = ...' % k)

    def get_buffer(self, name):
        return self._buffers[name]

    def set_view(self, name, view):
        self._views[name] = view

    def clear_views(self):
        self._views.clear()

    def keys(self):
        return self._buffers.keys()

    def items(self):
        return self._buffers.items()

    def __iter__(self):
        return iter(self._buffers)


class MergeAttrMapping:
    """attribute/item access to flat arrays for ``merge(dest, src)`` (base.py:703-735)."""

    def __init__(self, arrays):
        object.__setattr__(self, '_arrays', dict(arrays))

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, '_arrays')[k]
        except KeyError:
            raise AttributeError(k)

    def __getitem__(self, k):
        return self._arrays[k]

    def __iter__(self):
        return iter(self._arrays)

    def keys(self):
        return self._arrays.keys()


class UDFMeta:
    """What a UDF may know about the run (base.py:330-478)."""

    def __init__(self, partition_slice=None, dataset_shape=None, roi=None, dataset_dtype=None,
                 input_dtype=None, tiling_scheme=None, device=None, valid_nav_mask=None):
        self.partition_slice = partition_slice
        self.dataset_shape = dataset_shape
        self.roi = roi
        self.dataset_dtype = None if dataset_dtype is None else np.dtype(dataset_dtype)
        self.input_dtype = None if input_dtype is None else np.dtype(input_dtype)
        self.tiling_scheme = tiling_scheme
        self.tiling_scheme_idx = 0
        self.slice = None
        self.device = device
        self.device_class = 'cuda'
        self.array_backend = 'cuda'
        self.threads_per_worker = 1
        self._valid_nav_mask = valid_nav_mask

    @property
    def sig_slice(self):
        if self.slice is not None:
            return self.slice.discard_nav()
        sig = self.dataset_shape.sig
        return Slice(origin=(0,) * len(sig), shape=Shape(tuple(sig), sig_dims=len(sig)))

    def get_valid_nav_mask(self, full_nav=False):
        """processed-so-far mask over the (roi-compressed or full) flat nav axis
        (base.py:449-478); after a complete run: every roi position."""
        n = self.dataset_shape.nav.size
        roi = None if self.roi is None else np.asarray(self.roi).reshape(-1).astype(bool)
        valid = self._valid_nav_mask
        if valid is None:
            valid = np.ones(n if roi is None else int(roi.sum()), dtype=bool)
        if full_nav and roi is not None:
            full = np.zeros(n, dtype=bool)
            full[roi] = valid
            return full
        return valid


def _default_merge_all(udf, ordered_results):
    """concatenate the partitions' nav buffers in partition order (base.py:985-1002)"""
    import torch
    if udf.requires_custom_merge_all:
        raise NotImplementedError(
            "Default merging only works for kind='nav' buffers. "
            "Please implement a suitable custom merge_all function.")
    chunks = {}
    for b in ordered_results.values():
        for key in b:
            chunks.setdefault(key, []).append(getattr(b, key))
    return {k: torch.cat(v, dim=0) for k, v in chunks.items()}


class UDF:
    USE_NATIVE_DTYPE = bool
    TILE_SIZE_BEST_FIT = object()
    TILE_SIZE_MAX = np.inf
    TILE_DEPTH_DEFAULT = object()
    TILE_DEPTH_MAX = np.inf
    BACKEND_NUMPY = 'numpy'
    BACKEND_CUDA = 'cuda'
    BACKEND_CUPY = 'cupy'
    BACKEND_ALL = ('cuda',)     # this runtime has exactly one device class

    def __init__(self, **kwargs):
        self._kwargs = kwargs
        self.params = UDFParams(kwargs)
        self.task_data = None
        self.results = None
        self._meta = None

    # -- life cycle ---------------------------------------------------------------------------
    def copy_for_partition(self):
        """fresh instance from the constructor kwargs (base.py:1316-1325)"""
        return self.__class__(**self._kwargs)

    @property
    def meta(self):
        return self._meta

    def set_meta(self, meta):
        self._meta = meta

    def get_result_buffers(self):
        raise NotImplementedError()

    def get_task_data(self):
        return {}

    def get_preferred_input_dtype(self):
        return np.float32

    def get_backends(self):
        return (self.BACKEND_CUDA,)

    def get_tiling_preferences(self):
        return {'depth': self.TILE_DEPTH_DEFAULT, 'total_size': self.TILE_SIZE_MAX}

    def get_method(self):
        """'tile' | 'frame' | 'partition': which process_* method this UDF implements
        (base.py:1148-1176)"""
        cls = type(self)
        for name in ('tile', 'frame', 'partition'):
            if getattr(cls, 'process_' + name, None) is not None:
                return name
        raise UDFException('UDF should implement one of the process_* methods')

    def preprocess(self):
        pass

    def postprocess(self):
        pass

    # -- declaring buffers ----------------------------------------------------------------------
    def buffer(self, kind, extra_shape=(), dtype='float32', where=None, use=None):
        return BufferWrapper(kind, extra_shape, dtype, where, use)

    @classmethod
    def aux_data(cls, data, kind, extra_shape=(), dtype='float32'):
        buf = AuxBufferWrapper(kind, extra_shape, dtype)
        buf.set_buffer(data)
        return buf

    def forbuf(self, arr, target):
        """condition ``arr`` for assignment into the buffer view ``target`` (base.py:1576-1606)"""
        import torch
        if isinstance(arr, np.ndarray):
            arr = torch.from_numpy(arr)
        if isinstance(target, torch.Tensor):
            arr = arr.to(device=target.device)
            return arr.reshape(target.shape)
        return arr

    # -- merging / results ------------------------------------------------------------------------
    @property
    def requires_custom_merge(self):
        return any(b.kind != 'nav' and b.use != 'result_only'
                   for b in self.get_result_buffers().values())

    def merge(self, dest, src):
        """default: copy partition rows into their slice of the dataset buffer; only valid
        for kind='nav' buffers (base.py:1420-1453)"""
        if self.requires_custom_merge:
            raise NotImplementedError(
                "Default merging only works for kind='nav' buffers. "
                "Please implement a suitable custom merge function.")
        for k in dest:
            check_cast(getattr(src, k), getattr(dest, k))
            getattr(dest, k)[:] = getattr(src, k)

    @property
    def requires_custom_merge_all(self):
        """any buffer with ``kind != 'nav'``: the default merge_all (concatenation) does not
        apply (base.py:1405-1418)"""
        return any(b.kind != 'nav' for b in self.get_result_buffers().values())

    #: UDFs whose ``merge`` is a plain ``dest += src`` on every buffer may set this so that the
    #: multi-rank merge uses one all-reduce(sum) instead of replaying ``merge`` per rank
    _additive_merge = False

    def _do_merge_all(self, ordered_results):
        """combine the ordered partition results ``{Slice: MergeAttrMapping}`` in one go
        (base.py:1208-1224): the UDF's own ``merge_all`` when it defines one, else
        concatenation of the nav buffers (``_default_merge_all``, base.py:985-1002)."""
        custom = getattr(self, 'merge_all', None)
        if custom is not None:
            tmp = custom(ordered_results)
        else:
            tmp = _default_merge_all(self, ordered_results)
        if not set(tmp.keys()).issubset(set(self.results.keys())):
            raise ValueError('Returned result names from merge_all (%s) are not contained within '
                             'declared result buffer names (%s)'
                             % ([*tmp.keys()], [*self.results.keys()]))
        for key, value in tmp.items():
            buf = self.results.get_buffer(key)
            cur = buf.tensor
            import torch
            if isinstance(value, np.ndarray):
                value = torch.from_numpy(value)
            if cur is not None:
                if tuple(value.shape) != tuple(cur.shape) and value.numel() != cur.numel():
                    raise ValueError("merge_all result '%s' has shape %s, buffer has %s"
                                     % (key, tuple(value.shape), tuple(cur.shape)))
                value = value.reshape(cur.shape).to(device=cur.device, dtype=cur.dtype)
            buf.replace_array(value)

    def get_results(self):
        """default: every non-private buffer as it is (base.py:1455-1493)"""
        decl = self.get_result_buffers()
        return {k: self.results.get_buffer(k).raw_data for k, v in decl.items()
                if v.use is None}
