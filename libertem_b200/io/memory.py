"""In-memory / synthetic datasets feeding the hot path.

``MemoryDataSet`` mirrors the constructor and partitioning of the reference's test/bench source
(src/libertem/io/dataset/memory.py:202-452; partition boundaries
io/dataset/base/partition.py:66-99) for data that lives either

* on the device (torch CUDA tensor): tiles are zero-copy views, one tile per partition for
  full frames -- the situation the HBM roofline is defined for; or
* on the host (numpy): depth-blocks of full frames are streamed H2D from (registered) pinned
  memory on a copy stream, double-buffered against the kernels (the `e2e` path).

``SyntheticDataSet`` generates any partition on the device from the counter-based generator
(twin of oracle/synth.py), so multi-hundred-GiB shapes never touch the host.
"""
import numpy as np
import torch

from ..common.shape import Shape
from ..common.slice import Slice
from .. import engine


def partition_boundaries(num_frames, num_partitions):
    """np.linspace(0, N, P+1, dtype=int): identical boundaries to Partition.make_slices
    (partition.py:72-88)"""
    num_partitions = max(1, min(int(num_partitions), int(num_frames)))
    b = np.linspace(0, num_frames, num=max(2, num_partitions + 1), endpoint=True, dtype=int)
    return [(int(a), int(c)) for a, c in zip(b[:-1], b[1:])]


def _np_to_torch(arr):
    if arr.dtype == np.uint16:
        return torch.from_numpy(arr.view(np.int16)).view(torch.uint16)
    if arr.dtype == np.uint32:
        return torch.from_numpy(arr.view(np.int32)).view(torch.uint32)
    if arr.dtype == np.uint64:
        return torch.from_numpy(arr.view(np.int64)).view(torch.uint64)
    return torch.from_numpy(arr)


class Partition:
    def __init__(self, dataset, idx, start, stop):
        self.dataset = dataset
        self.idx = idx
        self.start = start
        self.stop = stop
        sig = tuple(dataset.shape.sig)
        self.slice = Slice(origin=(start,) + (0,) * len(sig),
                           shape=Shape((stop - start,) + sig, sig_dims=len(sig)))

    @property
    def shape(self):
        return self.slice.shape

    def get_tiles(self, device, roi=None, tileshape=None):
        """yield ``(tile, f0, f1, tile_slice)``: ``tile`` is a CUDA tensor (frames, *sig_tile);
        f0/f1 are dataset frame indices; with an roi only selected frames are delivered and
        ``tile_slice.origin[0]`` counts roi-compressed frames (common/slice.py:376-395)."""
        yield from self.dataset._iter_tiles(self, device, roi, tileshape)


class _DataSetBase:
    def initialize(self, executor=None):
        return self

    @property
    def dtype(self):
        return self._dtype

    @property
    def shape(self):
        return self._shape

    def get_partitions(self):
        n = self._shape.nav.size
        for i, (a, b) in enumerate(partition_boundaries(n, self.num_partitions)):
            yield Partition(self, i, a, b)

    def _sig_slices(self, tileshape):
        sig = tuple(self._shape.sig)
        if tileshape is None:
            return [tuple(slice(0, s) for s in sig)]
        tsig = tuple(tileshape[1:])
        ranges = [range(0, s, t) for s, t in zip(sig, tsig)]
        out = []
        for idx in np.ndindex(*[len(r) for r in ranges]):
            out.append(tuple(slice(r[i], min(r[i] + t, s))
                             for r, i, t, s in zip(ranges, idx, tsig, sig)))
        return out

    def _emit(self, block, f0, f1, roi_flat, roi_before, tileshape):
        """split a depth block (device tensor of frames f0..f1) into sig tiles, apply the roi"""
        sig = tuple(self._shape.sig)
        if roi_flat is not None:
            sel = roi_flat[f0:f1]
            n_sel = int(sel.sum())
            if n_sel == 0:
                return
            if n_sel != f1 - f0:
                idx = torch.from_numpy(np.nonzero(sel)[0]).to(block.device)
                block = block.index_select(0, idx)
            origin0 = roi_before
        else:
            n_sel = f1 - f0
            origin0 = f0
        for sl in self._sig_slices(tileshape):
            full = all(s.start == 0 and s.stop == d for s, d in zip(sl, sig))
            tile = block if full else block[(slice(None),) + sl].contiguous()
            tshape = (n_sel,) + tuple(s.stop - s.start for s in sl)
            tslice = Slice(origin=(origin0,) + tuple(s.start for s in sl),
                           shape=Shape(tshape, sig_dims=len(sig)))
            yield tile, f0, f1, tslice


class MemoryDataSet(_DataSetBase):
    """``MemoryDataSet(data=..., tileshape=None, num_partitions=None, sig_dims=2)``.

    data: numpy array (host) or torch tensor (host or CUDA), shape ``nav + sig``.
    tileshape: optional ``(depth, *sig_tile)`` forcing the reference's sub-frame tiling;
    default: full frames, depth = whole partition for device data or ``tile_depth`` frames
    (about 256 MiB) for host data.  num_partitions defaults to 1 per visible shard.
    """

    def __init__(self, data=None, tileshape=None, num_partitions=None, sig_dims=2,
                 tile_depth=None, pin=True, **kwargs):
        if data is None:
            raise ValueError('MemoryDataSet needs data')
        self._is_torch = isinstance(data, torch.Tensor)
        self.data = data
        self._shape = Shape(tuple(data.shape), sig_dims=sig_dims)
        if self._is_torch:
            self._dtype = np.dtype(str(data.dtype).replace('torch.', ''))
        else:
            self._dtype = np.dtype(data.dtype)
        self.tileshape = tileshape
        self.num_partitions = 1 if num_partitions is None else num_partitions
        self.tile_depth = tile_depth
        self._pin = pin
        self._registered = False
        self._stage = {}

    def _flat(self):
        n = self._shape.nav.size
        sig = tuple(self._shape.sig)
        return self.data.reshape((n,) + sig)

    def _register(self):
        """page-lock the host array in place (no copy) so H2D copies are asynchronous"""
        if self._registered or not self._pin or self._is_torch:
            return
        arr = self.data
        if not arr.flags['C_CONTIGUOUS']:
            return
        rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
        self._registered = (int(rc) == 0)

    def release(self):
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.data.ctypes.data)
            self._registered = False

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _iter_tiles(self, part, device, roi, tileshape):
        tileshape = self.tileshape if tileshape is None else tileshape
        roi_flat = None if roi is None else np.asarray(roi).reshape(-1).astype(bool)
        flat = self._flat()
        sig = tuple(self._shape.sig)
        on_device = self._is_torch and flat.is_cuda
        if tileshape is not None:
            depth = int(tileshape[0])
        elif on_device:
            depth = part.stop - part.start
        elif self.tile_depth is not None:
            depth = int(self.tile_depth)
        else:
            frame_bytes = int(np.prod(sig)) * self._dtype.itemsize
            depth = max(1, (256 << 20) // max(frame_bytes, 1))
        blocks = [(f0, min(f0 + depth, part.stop)) for f0 in range(part.start, part.stop, depth)]
        roi_pos = None if roi_flat is None else int(roi_flat[:part.start].sum())

        def emit(block, f0, f1):
            nonlocal roi_pos
            before = roi_pos
            if roi_flat is not None:
                roi_pos += int(roi_flat[f0:f1].sum())
            yield from self._emit(block, f0, f1, roi_flat, before, tileshape)

        if on_device:
            for f0, f1 in blocks:
                yield from emit(flat[f0:f1], f0, f1)
            return
        # host data: double-buffered H2D on a side stream.  The staging buffers, the copy
        # stream and the "buffer free" events live on the dataset (one set per device / depth),
        # so they persist across partitions and runs: a buffer is never handed back to the
        # caching allocator while a kernel of an earlier partition may still read it, and
        # every copy into a buffer waits for the event recorded after its last consumer.
        self._register()
        src = flat if self._is_torch else _np_to_torch(flat)
        st = self._staging(device, depth, sig, src.dtype)
        main = torch.cuda.current_stream(device)
        copy_stream = st['stream']
        bufs, ready, free = st['bufs'], st['ready'], st['free']

        def launch(f0, f1):
            b = st['next'] & 1
            st['next'] += 1
            n = f1 - f0
            with torch.cuda.stream(copy_stream):
                if free[b] is not None:
                    copy_stream.wait_event(free[b])
                bufs[b][:n].copy_(src[f0:f1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready[b] = ev
            return b

        pending = launch(*blocks[0]) if blocks else None
        for i, (f0, f1) in enumerate(blocks):
            b = pending
            if i + 1 < len(blocks):
                pending = launch(*blocks[i + 1])
            main.wait_event(ready[b])
            try:
                yield from emit(bufs[b][:f1 - f0], f0, f1)
            finally:
                # also on an abandoned generator: whatever main has queued so far is the
                # last possible reader of this buffer
                ev = torch.cuda.Event()
                ev.record(main)
                free[b] = ev

    def _staging(self, device, depth, sig, dtype):
        """persistent double buffer + copy stream for host -> device streaming"""
        key = (str(device), int(depth), tuple(sig), dtype)
        st = self._stage.get(key)
        if st is None:
            main = torch.cuda.current_stream(device)
            copy_stream = torch.cuda.Stream(device=device)
            bufs = [torch.empty((depth,) + tuple(sig), dtype=dtype, device=device)
                    for _ in range(2)]
            # the allocator may hand out blocks whose previous users are still queued on the
            # main stream: order the first copies after everything queued so far
            copy_stream.wait_stream(main)
            for b in bufs:
                b.record_stream(copy_stream)
            st = {'stream': copy_stream, 'bufs': bufs, 'ready': [None, None],
                  'free': [None, None], 'next': 0}
            self._stage[key] = st
        return st


class SyntheticDataSet(_DataSetBase):
    """Counter-based synthetic 4D-STEM data generated on the device (float32 uniform [0,1) or
    uint16 Poisson(3)); value depends only on (flat element index, seed) so any slice can be
    re-created on the host by oracle/synth.py."""

    def __init__(self, shape, dtype, seed, num_partitions=1, sig_dims=2, resident=True):
        self._shape = Shape(tuple(shape), sig_dims=sig_dims)
        self._dtype = np.dtype(dtype)
        self.seed = int(seed)
        self.num_partitions = num_partitions
        self.resident = resident
        self.tileshape = None
        self._cache = {}

    def partition_tensor(self, part, device):
        key = (part.start, part.stop, str(device))
        t = self._cache.get(key)
        if t is None:
            sig = tuple(self._shape.sig)
            per_frame = int(np.prod(sig))
            t = engine.synth_fill((part.stop - part.start,) + sig, self._dtype, self.seed,
                                  device, start=part.start * per_frame)
            if self.resident:
                self._cache[key] = t
        return t

    def materialize(self, device, partitions=None):
        for part in (partitions if partitions is not None else self.get_partitions()):
            self.partition_tensor(part, device)

    def _iter_tiles(self, part, device, roi, tileshape):
        roi_flat = None if roi is None else np.asarray(roi).reshape(-1).astype(bool)
        before = None if roi_flat is None else int(roi_flat[:part.start].sum())
        block = self.partition_tensor(part, device)
        yield from self._emit(block, part.start, part.stop, roi_flat, before, tileshape)
