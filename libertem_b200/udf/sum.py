"""SumUDF: sum of all frames, sig-shaped (reference src/libertem/udf/sum.py:6-58)."""
import numpy as np
import torch

from .base import UDF
from .masks import as_device_tile
from .. import engine


class SumUDF(UDF):
    def __init__(self, dtype='float32'):
        super().__init__(dtype=dtype)

    def get_preferred_input_dtype(self):
        return self.params.dtype

    def get_result_buffers(self):
        return {'intensity': self.buffer(kind='sig', dtype=self.meta.input_dtype,
                                         where='device')}

    def process_tile(self, tile):
        dev = self.meta.device if self.meta.device is not None else torch.device('cuda')
        tile = as_device_tile(tile, dev)
        flat = tile.reshape(tile.shape[0], -1)
        view = self.results.intensity
        sig_slice = self.meta.sig_slice
        full = sig_slice.shape.size == self.meta.dataset_shape.sig.size
        if view.dtype == torch.float32 and full and flat.dtype in (
                torch.float32, torch.uint16, torch.uint8, torch.int16, torch.int8):
            # sig_sum[k] += sum_f tile[f, k] inside the library (deterministic column sum)
            empty = torch.empty((0, flat.shape[1]), dtype=torch.float32, device=dev)
            engine.masks_dense(flat, empty, sig_sum=view.reshape(-1))
        else:
            acc = flat.to(view.dtype).sum(dim=0)
            sig = tuple(self.meta.dataset_shape.sig)
            v = view.reshape(sig)[sig_slice.get()]
            v += acc.reshape(v.shape)

    _additive_merge = True

    def merge(self, dest, src):
        dest.intensity[:] += src.intensity

    def merge_all(self, ordered_results):
        chunks = [b.intensity for b in ordered_results.values()]
        return {'intensity': torch.stack(chunks, dim=0).sum(dim=0)}

    def _fused_spec(self):
        if np.dtype(self.meta.input_dtype) != np.float32:
            return None
        return {'kind': 'sig_sum', 'buffer': 'intensity'}
