"""SumSigUDF: per-frame sum over all detector pixels, nav-shaped
(reference src/libertem/udf/sumsigudf.py:6-39) == ApplyMasks with an all-ones mask row."""
import numpy as np
import torch

from .base import UDF
from .masks import as_device_tile
from .. import engine

_ones_cache = {}


def ones_row(k, device):
    key = (int(k), str(device))
    if key not in _ones_cache:
        _ones_cache[key] = torch.ones((1, int(k)), dtype=torch.float32, device=device)
    return _ones_cache[key]


class SumSigUDF(UDF):
    _slab_buffer = 'intensity'     # float32 nav buffer the fused dense kernel may write directly

    def get_result_buffers(self):
        dtype = np.result_type(self.meta.input_dtype, np.float32)
        return {'intensity': self.buffer(kind='nav', dtype=dtype, where='device')}

    def process_tile(self, tile):
        dev = self.meta.device if self.meta.device is not None else torch.device('cuda')
        tile = as_device_tile(tile, dev)
        flat = tile.reshape(tile.shape[0], -1)
        view = self.results.intensity
        if view.dtype == torch.float32:
            engine.masks_dense(flat, ones_row(flat.shape[1], dev), out=view.reshape(-1, 1),
                               accumulate=True)
        else:
            ones = torch.ones((1, flat.shape[1]), dtype=torch.float64, device=dev)
            view[:] += engine.masks_dense(flat, ones).reshape(-1).to(view.dtype)

    def _fused_spec(self):
        if np.result_type(self.meta.input_dtype, np.float32) != np.float32:
            return None
        return {'kind': 'ones', 'buffer': 'intensity'}
