"""Detector corrections restated from the reference (test infrastructure, see oracle/__init__.py):
``(I - dark) * gain`` then every excluded pixel := mean of its good 3x3 neighbours
(src/libertem/io/corrections/detector.py:18-108 work horse, :111-150 environments,
:156-190 flatten_filter; applied per tile by corrset.py:141-169)."""
import numpy as np


def repair_environments(excluded_mask):
    """for each excluded pixel (row-major order): flat indices of its in-bounds 3x3
    neighbours that are not excluded themselves"""
    sy, sx = excluded_mask.shape
    ys, xs = np.nonzero(excluded_mask)
    out = []
    for y, x in zip(ys, xs):
        env = []
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dy == 0 and dx == 0:
                    continue
                yy, xx = y + dy, x + dx
                if 0 <= yy < sy and 0 <= xx < sx and not excluded_mask[yy, xx]:
                    env.append(yy * sx + xx)
        out.append((y * sx + x, env))
    return out


def correct(data, dark=None, gain=None, excluded_mask=None):
    """data: (..., sy, sx) -> corrected copy in result_type(float32, data)"""
    sig = data.shape[-2:]
    out = data.astype(np.result_type(np.float32, data.dtype)).reshape((-1, sig[0] * sig[1]))
    if dark is not None:
        out = out - dark.reshape(-1)
    if gain is not None:
        out = out * gain.reshape(-1)
    out = np.ascontiguousarray(out, dtype=np.result_type(np.float32, data.dtype))
    if excluded_mask is not None:
        for p, env in repair_environments(excluded_mask):
            if env:
                acc = np.zeros(out.shape[0], dtype=out.dtype)
                for q in env:
                    acc += out[:, q]
                out[:, p] = acc / len(env)
    return out.reshape(data.shape)
