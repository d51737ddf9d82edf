"""Detector corrections fused into the hot path.

Mirrors the reference's ``CorrectionSet`` (src/libertem/io/corrections/corrset.py:72-204):
dark frame subtraction, gain map multiplication and patching of excluded (dead/hot) pixels
with the mean of their good 3x3 neighbours (io/corrections/detector.py:18-108,111-190), applied
to every tile before the UDFs see it.

The corrected frame is an affine function of the raw frame, ``c = A @ I + b``.  For the masked
reductions ``sum_p m[p] * c[p]`` the correction therefore folds into the masks once per run
(the reference has the same idea for gain + dead pixels: ``correct_dot_masks``,
detector.py:315-340):

    m' = A^T m          (computed on the host in float64, uploaded as float32 rows)
    result = I . m' + m . b   (the constant is added to the kernel's output columns)

so the frames are still read exactly once, in their native dtype, with zero extra per-pixel
work.  SumUDF's frame sum is corrected after the fact (``A @ sum + n_frames * b``).  UDFs that
are not fused get explicitly corrected tiles (``apply``), like in the reference.
"""
import numpy as np
import torch


class RepairValueError(ValueError):
    pass


class CorrectionSet:
    def __init__(self, dark=None, gain=None, excluded_pixels=None, allow_empty=False):
        self._dark = None if dark is None else np.asarray(dark)
        self._gain = None if gain is None else np.asarray(gain)
        if excluded_pixels is not None:
            if hasattr(excluded_pixels, 'todense'):
                excluded_pixels = excluded_pixels.todense()
            elif hasattr(excluded_pixels, 'toarray'):
                excluded_pixels = excluded_pixels.toarray()
            excluded_pixels = np.asarray(excluded_pixels) != 0
        self._excluded = excluded_pixels
        self._allow_empty = allow_empty
        self._env = None
        self._dev = {}
        if excluded_pixels is not None and not allow_empty:
            empty = [p for p, env in self.repair_environments() if not env]
            if empty:
                raise RepairValueError(f'Empty repair environments for pixel(s) number {empty}.')

    def get_dark_frame(self):
        return self._dark

    def get_gain_map(self):
        return self._gain

    def get_excluded_pixels(self):
        return self._excluded

    def have_corrections(self):
        return any(c is not None for c in (self._dark, self._gain, self._excluded))

    def repair_environments(self):
        """[(flat index of excluded pixel, [flat indices of its good 3x3 neighbours])]"""
        if self._env is None:
            env_list = []
            if self._excluded is not None:
                ex = self._excluded
                sy, sx = ex.shape
                for y, x in zip(*np.nonzero(ex)):
                    env = [(y + dy) * sx + (x + dx)
                           for dy in (-1, 0, 1) for dx in (-1, 0, 1)
                           if (dy or dx) and 0 <= y + dy < sy and 0 <= x + dx < sx
                           and not ex[y + dy, x + dx]]
                    env_list.append((int(y * sx + x), env))
            self._env = env_list
        return self._env

    # -- folding into linear functionals -------------------------------------------------
    def fold_masks(self, rows):
        """rows: (M, K) mask rows -> (rows', const) with  rows @ corrected == rows' @ raw + const"""
        rows = np.asarray(rows, dtype=np.float64)
        K = rows.shape[1]
        eff = rows.copy()
        for p, env in self.repair_environments():
            w = eff[:, p].copy()
            eff[:, p] = 0
            if env:
                eff[:, env] += (w / len(env))[:, None]
            else:
                eff[:, p] = w        # allow_empty: pixel left uncorrected
        gain = np.ones(K) if self._gain is None else self._gain.reshape(-1).astype(np.float64)
        dark = np.zeros(K) if self._dark is None else self._dark.reshape(-1).astype(np.float64)
        folded = eff * gain
        const = -(folded @ dark)
        return folded, const

    def correct_frame_sum(self, sig_sum, n_frames):
        """corrected sum over frames from the raw sum (float32 (K,) tensor or array)"""
        s = np.asarray(sig_sum, dtype=np.float64).reshape(-1)
        if self._dark is not None:
            s = s - n_frames * self._dark.reshape(-1).astype(np.float64)
        if self._gain is not None:
            s = s * self._gain.reshape(-1).astype(np.float64)
        for p, env in self.repair_environments():
            if env:
                s[p] = s[env].sum() / len(env)
        return s

    # -- explicit application (non-fused UDFs) ---------------------------------------------
    def apply(self, tile):
        """corrected copy of a device tile (frames, sy, sx) in float32/float64"""
        dev = tile.device
        dt = torch.float64 if tile.dtype in (torch.float64, torch.int32, torch.int64,
                                             torch.uint32) else torch.float32
        out = tile.to(dt).reshape(tile.shape[0], -1).clone()
        key = (str(dev), dt)
        if key not in self._dev:
            dark = None if self._dark is None else torch.from_numpy(
                self._dark.reshape(-1).astype(np.float64)).to(dev, dt)
            gain = None if self._gain is None else torch.from_numpy(
                self._gain.reshape(-1).astype(np.float64)).to(dev, dt)
            self._dev[key] = (dark, gain)
        dark, gain = self._dev[key]
        if dark is not None:
            out -= dark
        if gain is not None:
            out *= gain
        for p, env in self.repair_environments():
            if env:
                out[:, p] = out[:, env].sum(dim=1) / len(env)
        return out.reshape(tile.shape)
