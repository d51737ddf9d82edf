"""CoMUDF on the B200 engine: centre-of-mass of every frame + field post-processing.

Mirrors the reference ``libertem.udf.com`` (src/libertem/udf/com.py): same parameters
(``CoMParams`` / ``CoMUDF.with_params``), same result buffers, same post-processing semantics.
The per-frame moments (m00, m10, m01) are three mask rows [D, y*D, x*D] of the dense kernel
(com.py:47-97,534-582), normally riding in the same pass as ApplyMasksUDF's masks (fused
runner).  The nav-sized post-processing (com.py:650-717) runs in the nav-space kernels (K9,
``libertem_b200/nav.py``) when the moments are in HBM -- float32 shifts, float64 from the
rotation matrix on, rounded once, like the reference's dtypes; the module-level helpers below
(``center_shifts``, ``apply_correction``, ``guess_corrections`` ...) keep the reference's numpy
call signatures for host arrays, ``libertem_b200.nav.guess_corrections`` is the device form.
"""
from enum import IntEnum
from typing import NamedTuple, Union

import numpy as np
import torch

from .. import masks
from ..common.container import MaskContainer
from .base import UDF
from .masks import ApplyMasksEngine, as_device_tile


class RegressionOptions(IntEnum):
    NO_REGRESSION = -1
    SUBTRACT_MEAN = 0
    SUBTRACT_LINEAR = 1


class CoMParams(NamedTuple):
    cy: Union[float, None] = None
    cx: Union[float, None] = None
    r: float = float('inf')
    ri: Union[float, None] = 0.
    scan_rotation: float = 0.
    flip_y: bool = False
    regression: object = RegressionOptions.NO_REGRESSION


# ---- coordinate transforms (reference corrections/coordinates.py:11-37) ----------------------

def rotate_deg(degrees):
    rad = np.pi * degrees / 180
    return np.array([(np.cos(rad), np.sin(rad)), (-np.sin(rad), np.cos(rad))])


def flip_y_matrix():
    return np.array([(-1, 0), (0, 1)])


def identity():
    return np.eye(2)


# ---- mask factories ---------------------------------------------------------------------------

def com_masks_factory(detector_y, detector_x, cy, cx, r):
    """[disk, y*disk, x*disk] (com.py:47-66)"""
    def disk_mask():
        return masks.circular(centerX=cx, centerY=cy, imageSizeX=detector_x,
                              imageSizeY=detector_y, radius=r)
    return [
        disk_mask,
        lambda: masks.gradient_y(imageSizeX=detector_x, imageSizeY=detector_y) * disk_mask(),
        lambda: masks.gradient_x(imageSizeX=detector_x, imageSizeY=detector_y) * disk_mask(),
    ]


def com_masks_generic(detector_y, detector_x, base_mask_factory):
    """[B, y*B, x*B] for an arbitrary selection mask B (com.py:69-97)"""
    return [
        base_mask_factory,
        lambda: masks.gradient_y(imageSizeX=detector_x, imageSizeY=detector_y)
        * base_mask_factory(),
        lambda: masks.gradient_x(imageSizeX=detector_x, imageSizeY=detector_y)
        * base_mask_factory(),
    ]


# ---- nav-space math (com.py:100-142) ------------------------------------------------------------

def center_shifts(img_sum, img_y, img_x, ref_y, ref_x):
    x_centers = np.divide(img_x, img_sum, where=img_sum != 0)
    y_centers = np.divide(img_y, img_sum, where=img_sum != 0)
    x_centers[img_sum == 0] = ref_x
    y_centers[img_sum == 0] = ref_y
    x_centers -= ref_x
    y_centers -= ref_y
    return (y_centers, x_centers)


def apply_correction(y_centers, x_centers, scan_rotation, flip_y, forward=True):
    shape = y_centers.shape
    transform = flip_y_matrix() if flip_y else identity()
    transform = rotate_deg(scan_rotation) @ transform
    if not forward:
        transform = np.linalg.inv(transform)
    y_t, x_t = transform @ (y_centers.reshape(-1), x_centers.reshape(-1))
    return (y_t.reshape(shape), x_t.reshape(shape))


def divergence(y_centers, x_centers):
    return np.gradient(y_centers, axis=0) + np.gradient(x_centers, axis=1)


def curl_2d(y_centers, x_centers):
    return np.gradient(y_centers, axis=1) - np.gradient(x_centers, axis=0)


def magnitude(y_centers, x_centers):
    return np.sqrt(y_centers ** 2 + x_centers ** 2)


def coordinate_check(y_centers, x_centers, roi=None):
    """RMS curl for every scan rotation 0..359 deg, straight and flipped (com.py:145-190)"""
    straight = np.zeros(360)
    flipped = np.zeros(360)
    if roi is None:
        roi = (slice(0, -1), slice(0, -1))
    for angle in range(360):
        for flip in (True, False):
            y_t, x_t = apply_correction(y_centers, x_centers, scan_rotation=angle, flip_y=flip)
            rms = np.sqrt(np.mean(curl_2d(y_t, x_t)[roi] ** 2))
            (flipped if flip else straight)[angle] = rms
    return (straight, flipped)


class GuessResult(NamedTuple):
    scan_rotation: int
    flip_y: bool
    cy: float
    cx: float


def guess_corrections(y_centers, x_centers, roi=None):
    """Guess (cy, cx), scan_rotation and flip_y from CoM data by minimising the RMS curl and
    checking the divergence polarity (com.py:208-295)."""
    if roi is None:
        roi = (slice(0, -1), slice(0, -1))
    straight, flipped = coordinate_check(y_centers, x_centers, roi=roi)
    flip = bool(np.min(flipped) < np.min(straight))
    angle = np.argmin(flipped) if flip else np.argmin(straight)
    cy_, cx_ = apply_correction(y_centers, x_centers, scan_rotation=angle, flip_y=flip)
    div = divergence(cy_, cx_)[roi]
    all_range = np.maximum(-np.min(div), np.max(div))
    hist, _ = np.histogram(div, range=(-all_range, all_range), bins=5)
    if np.sum(hist[:1]) < np.sum(hist[-1:]):
        angle += 180
    if angle > 180:
        angle -= 360
    return GuessResult(scan_rotation=int(angle), flip_y=flip,
                       cy=np.mean(y_centers[roi]), cx=np.mean(x_centers[roi]))


class CoMUDF(UDF):
    """Centre-of-mass analysis; result buffers (all nav): raw_com, raw_shifts, field (y, x),
    field_y, field_x, magnitude, divergence, curl, plus 'regression' (3, 2)
    (com.py:298-510)."""

    _slab_buffer = 'raw_mask_result'     # float32 nav buffer the fused dense kernel may write directly

    def __init__(self, com_params: CoMParams = CoMParams()):
        super().__init__(com_params=com_params)
        self._containers = {}

    def copy_for_partition(self):
        new = super().copy_for_partition()
        new._containers = self._containers      # share the cached CoM mask stack
        return new

    @classmethod
    def with_params(cls, *, cy=None, cx=None, r=float('inf'), ri=0., scan_rotation=0.,
                    flip_y=False, regression=RegressionOptions.NO_REGRESSION):
        if ri is not None and ri >= r:
            raise ValueError('Inner radius must be less than outer radius for annular CoM')
        return cls(com_params=CoMParams(cy=cy, cx=cx, r=r, ri=ri, scan_rotation=scan_rotation,
                                        flip_y=flip_y, regression=regression))

    def get_result_buffers(self):
        dtype = np.result_type(self.meta.input_dtype, np.float32)
        nav2 = dict(kind='nav', dtype=dtype, extra_shape=(2,), use='result_only')
        nav0 = dict(kind='nav', dtype=dtype, use='result_only')
        return {
            'raw_mask_result': self.buffer(kind='nav', dtype=dtype, extra_shape=(3,),
                                           where='device', use='private'),
            'raw_com': self.buffer(**nav2), 'raw_shifts': self.buffer(**nav2),
            'field': self.buffer(**nav2),
            'field_y': self.buffer(**nav0), 'field_x': self.buffer(**nav0),
            'magnitude': self.buffer(**nav0), 'divergence': self.buffer(**nav0),
            'curl': self.buffer(**nav0),
            'regression': self.buffer(kind='single', extra_shape=(3, 2), dtype=np.float64,
                                      use='result_only'),
        }

    def get_params(self) -> CoMParams:
        sig_shape = tuple(self.meta.dataset_shape.sig)
        p = self.params.com_params
        cy = sig_shape[0] // 2 if p.cy is None else p.cy
        cx = sig_shape[1] // 2 if p.cx is None else p.cx
        return CoMParams(cy=cy, cx=cx, r=p.r, ri=p.ri, scan_rotation=p.scan_rotation,
                         flip_y=p.flip_y, regression=p.regression)

    def get_task_data(self):
        sig_shape = tuple(self.meta.dataset_shape.sig)
        p = self.get_params()
        if len(sig_shape) != 2:
            raise ValueError('CoMUDF only works with 2D sig shape.')
        if len(self.meta.dataset_shape.nav) != 2:
            raise ValueError('CoMUDF only works with 2D nav shape.')
        if p.ri is None or np.isclose(p.ri, 0.):
            fac = com_masks_factory(detector_y=sig_shape[0], detector_x=sig_shape[1],
                                    cx=p.cx, cy=p.cy, r=p.r)
        else:
            fac = com_masks_generic(
                detector_y=sig_shape[0], detector_x=sig_shape[1],
                base_mask_factory=lambda: masks.ring(
                    imageSizeY=sig_shape[0], imageSizeX=sig_shape[1], centerY=p.cy,
                    centerX=p.cx, radius=p.r, radius_inner=p.ri))
        key = (sig_shape, p.cy, p.cx, p.r, p.ri)
        container = self._containers.get(key)
        if container is None:
            container = MaskContainer(mask_factories=fac, dtype=np.float32, use_sparse=False,
                                      count=3, backend='cuda')
            self._containers[key] = container
        return {'com_params': p,
                'engine': ApplyMasksEngine(masks=container, meta=self.meta, use_torch=True)}

    def process_tile(self, tile):
        eng = self.task_data['engine']
        view = self.results.raw_mask_result
        tile = as_device_tile(tile, eng.device)
        flat = tile.reshape(tile.shape[0], -1)
        if view.dtype == torch.float32 and view.is_cuda:
            eng.process_flat(flat, out=view, accumulate=True)
        else:
            view[:] += self.forbuf(eng.process_tile(tile), view)

    def _fused_spec(self):
        eng = self.task_data['engine']
        if eng.compute != np.float32:
            return None
        return {'kind': 'dense', 'buffer': 'raw_mask_result', 'engine': eng, 'columns': 3}

    # -- results (com.py:584-717) -------------------------------------------------------------
    def get_field_results(self, field_y, field_x):
        return {'magnitude': magnitude(y_centers=field_y, x_centers=field_x),
                'divergence': divergence(y_centers=field_y, x_centers=field_x),
                'curl': curl_2d(y_centers=field_y, x_centers=field_x)}

    def get_regression(self, field, valid_mask):
        inp = None
        result = np.zeros((3, 2))
        reg = self.get_params().regression

        def get_inp():
            a = np.ones(field.shape[:-1] + (3,))
            y, x = np.ogrid[:field.shape[0], :field.shape[1]]
            a[..., 1] = y
            a[..., 2] = x
            return a

        if isinstance(reg, (int, np.integer)):
            if reg == -1:
                pass
            elif reg == 0:
                result[0] = np.mean(field[valid_mask], axis=0)
            elif reg == 1:
                inp = get_inp()
                result[:] = np.linalg.lstsq(inp[valid_mask], field[valid_mask], rcond=None)[0]
            else:
                raise ValueError(f'Unrecognized regression option {reg}')
        else:
            reg = np.array(reg)
            if reg.shape != (3, 2):
                raise ValueError(f"Regression parameter {reg} doesn't have required shape (3, 2).")
            result[:] = reg
        has_lin = not np.allclose(result[1:], 0)
        if has_lin and inp is None:
            inp = get_inp()
        if not has_lin:
            inp = None
        return result, inp

    def apply_mean_regression(self, regression, field_inout, valid_mask):
        field_inout[valid_mask] -= regression[0]

    def apply_lin_regression(self, regression, inp, field_inout, valid_mask):
        field_inout[valid_mask] -= inp[valid_mask] @ regression

    def _get_results_device(self, raw):
        """the same pipeline in the nav-space kernels (libertem_b200/nav.py, csrc/k9_nav.cu)
        when the moments live in HBM: one upload of the roi / valid maps, four small launches,
        one D2H of the finished float32 buffers"""
        from .. import nav
        p = self.get_params()
        nav_shape = tuple(self.meta.dataset_shape.nav)
        n = int(np.prod(nav_shape))
        roi = self.meta.roi
        row_of_nav = None
        if roi is not None:
            flat = np.asarray(roi).reshape(-1).astype(bool)
            row_of_nav = np.where(flat, np.cumsum(flat) - 1, -1).astype(np.int32)
        valid = self.meta.get_valid_nav_mask(full_nav=True).reshape(-1)
        reg = p.regression
        coeffs = None
        if isinstance(reg, (int, np.integer)):
            if reg not in (-1, 0, 1):
                raise ValueError(f'Unrecognized regression option {reg}')
            mode = int(reg)
        else:
            coeffs = np.array(reg, dtype=np.float64)
            if coeffs.shape != (3, 2):
                raise ValueError(
                    f"Regression parameter {reg} doesn't have required shape (3, 2).")
            mode = 2
        transform = rotate_deg(p.scan_rotation) @ (flip_y_matrix() if p.flip_y else identity())
        out = nav.com_postprocess(raw, nav_shape, p.cy, p.cx, transform, mode, coeffs,
                                  row_of_nav=row_of_nav, valid=valid)
        results = {}
        sel = None if roi is None else torch.from_numpy(
            np.nonzero(np.asarray(roi).reshape(-1))[0]).to(raw.device)
        for key, t in out.items():
            if key == 'regression':
                results[key] = t.cpu().numpy()
                continue
            if sel is not None:
                t = t.index_select(0, sel)
            arr = t.cpu().numpy()
            results[key] = arr if (sel is not None or arr.ndim == 2) else arr.reshape(n, 1)
        return results

    def get_results(self):
        raw_buf = self.results.get_buffer('raw_mask_result').tensor
        nav_shape = tuple(self.meta.dataset_shape.nav)
        if (isinstance(raw_buf, torch.Tensor) and raw_buf.is_cuda
                and raw_buf.dtype == torch.float32 and len(nav_shape) == 2
                and min(nav_shape) >= 2):
            return self._get_results_device(raw_buf)
        # host tensors (and complex moments): numpy, operation for operation as the reference
        p = self.get_params()
        rmr = self.results.get_buffer('raw_mask_result').data     # (*nav, 3), NaN off-roi
        raw_shifts = center_shifts(img_sum=rmr[..., 0], img_y=rmr[..., 1], img_x=rmr[..., 2],
                                   ref_y=p.cy, ref_x=p.cx)
        raw_com = (raw_shifts[0].copy() + p.cy, raw_shifts[1].copy() + p.cx)
        field = apply_correction(y_centers=raw_shifts[0], x_centers=raw_shifts[1],
                                 scan_rotation=p.scan_rotation, flip_y=p.flip_y)
        roi = self.meta.roi
        raw_shifts = np.moveaxis(np.array(raw_shifts), 0, -1)
        raw_com = np.moveaxis(np.array(raw_com), 0, -1)
        field = np.moveaxis(np.array(field), 0, -1)
        nav_size = self.meta.dataset_shape.nav.size
        valid_mask = self.meta.get_valid_nav_mask(full_nav=True).reshape(
            tuple(self.meta.dataset_shape.nav))
        regression, inp = self.get_regression(field, valid_mask=valid_mask)
        if inp is not None:
            self.apply_lin_regression(regression, inp, field, valid_mask)
        elif not np.allclose(regression[0], 0):
            self.apply_mean_regression(regression, field, valid_mask)
        results = {'raw_shifts': raw_shifts, 'raw_com': raw_com, 'field': field,
                   'field_y': field[..., 0], 'field_x': field[..., 1],
                   'regression': regression.astype(np.float64)}
        results.update(self.get_field_results(field_y=field[..., 0], field_x=field[..., 1]))
        decl = self.get_result_buffers()
        for key, buf in decl.items():
            if buf.kind == 'nav' and key in results:
                if roi is not None:
                    results[key] = results[key][np.asarray(roi).reshape(
                        tuple(self.meta.dataset_shape.nav)).astype(bool)]
                else:
                    results[key] = results[key].reshape((nav_size, -1))
        return results
