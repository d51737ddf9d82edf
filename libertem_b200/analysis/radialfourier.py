"""RadialFourierAnalysis: Fourier coefficients of ring-shaped detector regions,
``c[b, o] = sum_pixels ring_b(r) * exp(i*o*phi) * I`` for b < n_bins, o <= max_order
(reference src/libertem/analysis/radialfourier.py:106-146 mask factory, :184-194 result
layout, :316-354 parameter defaults).  The masks are complex64; on real data they run through
the dense kernel as interleaved (re, im) rows -- see ApplyMasksEngine."""
import numpy as np

from .. import masks
from .base import AnalysisResult, AnalysisResultSet
from .masks import BaseMasksAnalysis


def radial_mask_factory(detector_y, detector_x, cx, cy, ri, ro, n_bins, max_order, use_sparse,
                        dtype=np.complex64):
    dtype = np.result_type(dtype, np.complex64)

    def stack():
        rings = masks.radial_bins(centerX=cx, centerY=cy, imageSizeX=detector_x,
                                  imageSizeY=detector_y, radius=ro, radius_inner=ri,
                                  n_bins=n_bins, use_sparse=False, dtype=dtype)
        orders = np.arange(max_order + 1, dtype=dtype)
        _, phi = masks.polar_map(centerX=cx, centerY=cy, imageSizeX=detector_x,
                                 imageSizeY=detector_y)
        modulator = np.exp(phi.astype(dtype) * orders[:, np.newaxis, np.newaxis] * 1j)
        ring_stack = (rings[:, np.newaxis, ...] * modulator).reshape(
            (-1, detector_y, detector_x))
        if use_sparse:
            return masks.SparseStack.from_dense(ring_stack)
        return ring_stack
    return stack


class RadialFourierAnalysis(BaseMasksAnalysis):
    def get_parameters(self, parameters):
        sy, sx = self.dataset.shape.sig
        cx = parameters.get('cx', sx / 2)
        cy = parameters.get('cy', sy / 2)
        ri = parameters.get('ri', 0)
        ro = parameters.get('ro', masks.bounding_radius(cx, cy, sx, sy))
        n_bins = parameters.get('n_bins', 1)
        max_order = parameters.get('max_order', 24)
        mask_count = n_bins * (max_order + 1)
        bin_width = (ro - ri) / n_bins
        bin_area = np.pi * ro ** 2 - np.pi * (ro - bin_width) ** 2
        stack_size = mask_count * sy * sx * 8
        default = 'scipy.sparse'
        if stack_size < 2 ** 18:
            default = False
        elif bin_area / (sx * sy) > 0.05 and n_bins < 10:
            default = False
        return {'cx': cx, 'cy': cy, 'ri': ri, 'ro': ro, 'n_bins': n_bins,
                'max_order': max_order, 'use_sparse': parameters.get('use_sparse', default),
                'mask_count': mask_count, 'mask_dtype': np.complex64}

    def get_mask_factories(self):
        if len(self.dataset.shape.sig) != 2:
            raise ValueError('can only handle 2D signals currently')
        sy, sx = self.dataset.shape.sig
        p = self.parameters
        return radial_mask_factory(detector_y=sy, detector_x=sx, cx=p['cx'], cy=p['cy'],
                                   ri=p['ri'], ro=p['ro'], n_bins=p['n_bins'],
                                   max_order=p['max_order'], use_sparse=p['use_sparse'])

    def get_udf_results(self, udf_results, roi, damage):
        shape = tuple(self.dataset.shape.nav)
        n = int(np.prod(shape))
        # NOTE transposed reshape, as in the reference (:189-194)
        raw = udf_results['intensity'].data.reshape((n, -1)).T
        orders = self.parameters['max_order'] + 1
        n_bins = self.parameters['n_bins']
        raw = raw.reshape((n_bins, orders, *shape))

        def resultlist():
            # derived channels, same keys / raw_data as the reference (:196-295)
            out = []
            absolute = np.absolute(raw)
            normal = np.maximum(1, absolute[:, 0])
            angle = np.angle(raw)
            higher = absolute[:, 1:, ...]
            threshold = higher.reshape((n_bins, -1)).max(axis=1) * 0.2 if orders > 1 else None
            if orders > 1:
                expand = (slice(None),) + (np.newaxis,) * (higher.ndim - 1)
                below = np.all(higher < threshold[expand], axis=1)
                dominant = np.argmax(higher, axis=1) + 1
                dominant[below] = 0
            else:
                dominant = np.zeros((n_bins,) + shape, dtype=np.int64)
            for b in range(n_bins):
                out.append(AnalysisResult(dominant[b], 'dominant_%s' % b))
                for o in range(orders):
                    out.append(AnalysisResult(absolute[b, o], f'absolute_{b}_{o}'))
            for b in range(n_bins):
                for o in range(orders):
                    out.append(AnalysisResult(angle[b, o], f'phase_{b}_{o}'))
            for b in range(n_bins):
                out.append(AnalysisResult(raw[b, 0], f'complex_{b}_{0}'))
                for o in range(1, orders):
                    out.append(AnalysisResult(raw[b, o] / normal[b], f'complex_{b}_{o}'))
            return out
        return AnalysisResultSet(resultlist, raw_results=raw)
