"""COMAnalysis: centre of mass through ApplyMasksUDF with the three CoM masks
(reference src/libertem/analysis/com.py:188-362).  NOTE the result set's ``field`` is ordered
(x, y) whereas CoMUDF's 'field' buffer is (y, x) -- as in the reference."""
import numpy as np

from .. import masks
from ..udf.com import (com_masks_factory, com_masks_generic, center_shifts, apply_correction,
                       magnitude, divergence, curl_2d)
from .base import AnalysisResult, AnalysisResultSet
from .masks import BaseMasksAnalysis


class COMAnalysis(BaseMasksAnalysis):
    def get_parameters(self, parameters):
        # analysis/com.py:310-334: float centre detector/2, radius = inf unless given
        sy, sx = self.dataset.shape.sig
        cx = parameters.get('cx', sx / 2)
        cy = parameters.get('cy', sy / 2)
        r = parameters.get('r', float('inf'))
        ri = parameters.get('ri', 0.0)
        return {'cx': cx, 'cy': cy, 'r': r, 'ri': ri,
                'scan_rotation': parameters.get('scan_rotation', 0.),
                'flip_y': parameters.get('flip_y', False),
                'mask_count': 3, 'mask_dtype': np.float32, 'use_sparse': False}

    def get_mask_factories(self):
        if len(self.dataset.shape.sig) != 2:
            raise ValueError('can only handle 2D signals currently')
        sy, sx = self.dataset.shape.sig
        p = self.parameters
        if p.get('ri'):
            return com_masks_generic(
                detector_y=sy, detector_x=sx,
                base_mask_factory=lambda: masks.ring(
                    imageSizeY=sy, imageSizeX=sx, centerY=p['cy'], centerX=p['cx'],
                    radius=p['r'], radius_inner=p['ri']))
        return com_masks_factory(detector_y=sy, detector_x=sx, cy=p['cy'], cx=p['cx'], r=p['r'])

    def get_udf_results(self, udf_results, roi, damage):
        data = udf_results['intensity'].data
        return self.get_generic_results(data[..., 0], data[..., 1], data[..., 2], damage=damage)

    def get_generic_results(self, img_sum, img_y, img_x, damage):
        p = self.parameters
        y_raw, x_raw = center_shifts(img_sum, img_y, img_x, p['cy'], p['cx'])
        shape = y_raw.shape
        y_c, x_c = apply_correction(y_raw, x_raw, scan_rotation=p['scan_rotation'],
                                    flip_y=p['flip_y'])
        if img_sum.dtype.kind == 'c':
            return AnalysisResultSet([
                AnalysisResult(np.real(x_c), 'x_real'), AnalysisResult(np.real(y_c), 'y_real'),
                AnalysisResult(np.imag(x_c), 'x_imag'), AnalysisResult(np.imag(y_c), 'y_imag'),
            ])
        results = [
            AnalysisResult((x_c, y_c), 'field'),
            AnalysisResult(magnitude(y_c, x_c), 'magnitude'),
            AnalysisResult(x_c, 'x'),
            AnalysisResult(y_c, 'y'),
        ]
        if all(s > 1 for s in shape):
            results[2:2] = [AnalysisResult(divergence(y_c, x_c), 'divergence'),
                            AnalysisResult(curl_2d(y_c, x_c), 'curl')]
        return AnalysisResultSet(results)
