"""Small Context mirroring the reference's entry points for the hot path
(src/libertem/api.py: Context.run_udf :914-1051, create_com_analysis :592-663,
create_mask_analysis :514-590, create_radial_fourier_analysis :665-707)."""
import torch

from .runner import run_udf as _run_udf, UDFRunner
from .io.memory import MemoryDataSet


class Context:
    def __init__(self, device=None):
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)

    def load(self, filetype, *args, **kwargs):
        if filetype not in ('memory', 'mem'):
            raise ValueError('only in-memory datasets are in scope of this runtime')
        return MemoryDataSet(*args, **kwargs)

    def run_udf(self, dataset, udf, roi=None, progress=False, corrections=None, backends=None):
        return _run_udf(dataset, udf, roi=roi, device=self.device, corrections=corrections)

    def run(self, analysis, roi=None):
        udf = analysis.get_udf()
        runner = UDFRunner([udf])
        res = runner.run_for_dataset(analysis.dataset, roi=roi, device=self.device)
        return analysis.get_udf_results(res.buffers[0], roi, res.damage)

    def create_com_analysis(self, dataset, cx=None, cy=None, mask_radius=None, flip_y=False,
                            scan_rotation=0.0, mask_radius_inner=None):
        from .analysis.com import COMAnalysis
        params = dict(flip_y=flip_y, scan_rotation=scan_rotation)
        if cx is not None:
            params['cx'] = cx
        if cy is not None:
            params['cy'] = cy
        if mask_radius is not None:
            params['r'] = mask_radius
        if mask_radius_inner is not None:
            params['ri'] = mask_radius_inner
        return COMAnalysis(dataset, params)

    def create_mask_analysis(self, dataset, factories, use_sparse=None, mask_count=None,
                             mask_dtype=None, dtype=None):
        from .analysis.masks import MasksAnalysis
        return MasksAnalysis(dataset, dict(factories=factories, use_sparse=use_sparse,
                                           mask_count=mask_count, mask_dtype=mask_dtype,
                                           dtype=dtype))

    def create_radial_fourier_analysis(self, dataset, cx=None, cy=None, ri=None, ro=None,
                                       n_bins=None, max_order=None, use_sparse=None):
        from .analysis.radialfourier import RadialFourierAnalysis
        params = {k: v for k, v in dict(cx=cx, cy=cy, ri=ri, ro=ro, n_bins=n_bins,
                                        max_order=max_order, use_sparse=use_sparse).items()
                  if v is not None}
        return RadialFourierAnalysis(dataset, params)
