"""MasksAnalysis: ApplyMasksUDF behind the analysis interface
(reference src/libertem/analysis/masks.py:6-184)."""
import numpy as np

from ..udf.masks import ApplyMasksUDF
from .base import AnalysisResult, AnalysisResultSet


class BaseMasksAnalysis:
    def __init__(self, dataset, parameters):
        self.dataset = dataset
        self.parameters = self.get_parameters(parameters)

    def get_parameters(self, parameters):
        return parameters

    def get_mask_factories(self):
        raise NotImplementedError()

    def get_use_sparse(self):
        return self.parameters.get('use_sparse', None)

    def get_preferred_dtype(self):
        return self.parameters.get('dtype', None)

    def get_udf(self):
        # analysis/masks.py:18-25
        return ApplyMasksUDF(
            mask_factories=self.get_mask_factories(),
            use_sparse=self.get_use_sparse(),
            mask_count=self.parameters.get('mask_count'),
            mask_dtype=self.parameters.get('mask_dtype'),
            preferred_dtype=self.get_preferred_dtype(),
        )


class MasksAnalysis(BaseMasksAnalysis):
    def get_mask_factories(self):
        return self.parameters['factories']

    def get_udf_results(self, udf_results, roi, damage):
        data = udf_results['intensity'].data
        results = []
        for i in range(data.shape[-1]):
            results.append(AnalysisResult(raw_data=data[..., i], key=f'mask_{i}',
                                          title=f'mask {i}',
                                          desc=f'integrated intensity for mask {i}'))
        # reference exposes the stack as mask_0 = (count, *nav) too (analysis/masks.py:120-184)
        return AnalysisResultSet(results, raw_results=np.moveaxis(data, -1, 0))
