from .base import AnalysisResult, AnalysisResultSet
from .masks import MasksAnalysis
from .com import COMAnalysis
from .radialfourier import RadialFourierAnalysis

__all__ = ['AnalysisResult', 'AnalysisResultSet', 'MasksAnalysis', 'COMAnalysis',
           'RadialFourierAnalysis']
