from .base import UDF, UDFMeta, UDFException
from .masks import ApplyMasksUDF
from .com import CoMUDF, CoMParams, RegressionOptions, guess_corrections
from .sum import SumUDF
from .sumsigudf import SumSigUDF

__all__ = ['UDF', 'UDFMeta', 'UDFException', 'ApplyMasksUDF', 'CoMUDF', 'CoMParams',
           'RegressionOptions', 'guess_corrections', 'SumUDF', 'SumSigUDF']
