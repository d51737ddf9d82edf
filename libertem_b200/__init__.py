"""libertem_b200 -- B200-native (sm_100a) masked-reduction engine behind LiberTEM's UDF API.

The hot path (ApplyMasksUDF / CoMUDF / SumUDF / SumSigUDF) runs in hand-written CUDA behind the
C ABI declared in ``include/ltb200.h``; this package is the host-side mirror of the reference's
UDF / MaskContainer interface for that path.  There is no CPU fallback: importing the compute
entry points without the built extension raises.
"""
__version__ = '0.1.0'
