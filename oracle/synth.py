"""Counter-based synthetic 4D-STEM data, bit-identical on host (numpy) and device (CUDA).

Twin of ``libertem_b200/csrc/synth.cuh``.  value(i) depends only on the flat
element index ``i`` (64 bit) and a 32-bit seed, so any slice of any dataset
(even the 256 GiB cfg5) can be regenerated on the host for the oracle.

Test infrastructure (see oracle/__init__.py).
"""
import numpy as np

_M1 = np.uint32(0x85EBCA6B)
_M2 = np.uint32(0xC2B2AE35)
_GOLD = np.uint32(0x9E3779B9)


def _fmix32(h):
    # murmur3 finalizer, uint32 wrap-around arithmetic
    h = h ^ (h >> np.uint32(16))
    h = h * _M1
    h = h ^ (h >> np.uint32(13))
    h = h * _M2
    h = h ^ (h >> np.uint32(16))
    return h


def hash_u32(start, count, seed):
    """uint32 hash of flat indices [start, start+count)."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    lo = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (idx >> np.uint64(32)).astype(np.uint32)
    with np.errstate(over='ignore'):
        h = _fmix32(hi ^ np.uint32(seed & 0xFFFFFFFF))
        h = _fmix32(lo ^ h ^ _GOLD)
    return h


def uniform_f32(start, count, seed):
    """float32 uniform [0, 1): top 24 bits of the hash * 2^-24 (exact)."""
    h = hash_u32(start, count, seed)
    return (h >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def _poisson_thresholds(lam, nmax=24):
    # cumulative Poisson CDF scaled to 2^32, as uint64 thresholds (computed in
    # float64, then frozen as integers: both twins compare integers only)
    import math
    cdf = 0.0
    out = []
    for k in range(nmax):
        cdf += math.exp(-lam) * lam ** k / math.factorial(k)
        out.append(min(int(cdf * 2.0 ** 32), 2 ** 32 - 1))
    return np.array(out, dtype=np.uint64)


#: thresholds for lambda = 3 (cfg3), frozen as integer literals and shared verbatim
#: with synth.cuh (== _poisson_thresholds(3.0)[:20], checked in tests)
POISSON3_THRESHOLDS = np.array([
    213833830, 855335321, 1817587558, 2779839795, 3501528972, 3934542479,
    4151049232, 4243837841, 4278633569, 4290232145, 4293711718, 4294660692,
    4294897936, 4294952684, 4294964416, 4294966763, 4294967203, 4294967280,
    4294967293, 4294967295,
], dtype=np.uint64)


def poisson3_u16(start, count, seed):
    """uint16 Poisson(3)-distributed counts by inverse CDF on the integer hash."""
    h = hash_u32(start, count, seed).astype(np.uint64)
    return np.searchsorted(POISSON3_THRESHOLDS, h, side='right').astype(np.uint16)


def dataset(shape, dtype, seed, start=0):
    n = int(np.prod(shape))
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return uniform_f32(start, n, seed).reshape(shape)
    if dtype == np.uint16:
        return poisson3_u16(start, n, seed).reshape(shape)
    raise ValueError(dtype)
