"""Numerical feasibility of the D2-symmetric formulation of the radial-Fourier contraction
(the planned kernel K9).  numpy only, runs on the CPU.

The reference's masks m[b, o](px) = ring_b(r) * exp(i*o*phi) (analysis/radialfourier.py:106-146)
obey, about the default centre (sx/2, sy/2) and away from the centre pixel:
    m(y, sx - x) = (-1)^o conj(m(y, x)),   m(sy - y, x) = conj(m(y, x)),
    m(sy - y, sx - x) = (-1)^o m(y, x).
So for the orbit {p0, p1, p2, p3} of a pixel under the two mirrors, with data I0..I3:
    re += a * (I0 + s I1 + I2 + s I3),   im += b * (I0 - s I1 - I2 + s I3),   s = (-1)^o,
    a + i b = the orbit-averaged weight of p0
-- every real output column needs ONE of the four Hadamard combinations of the four pixels and
ONE real weight per orbit: 4x fewer multiply-adds and a 4x smaller weight table than the direct
sum.  This script measures (1) how far the reference's complex64 masks are from exact symmetry
and (2) the error of the symmetric evaluation (float32 butterflies) against float64."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200.analysis.radialfourier import radial_mask_factory  # noqa: E402
from libertem_b200 import masks  # noqa: E402


def main(S=128, n_bins=8, max_order=24, F=64):
    ro = masks.bounding_radius(S / 2, S / 2, S, S)
    stack = np.asarray(radial_mask_factory(S, S, S / 2, S / 2, 0, ro, n_bins, max_order,
                                           use_sparse=False)()).reshape(-1, S, S)
    M = stack.shape[0]
    c = S // 2
    data = np.random.default_rng(5).random((F, S, S), dtype=np.float32)
    flat = data.reshape(F, -1)
    ref = flat.astype(np.float64) @ stack.reshape(M, -1).astype(np.complex128).T
    scale = (np.abs(flat).astype(np.float64) @ np.abs(stack.reshape(M, -1)).astype(np.float64).T).max()
    # fundamental domain 1 <= y < c, 1 <= x < c and its three mirror images
    q0 = stack[:, 1:c, 1:c].astype(np.complex128)
    q1 = stack[:, 1:c, S - 1:c:-1].astype(np.complex128)          # (y, S - x)
    q2 = stack[:, S - 1:c:-1, 1:c].astype(np.complex128)          # (S - y, x)
    q3 = stack[:, S - 1:c:-1, S - 1:c:-1].astype(np.complex128)
    s = ((-1.0) ** (np.arange(M) % (max_order + 1)))[:, None, None]
    avg = (q0 + s * np.conj(q1) + np.conj(q2) + s * q3) / 4
    dev = max(np.abs(q0 - avg).max(), np.abs(q1 - s * np.conj(avg)).max(),
              np.abs(q2 - np.conj(avg)).max(), np.abs(q3 - s * avg).max())
    print(f'{S}x{S}, {n_bins} bins: max |weight - orbit average| = {dev:.3e}')
    I0, I1 = data[:, 1:c, 1:c], data[:, 1:c, S - 1:c:-1]
    I2, I3 = data[:, S - 1:c:-1, 1:c], data[:, S - 1:c:-1, S - 1:c:-1]
    comb = {(+1, 're'): I0 + I1 + I2 + I3, (-1, 're'): I0 - I1 + I2 - I3,
            (+1, 'im'): I0 - I1 - I2 + I3, (-1, 'im'): I0 + I1 - I2 - I3}      # float32 adds
    comb = {k: v.reshape(F, -1).astype(np.float64) for k, v in comb.items()}
    a, b = avg.real.reshape(M, -1), avg.imag.reshape(M, -1)
    res = np.zeros((F, M), dtype=np.complex128)
    for j in range(M):
        sj = int(s[j, 0, 0])
        res[:, j] = comb[(sj, 're')] @ a[j] + 1j * (comb[(sj, 'im')] @ b[j])
    # rows 0 and c, columns 0 and c have no (or degenerate) orbits: direct sum (0.8 % at 512^2)
    irr = np.zeros((S, S), bool)
    irr[[0, c], :] = True
    irr[:, [0, c]] = True
    res += data[:, irr].astype(np.float64) @ stack[:, irr].astype(np.complex128).T
    print(f'symmetric evaluation vs direct float64: max err / scale = '
          f'{np.abs(res - ref).max() / scale:.3e}  (parity tolerance 1e-5)')


if __name__ == '__main__':
    main(*[int(v) for v in sys.argv[1:]])
