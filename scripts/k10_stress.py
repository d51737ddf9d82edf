"""K10 determinism stress: repeat one launch many times, report how often the result differs
from the float64 reference by more than the tolerance.   python scripts/k10_stress.py [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import group_masks as gm, masks as M  # noqa: E402
from libertem_b200.analysis.radialfourier import radial_mask_factory  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device('cuda')
bad_total = 0
CASES = [(128, 8, 24, 300), (256, 16, 24, 257), (128, 8, 24, 1000), (128, 16, 6, 2000),
         (128, 11, 24, 1000), (128, 8, 6, 2000), (128, 16, 24, 1000), (256, 16, 6, 2000),
         (128, 16, 6, 1000), (128, 16, 24, 2000), (128, 16, 5, 2000), (128, 16, 7, 2000),
         (512, 32, 24, 1024), (256, 16, 7, 3000), (128, 8, 7, 4000), (128, 8, 11, 2000)]
if len(sys.argv) > 2:
    CASES = [CASES[int(a)] for a in sys.argv[2:]]
for S, nb, mo, F in CASES:
    ro = M.bounding_radius(S / 2, S / 2, S, S)
    st = np.asarray(radial_mask_factory(S, S, S / 2, S / 2, 0, ro, nb, mo, use_sparse=False)())
    flat = st.reshape(st.shape[0], -1).astype(np.complex64)
    plan = gm.build_plan(flat, mo + 1, dev, walk_max_dup=1e9)
    rng = np.random.default_rng(F)
    data = rng.random((F, S * S), dtype=np.float32)
    t = torch.from_numpy(data).cuda()
    ref = torch.from_numpy((data.astype(np.float64) @ flat.astype(np.complex128).T)).cuda()
    scale = float((np.abs(data).astype(np.float64) @ np.abs(flat).astype(np.float64).T).max())
    bad = 0
    worst = 0.0
    for r in range(reps):
        acc = (r % 2) == 1
        base = torch.zeros((F, flat.shape[0]), dtype=torch.complex64, device=dev)
        out = gm.group_masks(t, plan, out=base, accumulate=acc, kernel='walk')
        err = float((out.to(torch.complex128) - ref).abs().max()) / scale
        worst = max(worst, err)
        bad += err > 3e-6
        if err > 3e-6 and bad <= 6:
            d = (out.to(torch.complex128) - ref).abs() / scale
            rows, cols = torch.nonzero(d > 3e-6, as_tuple=True)
            gs = sorted(set((cols // (mo + 1)).tolist()))
            fbs = sorted(set((rows // 128).tolist()))
            print(f'  rep {r} acc={acc}: err {err:.2e} rings {gs} frame blocks {fbs} '
                  f'orders {sorted(set((cols % (mo + 1)).tolist()))[:8]} n={len(rows)} '
                  f'rows {sorted(set(rows.tolist()))[:6]}', flush=True)
            r0, c0 = int(rows[0]), int(cols[0])
            print('    got', complex(out[r0, c0]), 'ref', complex(ref[r0, c0]), flush=True)
    print(f'{S}x{S} bins={nb} order={mo} F={F}: {bad}/{reps} bad, worst err {worst:.2e}', flush=True)
    bad_total += bad
print('STRESS OK' if bad_total == 0 else 'STRESS FAILED')
sys.exit(0 if bad_total == 0 else 1)
