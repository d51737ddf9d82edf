"""K10 (dense-walk) correctness on the GPU: radial masks and generic stacks vs numpy float64.
    python scripts/k10_check.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine, group_masks as gm, masks as M  # noqa: E402
from libertem_b200.analysis.radialfourier import radial_mask_factory  # noqa: E402


def check(stack, size, F, tag, **kw):
    dev = torch.device('cuda')
    flat = stack.reshape(stack.shape[0], -1)
    plan = gm.build_plan(flat, size, dev, walk_max_dup=1e9, **kw)
    assert plan.walk is not None, tag
    K = flat.shape[1]
    rng = np.random.default_rng(F)
    data = rng.random((F, K), dtype=np.float32)
    t = torch.from_numpy(data).cuda()
    out = gm.group_masks(t, plan, kernel='walk')
    torch.cuda.synchronize()
    assert engine.last_kernel() == 10
    out = out.cpu().numpy()
    ref = data.astype(np.float64) @ flat.astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(flat).astype(np.float64).T).max() + 1e-30
    err = np.abs(out - ref).max() / scale
    out2 = gm.group_masks(t, plan, out=torch.from_numpy(out).cuda(), accumulate=True,
                          kernel='walk').cpu().numpy()
    err2 = np.abs(out2 - 2 * ref).max() / scale
    out3 = gm.group_masks(t, plan, kernel='walk').cpu().numpy()
    print(f'{tag}: F={F} K={K} groups={plan.n_groups} size={size} segs={plan.walk["n_segments"]} '
          f'err={err:.2e} acc_err={err2:.2e} deterministic={np.array_equal(out, out3)}', flush=True)
    return err < 3e-6 and err2 < 6e-6


def make_stack(n_groups, size, K, seed):
    rng = np.random.default_rng(seed)
    stack = np.zeros((n_groups * size, K), dtype=np.complex64)
    for g in range(n_groups):
        n = int(rng.integers(1, max(2, K // 3)))
        px = np.sort(rng.choice(K, size=n, replace=False))
        vals = (rng.random((size, n)) - 0.5 + 1j * (rng.random((size, n)) - 0.5))
        stack[g * size:(g + 1) * size, px] = vals.astype(np.complex64)
    return stack


ok = True
for S, nb, mo, F in [(64, 4, 6, 128), (128, 8, 24, 300), (128, 8, 24, 1000), (256, 16, 24, 257)]:
    ro = M.bounding_radius(S / 2, S / 2, S, S)
    st = np.asarray(radial_mask_factory(S, S, S / 2, S / 2, 0, ro, nb, mo, use_sparse=False)())
    ok &= check(st.astype(np.complex64), mo + 1, F, f'radial {S}x{S} bins={nb} order={mo}')
for F, K, ng, size in [(64, 512, 3, 25), (100, 1024, 5, 7), (7, 320, 2, 28), (200, 4096, 32, 25),
                       (129, 992, 5, 4), (1000, 4096, 8, 25)]:
    ok &= check(make_stack(ng, size, K, F + K), size, F, 'generic')
print('ALL OK' if ok else 'FAILED')
sys.exit(0 if ok else 1)
