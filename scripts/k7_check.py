"""Bring-up / timing of the tensor-core group-sparse kernel (K7) on the cfg4 geometry:
512x512 signal, radial Fourier 32 bins x 25 orders.  python scripts/k7_check.py [frames]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine, group_masks as gm  # noqa: E402
from libertem_b200.analysis.radialfourier import radial_mask_factory  # noqa: E402


def bench(fn, n=4):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    dev = torch.device('cuda')
    # small structured check first
    rng = np.random.default_rng(1)
    K, ng, size = 2048, 3, 25
    stack = np.zeros((ng * size, K), np.complex64)
    for g in range(ng):
        px = np.sort(rng.choice(K, 300 + 50 * g, replace=False))
        stack[g * size:(g + 1) * size, px] = (rng.random((size, len(px))) - 0.5 +
                                              1j * (rng.random((size, len(px))) - 0.5))
    plan = gm.build_plan(stack, size, dev)
    data = rng.random((300, K), dtype=np.float32)
    t = torch.from_numpy(data).cuda()
    ref = data.astype(np.float64) @ stack.astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(stack).astype(np.float64).T).max()
    for chain in (1, 2, 4, 8):
        out = gm.group_masks(t, plan, kernel='tc', chain=chain).cpu().numpy()
        print(f'small check chain={chain}: max err / scale = {np.abs(out - ref).max() / scale:.3e}',
              flush=True)
    out = gm.group_masks(t, plan, kernel='ffma').cpu().numpy()
    print(f'small check ffma: max err / scale = {np.abs(out - ref).max() / scale:.3e}', flush=True)

    t0 = time.time()
    fac = radial_mask_factory(512, 512, 256, 256, 0, 364.0, 32, 24, use_sparse=False)
    stack = np.asarray(fac()).reshape(800, -1)
    plan = gm.build_plan(stack, 25, dev)
    print(f'plan: {plan.n_entries} entries, table_split {tuple(plan.table_split.shape)}, '
          f'{time.time() - t0:.1f} s', flush=True)
    data = engine.synth_fill((F, 512 * 512), np.float32, 104, dev)
    gb = F * 512 * 512 * 4 / 1e9
    out_tc = gm.group_masks(data, plan, kernel='tc')
    out_ff = gm.group_masks(data, plan, kernel='ffma')
    torch.cuda.synchronize()
    sel = torch.arange(0, F, max(1, F // 16), device=dev)
    st = torch.from_numpy(stack).to(dev)
    ref = data[sel].double().to(torch.complex128) @ st.to(torch.complex128).T
    scale = (data[sel].double() @ st.abs().double().T).max().item()
    print(f'cfg4 tc   err/scale {((out_tc[sel] - ref).abs().max().item()) / scale:.3e}')
    print(f'cfg4 ffma err/scale {((out_ff[sel] - ref).abs().max().item()) / scale:.3e}', flush=True)
    for name, fn in (('ffma', lambda: gm.group_masks(data, plan, out=out_ff, kernel='ffma')),
                     ('tc chain=2', lambda: gm.group_masks(data, plan, out=out_tc, kernel='tc', chain=2)),
                     ('tc chain=4', lambda: gm.group_masks(data, plan, out=out_tc, kernel='tc', chain=4)),
                     ('tc chain=1', lambda: gm.group_masks(data, plan, out=out_tc, kernel='tc', chain=1))):
        ms = bench(fn)
        print(f'{name}: {ms:.3f} ms for {F} frames = {F / ms * 1e3 / 1e6:.3f} M frames/s, '
              f'{gb / ms * 1e3:.0f} GB/s = {gb / ms * 1e3 / 6549.4:.3f} of roofline', flush=True)


if __name__ == '__main__':
    main()
