"""K7 on the cfg4 geometry, ring-major vs banded schedule.
    python scripts/k7_band_exp.py <n_bands> <kernel tc|banded|sym|walk> [frames] [ncu]
(sym: the experimental mirror-symmetric plan, needs LTB200_K7_SYM=1)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine, group_masks as gm  # noqa: E402
from libertem_b200.analysis.radialfourier import radial_mask_factory  # noqa: E402
from k7_check import bench  # noqa: E402


def main():
    n_bands = int(sys.argv[1])
    kernel = sys.argv[2]
    F = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
    ncu = len(sys.argv) > 4
    dev = torch.device('cuda')
    fac = radial_mask_factory(512, 512, 256, 256, 0, 364.0, 32, 24, use_sparse=False)
    stack = np.asarray(fac()).reshape(800, -1)
    plan = gm.build_plan(stack, 25, dev, n_bands=n_bands, sig_shape=(512, 512))
    n_ent = (plan.n_entries if kernel == 'tc' else plan.walk['n_entries'] if kernel == 'walk' else
             int((plan.sym['main'] if kernel == 'sym' else plan.banded)['group_off_host'][-1]))
    data = engine.synth_fill((F, 512 * 512), np.float32, 104, dev)
    out = gm.group_masks(data, plan, kernel=kernel)
    if ncu:
        gm.group_masks(data, plan, out=out, kernel=kernel)
        torch.cuda.synchronize()
        return
    gb = F * 512 * 512 * 4 / 1e9
    ms = bench(lambda: gm.group_masks(data, plan, out=out, kernel=kernel))
    print(f'bands={n_bands} kernel={kernel} FBG={os.environ.get("LTB200_K7_FBG", "-")} entries={n_ent}: '
          f'{ms:.3f} ms, {gb / ms * 1e3 / 6551:.3f} of roofline', flush=True)


if __name__ == '__main__':
    main()
