"""K7 schedules at small tile sizes (cfg4 geometry): python scripts/k7_small_exp.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine, group_masks as gm  # noqa: E402
from libertem_b200.analysis.radialfourier import radial_mask_factory  # noqa: E402
from k7_check import bench  # noqa: E402

dev = torch.device('cuda')
fac = radial_mask_factory(512, 512, 256, 256, 0, 364.0, 32, 24, use_sparse=False)
stack = np.asarray(fac()).reshape(800, -1)
for nb in (1, 4):
    plan = gm.build_plan(stack, 25, dev, n_bands=nb)
    for F in (256, 1024, 8192):
        data = engine.synth_fill((F, 512 * 512), np.float32, 104, dev)
        for kernel in ('tc', 'banded') if nb == 4 else ('banded',):
            for late in ('0', '1'):
                os.environ['LTB200_K7_LATE'] = late
                out = gm.group_masks(data, plan, kernel=kernel)
                ms = bench(lambda: gm.group_masks(data, plan, out=out, kernel=kernel))
                print(f'bands={nb} F={F} kernel={kernel} late={late}: {ms:.3f} ms, '
                      f'{F * 1048576 / ms / 1e6 / 6551:.3f} of roofline', flush=True)
        del data
