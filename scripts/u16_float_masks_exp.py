"""cfg3 geometry (512x512 nav x 128x128 uint16) with NON-integer masks: the fixed-point int8
path (K8) vs the float kernels (LTB200_FLOAT_MASKS_INT8=0).  python scripts/u16_float_masks_exp.py [n_masks]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libertem_b200 import engine
from libertem_b200.io import SyntheticDataSet
from libertem_b200.runner import UDFRunner
from libertem_b200.udf import ApplyMasksUDF, SumUDF, SumSigUDF
import libertem_b200.runner as R

dev = torch.device('cuda')
n_masks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ds = SyntheticDataSet((512, 512, 128, 128), np.uint16, seed=103, num_partitions=1)
ds.materialize(dev)
yy, xx = np.mgrid[:128, :128]
r = np.hypot(yy - 63.5, xx - 62.2)
masks = np.stack([np.exp(-((r - 6 * (i + 1)) / 5.0) ** 2) * (0.3 + i) for i in range(n_masks)]
                 ).astype(np.float32)
for flag in (True, False):
    R.FLOAT_MASKS_INT8 = flag
    for with_sum in (True, False):
        udfs = ([SumUDF(), SumSigUDF()] if with_sum else []) + [
            ApplyMasksUDF(mask_factories=lambda: masks)]
        runner = UDFRunner(udfs)
        for _ in range(3):
            runner.run_for_dataset(ds, device=dev, finalize=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            runner.run_for_dataset(ds, device=dev, finalize=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = ds.shape.size * 2
        print(json.dumps(dict(fixed_point_int8=flag, n_float_masks=n_masks, sum_udfs=with_sum,
                              ms=round(ms, 4), GBps=round(nbytes / ms / 1e6, 1),
                              roofline_frac=round(nbytes / ms / 1e6 / 6551, 4),
                              kernel=engine.last_kernel(),
                              int8_passes=runner.stats.get('int8_passes', 0))), flush=True)
