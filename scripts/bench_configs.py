"""Device-timed throughput of the BASELINE.json configs other than the headline (bench.py):
cfg1, cfg3, cfg4 (nav sub-sample), cfg5 (one GPU's shard).  Writes gpurun_out/configs.json.
Usage: python scripts/bench_configs.py [cfg1,cfg3,cfg4,cfg5]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libertem_b200 import engine, masks as M
from libertem_b200.io import SyntheticDataSet
from libertem_b200.runner import UDFRunner
from libertem_b200.udf import ApplyMasksUDF, CoMUDF, SumUDF, SumSigUDF
from libertem_b200.api import Context
from bench import bench_masks

PEAK = 6551.0
dev = torch.device('cuda')


def uni(n, seed):
    return engine.synth_fill((n,), np.float32, seed, dev).cpu().numpy()


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    engine.launch_count(reset=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, engine.launch_count() / steps


def report(name, ds, udfs, note='', steps=10):
    runner = UDFRunner(udfs)
    ds.materialize(dev)
    ms, launches = timed(lambda: runner.run_for_dataset(ds, device=dev, finalize=False), steps)
    frames = ds.shape.nav.size
    nbytes = ds.shape.size * ds.dtype.itemsize
    gbs = nbytes / ms / 1e6
    line = dict(config=name, frames=frames, ms_per_pass=round(ms, 4),
                frames_per_s=round(frames / ms * 1e3), GBps=round(gbs, 1),
                roofline_frac=round(gbs / PEAK, 4), launches_per_pass=launches, note=note,
                unfused_calls=runner.stats['unfused_calls'],
                int8_passes=runner.stats.get('int8_passes', 0), last_kernel=engine.last_kernel())
    print(json.dumps(line), flush=True)
    return line


which = (sys.argv[1] if len(sys.argv) > 1 else 'cfg1,cfg3,cfg4,cfg5').split(',')
out = []
if 'cfg1' in which:
    ds = SyntheticDataSet((32, 32, 64, 64), np.float32, seed=101, num_partitions=1)
    mask = uni(4096, 201).reshape(64, 64)
    out.append(report('cfg1: 32x32 nav x 64x64 sig f32, 1 dense mask', ds,
                      [ApplyMasksUDF(mask_factories=[lambda: mask])],
                      'latency-bound (16 MiB)', steps=50))
if 'cfg3' in which:
    ds = SyntheticDataSet((512, 512, 128, 128), np.uint16, seed=103, num_partitions=1)
    rings = [(8, 16), (20, 28), (32, 40), (44, 52)]
    facs = [lambda ri=ri, ro=ro: M.ring(64, 64, 128, 128, ro, ri) for ri, ro in rings]
    out.append(report('cfg3: 512x512 nav x 128x128 sig u16, SumUDF+SumSigUDF+4 sparse ring masks',
                      ds, [SumUDF(), SumSigUDF(),
                           ApplyMasksUDF(mask_factories=facs, use_sparse=True,
                                         mask_dtype=np.float32)],
                      'one fused pass on the int8 tensor cores (K8): u16 TMA ingest, 5 columns + frame sum as MMAs'))
    del ds
    torch.cuda.empty_cache()
if 'cfg5' in which:
    # one rank's shard of 1024x1024 nav on 8 GPUs = 131072 frames (32 GiB), 16 masks + CoM
    ds = SyntheticDataSet((128, 1024, 256, 256), np.float32, seed=105, num_partitions=1)
    stack = bench_masks(256, 256, 16, 2005, uni)
    out.append(report('cfg5 shard: 128x1024 nav x 256x256 sig f32, 16 dense masks + CoM (1 of 8 GPUs)',
                      ds, [ApplyMasksUDF(mask_factories=lambda: stack, mask_count=16,
                                         mask_dtype=np.float32, use_sparse=False), CoMUDF()],
                      '19 fused columns', steps=5))
    del ds
    torch.cuda.empty_cache()
if 'cfg4' in which:
    # nav sub-sample of 256x256 nav x 512x512 sig (1 MiB / frame)
    ds = SyntheticDataSet((int(os.environ.get('CFG4_NAV0', '128')), 64, 512, 512), np.float32, seed=104, num_partitions=1)
    ctx = Context()
    a = ctx.create_radial_fourier_analysis(ds, n_bins=32)
    out.append(report('cfg4 (nav %dx64 sub-sample): 512x512 sig f32, radial Fourier 32 bins x 25 orders' % ds.shape[0],
                      ds, [a.get_udf()], 'K7 group-sparse tensor-core kernel (quad gather, banded schedule), %d complex masks' % a.parameters['mask_count'], steps=3))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/configs.json', 'w'), indent=1)
