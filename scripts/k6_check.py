"""Bring-up / tuning script for the tcgen05 dense kernel (K6): structured patterns first (they
localise a wrong lane / column / k mapping), then random parity vs float64, then timing vs the
FFMA2 kernel.  Run on the GPU box: python scripts/k6_check.py [quick|time|all]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from libertem_b200 import engine  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run(data, masks, chain=0):
    out = engine.masks_dense_tc(dev(data), dev(masks), chain=chain)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def report(name, got, truth, scale=None):
    if scale is None:
        scale = np.abs(truth).max(axis=0, keepdims=True) + 1e-30
    err = np.abs(got - truth) / scale
    bad = np.argwhere(err > 1e-5)
    print(f'{name}: max rel err {err.max():.3e} mean {err.mean():.3e} bad {len(bad)}/{err.size}',
          flush=True)
    if len(bad):
        for f, m in bad[:6]:
            print(f'   out[{f},{m}] = {got[f, m]!r} expected {truth[f, m]!r}')
    return err.max()


def patterns():
    F, K, M = 256, 128, 8
    ones = np.ones((F, K), np.float32)
    mo = np.ones((M, K), np.float32)
    report('ones x ones', run(ones, mo), np.full((F, M), K, np.float64))
    rows = np.repeat(np.arange(F, dtype=np.float32)[:, None], K, 1)
    report('row index', run(rows, mo), rows.astype(np.float64) @ mo.T.astype(np.float64))
    mc = np.repeat(np.arange(1, M + 1, dtype=np.float32)[:, None], K, 1)
    report('mask index', run(ones, mc), ones.astype(np.float64) @ mc.T.astype(np.float64))
    kk = np.tile(np.arange(K, dtype=np.float32)[None, :], (F, 1)) + 1000 * np.arange(F)[:, None]
    kk = kk.astype(np.float32)
    for base in (0, 8, 24, 32, 100, 120):
        delta = np.zeros((M, K), np.float32)
        for m in range(M):
            delta[m, base + m] = 1.0
        report(f'delta k={base}..', run(kk, delta), kk.astype(np.float64) @ delta.T.astype(np.float64))
    rng = np.random.default_rng(0)
    for (F, K, M) in [(256, 128, 8), (256, 4096, 8), (300, 4096, 11), (1000, 2048, 19),
                      (513, 65536, 19), (2048, 16384, 32), (700, 8192, 40), (256, 132, 3)]:
        d = rng.random((F, K), dtype=np.float32)
        m = rng.random((M, K), dtype=np.float32) - 0.25
        truth = d.astype(np.float64) @ m.T.astype(np.float64)
        scale = (np.abs(d).astype(np.float64) @ np.abs(m).T.astype(np.float64)).max(axis=0, keepdims=True)
        for chain in (1, 4, 8, 64, 100000):
            report(f'random F={F} K={K} M={M} chain={chain}', run(d, m, chain), truth, scale)
    # positive data x positive masks: a truncating accumulator shows up as a bias here
    F, K, M = 512, 65536, 19
    d = rng.random((F, K), dtype=np.float32)
    m = rng.random((M, K), dtype=np.float32)
    truth = d.astype(np.float64) @ m.T.astype(np.float64)
    for chain in (1, 2, 4, 8, 16, 64, 100000):
        got = run(d, m, chain)
        rel = (got - truth) / truth
        print(f'positive chain={chain}: mean rel {rel.mean():+.3e} max |rel| {np.abs(rel).max():.3e}',
              flush=True)
    engine.set_k1_variant(2)
    got = engine.masks_dense(dev(d), dev(m)).cpu().numpy()
    engine.set_k1_variant(0)
    rel = (got - truth) / truth
    print(f'positive FFMA2 pair kernel: mean rel {rel.mean():+.3e} max |rel| {np.abs(rel).max():.3e}',
          flush=True)
    # accumulate + strided tile
    F, K, M = 300, 1024, 6
    big = rng.random((F, K + 64), dtype=np.float32)
    mk = rng.random((M, K), dtype=np.float32)
    tile = dev(big)[:, 32:32 + K]
    base = rng.random((F, M), dtype=np.float32)
    out = dev(base.copy())
    engine.masks_dense_tc(tile, dev(mk), out=out, accumulate=True)
    truth = base + big[:, 32:32 + K].astype(np.float64) @ mk.T.astype(np.float64)
    report('strided accumulate', out.cpu().numpy(), truth)


def timing():
    for (F, K, cols) in [(16384, 65536, (8, 11, 12, 16, 19, 24, 32))]:
        data = engine.synth_fill((F, K), np.float32, 1, 'cuda')
        gb = F * K * 4 / 1e9
        for M in cols:
            masks = engine.synth_fill((M, K), np.float32, 2, 'cuda')
            res = {}
            for name, fn in (('tc', lambda: engine.masks_dense_tc(data, masks)),
                             ('ffma2', lambda: engine.masks_dense(data, masks))):
                if name == 'ffma2':
                    engine.set_k1_variant(2)
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(8):
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                engine.set_k1_variant(0)
                res[name] = (min(ts), float(np.median(ts)))
            print(f'F={F} K={K} M={M}: ' + '  '.join(
                f'{n} best {b:.3f} ms ({gb / b * 1e3:.0f} GB/s) median {md:.3f} ms'
                for n, (b, md) in res.items()), flush=True)
        for chain in (2, 4, 8, 16, 64):
            masks = engine.synth_fill((19, K), np.float32, 2, 'cuda')
            for _ in range(2):
                engine.masks_dense_tc(data, masks, chain=chain)
            torch.cuda.synchronize()
            ts = []
            for _ in range(6):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                engine.masks_dense_tc(data, masks, chain=chain)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(f'M=19 chain={chain}: best {min(ts):.3f} ms ({gb / min(ts) * 1e3:.0f} GB/s)', flush=True)


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    t0 = time.time()
    if what in ('quick', 'all'):
        patterns()
    if what in ('time', 'all'):
        timing()
    print(f'done in {time.time() - t0:.1f} s')
