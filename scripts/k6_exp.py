"""K6 bottleneck experiments: LTB200_K6_DEBUG switches (1 no convert, 2 no MMA, 4 no drain)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine  # noqa: E402


def bench(fn, n=8):
    for _ in range(3):
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
    return min(ts), float(np.median(ts))


def main():
    F, K = 16384, 65536
    only = sys.argv[1] if len(sys.argv) > 1 else None
    data = engine.synth_fill((F, K), np.float32, 1, 'cuda')
    gb = F * K * 4 / 1e9
    for M in [int(a) for a in os.environ.get('K6_EXP_M', '11,19,32').split(',')]:
        masks = engine.synth_fill((M, K), np.float32, 2, 'cuda')
        if only == 'ncu':
            os.environ['LTB200_K6_DEBUG'] = '0'
            for _ in range(3):
                engine.masks_dense_tc(data, masks)
            torch.cuda.synchronize()
            continue
        for dbg, chain in ((0, 1), (0, 2), (0, 4), (0, 8), (2, 1), (1, 1), (3, 1)):
            os.environ['LTB200_K6_DEBUG'] = str(dbg)
            b, md = bench(lambda: engine.masks_dense_tc(data, masks, chain=chain))
            print(f'M={M} debug={dbg} chain={chain or "default"}: best {b:.3f} ms '
                  f'({gb / b * 1e3:.0f} GB/s) median {md:.3f}', flush=True)
        os.environ['LTB200_K6_DEBUG'] = '0'


if __name__ == '__main__':
    main()
