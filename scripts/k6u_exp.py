"""K6 uint16 form (cfg3 geometry): where the time goes.  LTB200_K6_DEBUG switches (1 no convert,
2 no MMA, 4 no drain), with / without the fused frame sum, against the FFMA2 kernel.
Usage: python scripts/k6u_exp.py [ncu]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine  # noqa: E402
from k6_exp import bench  # noqa: E402


def main():
    F, K = 262144, 16384
    only = sys.argv[1] if len(sys.argv) > 1 else None
    data = engine.synth_fill((F, K), np.uint16, 1, 'cuda')
    gb = F * K * 2 / 1e9
    sig = torch.zeros(K, dtype=torch.float32, device='cuda')
    m5 = engine.synth_fill((5, K), np.float32, 2, 'cuda')
    m9 = engine.synth_fill((9, K), np.float32, 3, 'cuda')
    if only == 'ncu':
        for _ in range(2):
            engine.masks_dense_tc_u16(data, m5, sig_sum=sig)
        torch.cuda.synchronize()
        return

    def show(name, fn):
        b, md = bench(fn)
        print(f'{name}: best {b:.3f} ms ({gb / b * 1e3:.0f} GB/s, {gb / b * 1e3 / 6551:.3f}) '
              f'median {md:.3f}', flush=True)

    for dbg in (0, 1, 2, 3, 4):
        os.environ['LTB200_K6_DEBUG'] = str(dbg)
        show(f'5 cols + sum  debug={dbg}', lambda: engine.masks_dense_tc_u16(data, m5, sig_sum=sig))
        show(f'5 cols no sum debug={dbg}', lambda: engine.masks_dense_tc_u16(data, m5))
    os.environ['LTB200_K6_DEBUG'] = '0'
    for chain in (2, 4):
        show(f'5 cols + sum chain={chain}',
             lambda: engine.masks_dense_tc_u16(data, m5, sig_sum=sig, chain=chain))
    show('9 cols + sum', lambda: engine.masks_dense_tc_u16(data, m9, sig_sum=sig))
    show('9 cols no sum', lambda: engine.masks_dense_tc_u16(data, m9))
    engine.set_k1_variant(2)
    show('FFMA2 5 cols + sum', lambda: engine.masks_dense(data, m5, sig_sum=sig))
    show('FFMA2 5 cols no sum', lambda: engine.masks_dense(data, m5))
    engine.set_k1_variant(0)


if __name__ == '__main__':
    main()
