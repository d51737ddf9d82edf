"""K8 (int8 tensor-core path, cfg3 geometry): timing with / without the fused frame sum and the
LTB200_K8_DEBUG switches (1 no mask MMAs, 2 no frame-sum MMAs).  Usage: k8_exp.py [ncu]"""
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
    m5 = torch.ones((5, K), dtype=torch.int8, device='cuda')
    m12 = torch.ones((12, K), dtype=torch.int8, device='cuda')
    if only == 'ncu':
        for _ in range(2):
            engine.masks_dense_i8(data, m5, sig_sum=sig)
        torch.cuda.synchronize()
        return

    def show(name, fn):
        b, md = bench(fn)
        print(f'{name}: best {b:.3f} ms ({gb / b * 1e3:.0f} GB/s, {gb / b * 1e3 / 6551:.3f}) '
              f'median {md:.3f}', flush=True)

    for dbg in (0, 1, 2, 3):
        os.environ['LTB200_K8_DEBUG'] = str(dbg)
        show(f'5 cols + sum  debug={dbg}', lambda: engine.masks_dense_i8(data, m5, sig_sum=sig))
        show(f'5 cols no sum debug={dbg}', lambda: engine.masks_dense_i8(data, m5))
    os.environ['LTB200_K8_DEBUG'] = '0'
    show('12 cols + sum', lambda: engine.masks_dense_i8(data, m12, sig_sum=sig))
    show('12 cols no sum', lambda: engine.masks_dense_i8(data, m12))


if __name__ == '__main__':
    main()
