"""A/B of two builds of the library on one box: burst (best of 10 after a pause) and sustained
(0.7 s back to back) K6 timings.   python scripts/k6_ab_lib_exp.py <path of libltb200.so | -> [M ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libertem_b200._lib as L  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] != '-':
    L.LIB_PATH = os.path.join(ROOT, sys.argv[1])
from libertem_b200 import engine  # noqa: E402
from k6_exp import bench  # noqa: E402

import pynvml  # noqa: E402


def main():
    cols = [int(a) for a in sys.argv[2:]] or [11, 19, 24]
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    F, K = 16384, 65536
    data = engine.synth_fill((F, K), np.float32, 1, 'cuda')
    gb = F * K * 4 / 1e9
    for M in cols:
        masks = engine.synth_fill((M, K), np.float32, 2, 'cuda')
        fn = lambda: engine.masks_dense_tc(data, masks)  # noqa: E731
        fn()
        torch.cuda.synchronize()
        time.sleep(1.5)
        burst = bench(fn, n=10)[0]
        sus = None
        for _ in range(2):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(800):
                fn()
            e1.record()
            mhz = []
            while not e1.query():
                mhz.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                time.sleep(0.02)
            torch.cuda.synchronize()
            sus = (e0.elapsed_time(e1) / 800, int(np.median(mhz[len(mhz) // 2:])))
        print(f'{os.path.relpath(L.LIB_PATH, ROOT)} M={M}: burst {burst:.3f} ms '
              f'({gb / burst * 1e3 / 6551:.3f}); sustained {sus[0]:.3f} ms '
              f'({gb / sus[0] * 1e3 / 6551:.3f}) at {sus[1]} MHz', flush=True)


if __name__ == '__main__':
    main()
