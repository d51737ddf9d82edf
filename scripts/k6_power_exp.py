"""Where does the power go?  K6 (11 columns) back to back for 0.7 s per visit with parts of the
kernel switched off (LTB200_K6_DEBUG: 1 no conversion, 2 no MMA, 4 no drain), and the read-only
streaming probe: ms per launch, SM clock and board power (NVML)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine  # noqa: E402

import pynvml  # noqa: E402


def sustained(fn, n):
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(2):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        mhz, w = [], []
        while not e1.query():
            mhz.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            w.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
            time.sleep(0.02)
        torch.cuda.synchronize()
        out = (e0.elapsed_time(e1) / n, int(np.median(mhz[len(mhz) // 2:])),
               float(np.median(w[len(w) // 2:])))
    return out


def main():
    pynvml.nvmlInit()
    F, K = 16384, 65536
    data = engine.synth_fill((F, K), np.float32, 1, 'cuda')
    gb = F * K * 4 / 1e9
    dst = torch.empty_like(data)
    ms, clk, w = sustained(lambda: dst.copy_(data), 400)
    print(f'torch copy (read + write): {ms:.3f} ms = {2 * gb / ms * 1e3:.0f} GB/s at {clk} MHz, {w:.0f} W',
          flush=True)
    del dst
    for M in (11, 32):
        masks = engine.synth_fill((M, K), np.float32, 2, 'cuda')
        for dbg, what in ((0, 'full kernel'), (1, 'no conversion'), (2, 'no MMA'),
                          (3, 'no conversion, no MMA'), (7, 'stream only')):
            os.environ['LTB200_K6_DEBUG'] = str(dbg)
            ms, clk, w = sustained(lambda: engine.masks_dense_tc(data, masks), 800)
            print(f'K6 M={M} {what}: {ms:.3f} ms ({gb / ms * 1e3 / 6551:.3f}) at {clk} MHz, '
                  f'{w:.0f} W', flush=True)
        os.environ.pop('LTB200_K6_DEBUG', None)


if __name__ == '__main__':
    main()
