"""K6 at 11..32 columns: one / two MMA-issuing warps x converters / extra warps draining.
    python scripts/k6_issuers_exp.py [frames]      (LTB200_K6_ISSUERS, LTB200_K6_DW are per call)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine  # noqa: E402
from k6_exp import bench  # noqa: E402


class Clocks:
    """SM clock / power sampled every 5 ms while a timing loop runs (NVML)"""

    def __init__(self):
        import threading
        import pynvml
        pynvml.nvmlInit()
        self.nv = pynvml
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.stop = False
        self.mhz, self.watt = [], []
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        import time
        while not self.stop:
            self.mhz.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.watt.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            time.sleep(0.005)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join()

    def summary(self):
        return f'{int(np.median(self.mhz))} MHz {np.max(self.watt):.0f} W'


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    K = 65536
    data = engine.synth_fill((F, K), np.float32, 1, 'cuda')
    gb = F * K * 4 / 1e9
    for M in [int(a) for a in os.environ.get('K6_EXP_M', '11,19,24,28,32').split(',')]:
        masks = engine.synth_fill((M, K), np.float32, 2, 'cuda')
        for chain in (1, 4):
            line = [f'M={M} chain={chain}:']
            for iss in (1, 2):
                for dw in ((0, 1, 2) if M > 24 else (0,)):
                    os.environ['LTB200_K6_ISSUERS'] = str(iss)
                    os.environ['LTB200_K6_DW'] = str(dw)
                    for three in (0, 1):
                        os.environ['LTB200_K6_THREE'] = str(three)
                        with Clocks() as ck:
                            b, md = bench(lambda: engine.masks_dense_tc(data, masks, chain=chain),
                                          n=30)
                        line.append(f'\n   iss{iss}/dw{dw}/three{three} {b:.3f} ms '
                                    f'({gb / b * 1e3 / 6551:.3f}) med {md:.3f} [{ck.summary()}];')
            print(' '.join(line), flush=True)
    # accuracy of the three-product form (sum|x||m| scale, float64 truth on the device)
    d2 = data[:2048]
    for M in (11, 32):
        masks = engine.synth_fill((M, K), np.float32, 2, 'cuda')
        ref = d2.double() @ masks.double().T
        scale = (d2.double().abs() @ masks.double().abs().T).max()
        for three in (0, 1):
            os.environ['LTB200_K6_THREE'] = str(three)
            out = engine.masks_dense_tc(d2, masks)
            err = (out.double() - ref)
            print(f'M={M} three={three}: max err {float(err.abs().max() / scale):.2e} '
                  f'mean err {float(err.mean() / scale):+.2e}', flush=True)
    os.environ.pop('LTB200_K6_ISSUERS', None)
    os.environ.pop('LTB200_K6_DW', None)
    os.environ.pop('LTB200_K6_THREE', None)


if __name__ == '__main__':
    main()
