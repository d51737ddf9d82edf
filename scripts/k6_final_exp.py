"""K6: the shipped defaults against the round-1 form (one issuer, converters drain, four
products) at steady state (0.7 s per visit, 3 visits in rotation) and in bursts (best of 10)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine  # noqa: E402
from k6_exp import bench  # noqa: E402

import pynvml  # noqa: E402

OLD = {'LTB200_K6_ISSUERS': '1', 'LTB200_K6_DW': '0', 'LTB200_K6_THREE': '0'}


def setenv(old):
    for k, v in OLD.items():
        if old:
            os.environ[k] = v
        else:
            os.environ.pop(k, None)


def main():
    F, K = 16384, 65536
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    data = engine.synth_fill((F, K), np.float32, 1, 'cuda')
    gb = F * K * 4 / 1e9
    for M in (11, 16, 25, 28, 32):
        masks = engine.synth_fill((M, K), np.float32, 2, 'cuda')
        burst = {}
        for old in (True, False):
            setenv(old)
            time.sleep(1.0)                       # let the board cool between bursts
            burst[old] = bench(lambda: engine.masks_dense_tc(data, masks), n=10)[0]
        res = {True: [], False: []}
        for visit in range(3):
            for old in (True, False):
                setenv(old)
                engine.masks_dense_tc(data, masks)
                torch.cuda.synchronize()
                n = 800
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    engine.masks_dense_tc(data, masks)
                e1.record()
                mhz = []
                while not e1.query():
                    mhz.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    time.sleep(0.02)
                torch.cuda.synchronize()
                res[old].append((e0.elapsed_time(e1) / n, int(np.median(mhz[len(mhz) // 2:])),
                                 pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        for old in (True, False):
            ms = float(np.mean([x[0] for x in res[old][1:]]))
            print(f'M={M} {"round-1 form" if old else "defaults    "}: burst {burst[old]:.3f} ms '
                  f'({gb / burst[old] * 1e3 / 6551:.3f}); sustained {ms:.3f} ms '
                  f'({gb / ms * 1e3 / 6551:.3f}) at {res[old][-1][1]} MHz, {res[old][-1][2]:.0f} W',
                  flush=True)
    setenv(False)


if __name__ == '__main__':
    main()
