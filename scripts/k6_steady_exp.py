"""K6 variants at thermal / power steady state: every configuration runs back to back for
~0.7 s per visit, three visits in rotation; reports ms per launch of the later visits and the SM
clock NVML saw.   python scripts/k6_steady_exp.py [columns ...]
(LTB200_K6_ISSUERS, LTB200_K6_DW, LTB200_K6_THREE are read per call)"""
import itertools
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import engine  # noqa: E402

import pynvml  # noqa: E402


def main():
    cols = [int(a) for a in sys.argv[1:]] or [32]
    F, K = 16384, 65536
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    data = engine.synth_fill((F, K), np.float32, 1, 'cuda')
    gb = F * K * 4 / 1e9
    for M in cols:
        masks = engine.synth_fill((M, K), np.float32, 2, 'cuda')
        nh = 8 if M <= 8 else 16 if M <= 16 else 24 if M <= 24 else 32
        dws = (0, 1, 2) if nh == 32 else (0,)
        threes = (0, 1) if nh % 16 == 0 else (0,)
        chains = [int(c) for c in os.environ.get('K6_EXP_CHAINS', '1').split(',')]
        cfgs = list(itertools.product(chains, (1, 2), dws, threes))
        res = {c: [] for c in cfgs}
        for visit in range(3):
            for cfg in cfgs:
                chain, iss, dw, three = cfg
                os.environ['LTB200_K6_ISSUERS'] = str(iss)
                os.environ['LTB200_K6_DW'] = str(dw)
                os.environ['LTB200_K6_THREE'] = str(three)
                fn = lambda: engine.masks_dense_tc(data, masks, chain=chain)  # noqa: E731
                fn()
                torch.cuda.synchronize()
                n = 800
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    fn()
                e1.record()
                mhz = []
                while not e1.query():
                    mhz.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    time.sleep(0.02)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                w = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
                res[cfg].append((ms, int(np.median(mhz[len(mhz) // 2:])) if mhz else 0, w))
        for cfg in cfgs:
            chain, iss, dw, three = cfg
            r = res[cfg]
            ms = float(np.mean([x[0] for x in r[1:]]))
            print(f'M={M} chain={chain} issuers={iss} dw={dw} three={three}: '
                  + ' '.join(f'{x[0]:.3f}' for x in r)
                  + f' ms -> {ms:.3f} ms = {gb / ms * 1e3 / 6551:.3f} of roofline '
                  f'[{r[-1][1]} MHz, {r[-1][2]:.0f} W]', flush=True)
    for k in ('LTB200_K6_ISSUERS', 'LTB200_K6_DW', 'LTB200_K6_THREE'):
        os.environ.pop(k, None)


if __name__ == '__main__':
    main()
