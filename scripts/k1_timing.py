"""Quick K1 timing sweep (device-resident data, CUDA events).  Usage: python scripts/k1_timing.py [F]"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from libertem_b200 import engine

F = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = 65536
data = engine.synth_fill((F, K), np.float32, 1, 'cuda')
res = []
for M in [int(a) for a in (sys.argv[2].split(',') if len(sys.argv) > 2 else '1,4,8,11,12,16,19,24'.split(','))]:
    masks = torch.rand((M, K), device='cuda')
    out = torch.zeros((F, M), device='cuda')
    for _ in range(3):
        engine.masks_dense(data, masks, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.masks_dense(data, masks, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    gbs = F * K * 4 / t / 1e6
    res.append(dict(M=M, F=F, ms=t, ms_med=float(np.median(ts)), GBps=gbs, frac=gbs / 6549.4,
                    tflops=2 * F * K * M / t / 1e9))
    print(json.dumps(res[-1]), flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/k1_timing.json', 'w'), indent=1)
