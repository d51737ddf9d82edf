"""Quick K1 timing sweep (device-resident data, CUDA events).
Usage: python scripts/k1_timing.py [F] [M list] [variants] [dtype]"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from libertem_b200 import engine

F = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
Ms = [int(a) for a in (sys.argv[2] if len(sys.argv) > 2 else '1,4,8,11,12,16,19,24').split(',')]
variants = (sys.argv[3] if len(sys.argv) > 3 else 'eo,pair').split(',')
dtype = sys.argv[4] if len(sys.argv) > 4 else 'float32'
K = int(sys.argv[5]) if len(sys.argv) > 5 else 65536
with_sig = len(sys.argv) > 6 and sys.argv[6] == 'sig'
data = engine.synth_fill((F, K), np.dtype(dtype), 1, 'cuda')
res = []
for var in variants:
    engine.set_k1_variant({'auto': 0, 'eo': 1, 'pair': 2}[var])
    for M in Ms:
        masks = torch.rand((M, K), device='cuda')
        out = torch.zeros((F, M), device='cuda')
        ss = torch.zeros(K, device='cuda') if with_sig else None
        for _ in range(3):
            engine.masks_dense(data, masks, out=out, sig_sum=ss)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.masks_dense(data, masks, out=out, sig_sum=ss)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = min(ts)
        gbs = F * K * data.element_size() / t / 1e6
        res.append(dict(variant=var, kernel=engine.last_kernel(), M=M, F=F, K=K, dtype=dtype,
                        sig=with_sig, ms=round(t, 4), ms_med=round(float(np.median(ts)), 4),
                        GBps=round(gbs, 1), frac=round(gbs / 6549.4, 4),
                        tflops=round(2 * F * K * M / t / 1e9, 2)))
        print(json.dumps(res[-1]), flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/k1_timing_%s.json' % dtype, 'w'), indent=1)
