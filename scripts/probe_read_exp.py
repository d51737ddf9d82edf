"""read-only streaming probes vs the copy peak (scripts: measurement only)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libertem_b200 import engine

dev = torch.device('cuda')
buf = torch.empty(16 << 30, dtype=torch.uint8, device=dev)
buf.zero_()
dst = torch.empty(8 << 30, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for mode in (0, 1):
    ms = timed(lambda: engine.probe_read(buf, mode))
    print(f'probe_read mode {mode}: {ms:.3f} ms  {buf.numel() / ms / 1e6:.1f} GB/s', flush=True)
ms = timed(lambda: dst.copy_(buf[:8 << 30]))
print(f'copy 8 GiB: {ms:.3f} ms  {2 * dst.numel() / ms / 1e6:.1f} GB/s (read+write)')
ms = timed(lambda: buf.sum(dtype=torch.int64) if False else torch.count_nonzero(buf.view(torch.int64)))
print(f'torch count_nonzero (read only): {ms:.3f} ms  {buf.numel() / ms / 1e6:.1f} GB/s')
