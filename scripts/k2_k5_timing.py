"""Device-timed throughput of the sparse-CSC kernel (K2) and the shifted-mask kernel (K5) on
detector-sized inputs; prints one JSON line per case (profiles/r2_k2_k5.json).
    python scripts/k2_k5_timing.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from libertem_b200 import engine, masks as M

PEAK = 6551.0
dev = torch.device('cuda')


def timed(fn, steps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def line(name, nbytes, ms, **kw):
    d = dict(case=name, ms=round(ms, 4), GBps=round(nbytes / ms / 1e6, 1),
             roofline_frac=round(nbytes / ms / 1e6 / PEAK, 4), **kw)
    print(json.dumps(d), flush=True)


# ---- K2: 65536 frames x 128x128 uint16, 4 sparse ring masks (cfg3 geometry) ------------------
F, sy, sx = 65536, 128, 128
data = engine.synth_fill((F, sy * sx), np.uint16, 3, dev)
rings = [(8, 16), (20, 28), (32, 40), (44, 52)]
stack = np.stack([M.ring(64, 64, sx, sy, ro, ri) for ri, ro in rings]).astype(np.float32)
flat = stack.reshape(4, -1)
indptr, indices, values = [0], [], []
for m in flat:
    nz = np.nonzero(m)[0]
    indices.append(nz.astype(np.int32))
    values.append(m[nz])
    indptr.append(indptr[-1] + len(nz))
ip = torch.tensor(indptr, dtype=torch.int32, device=dev)
ix = torch.from_numpy(np.concatenate(indices)).to(dev)
vv = torch.from_numpy(np.concatenate(values)).to(dev)
out = torch.zeros((F, 4), dtype=torch.float32, device=dev)
ms = timed(lambda: engine.masks_csc(data, ip, ix, vv, 4, out=out))
nnz = int(indptr[-1])
line('K2 csc: 65536 x 128x128 u16, 4 ring masks (%d nnz = %.1f %% fill)' % (nnz, 100 * nnz / 4 / sy / sx),
     F * sy * sx * 2, ms, gathered_GBps=round(F * nnz * 2 / ms / 1e6, 1),
     note='roofline vs the WHOLE frame bytes; the kernel only touches the ring pixels')
rows = torch.from_numpy(flat).to(dev)
ms_dense = timed(lambda: engine.masks_dense(data, rows, out=out))
line('same masks through the dense pass (what the runner does for <= 24 sparse masks)',
     F * sy * sx * 2, ms_dense, kernel=engine.last_kernel())
del data, out

# ---- K5: 16384 frames x 256x256 float32, 8 masks, per-frame shifts ----------------------------
F, sy, sx, NM = 16384, 256, 256, 8
tile = engine.synth_fill((F, sy, sx), np.float32, 5, dev)
masks = engine.synth_fill((NM, sy * sx), np.float32, 6, dev)
sh = (torch.randint(-12, 13, (F, 2), device=dev, dtype=torch.int32))   # device shifts: warp-per-frame kernel
out = torch.zeros((F, NM), dtype=torch.float32, device=dev)
ms = timed(lambda: engine.masks_shifted(tile, masks, sh, out=out))
line('K5 shifted: 16384 x 256x256 f32, 8 masks, per-frame (dy, dx) in [-12, 12]', F * sy * sx * 4, ms)
for nm, dmax in ((8, 12), (8, 4), (3, 12), (3, 4), (4, 2)):
    shd = torch.randint(-dmax, dmax + 1, (F, 2), dtype=torch.int32)        # host shifts: banded
    o = torch.zeros((F, nm), dtype=torch.float32, device=dev)
    ms = timed(lambda: engine.masks_shifted(tile, masks[:nm], shd, out=o))
    line('K5 banded: same frames, %d masks, |dy|,|dx| <= %d' % (nm, dmax), F * sy * sx * 4, ms,
         kernel=engine.last_kernel())
out3 = torch.zeros((F, 3), dtype=torch.float32, device=dev)
ms = timed(lambda: engine.masks_shifted(tile, masks[:3], sh, out=out3))
line('K5 shifted: same, 3 masks', F * sy * sx * 4, ms)
m64 = masks.double()
out64 = torch.zeros((F, NM), dtype=torch.float64, device=dev)
ms = timed(lambda: engine.masks_shifted(tile, m64, sh, out=out64), steps=5)
line('K5 shifted: same, float64 masks / accumulation / result', F * sy * sx * 4, ms)
