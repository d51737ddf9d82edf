"""find the boxes whose contribution to a ring is wrong (narrow-ring geometry, gate lifted)"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libertem_b200 import group_masks as gm, masks as M, walk_plan as wp  # noqa: E402
from libertem_b200.analysis.radialfourier import radial_mask_factory  # noqa: E402

S, nb, mo, F = 128, 16, 6, 1000
dev = torch.device('cuda')
ro = M.bounding_radius(S / 2, S / 2, S, S)
st = np.asarray(radial_mask_factory(S, S, S / 2, S / 2, 0, ro, nb, mo, use_sparse=False)())
flat = st.reshape(st.shape[0], -1).astype(np.complex64)
plan = gm.build_plan(flat, mo + 1, dev, walk_max_dup=1e9)
w = wp.build_walk(flat, mo + 1, max_dup=np.inf, max_per_slice=1 << 30)
ring = 14
sup = np.nonzero(np.abs(flat[ring * (mo + 1)]) > 0)[0]
boxes = sorted(set((sup // 32).tolist()))
print('ring', ring, 'pixels', len(sup), 'boxes', len(boxes), 'segments', w['n_segments'], flush=True)
rng = np.random.default_rng(0)
base = rng.random(S * S).astype(np.float32)
bad_boxes = {}
for rep in range(3):
    for b in boxes:
        x = np.zeros(S * S, dtype=np.float32)
        x[b * 32:(b + 1) * 32] = base[b * 32:(b + 1) * 32]
        t = torch.from_numpy(np.broadcast_to(x, (F, S * S)).copy()).cuda()
        out = gm.group_masks(t, plan, kernel='walk')
        got = out[:, ring * (mo + 1)].real.double().cpu().numpy()
        ref = float(x.astype(np.float64) @ flat[ring * (mo + 1)].real.astype(np.float64))
        dev_rows = np.nonzero(np.abs(got - ref) > 1e-4 * max(1.0, abs(ref)))[0]
        if len(dev_rows):
            bad_boxes.setdefault(b, []).append((len(dev_rows), float(got[dev_rows[0]]), ref))
for b, v in sorted(bad_boxes.items()):
    y, xb = divmod(b, S // 32)
    visits = [i for i, bw in enumerate(w['boxes']) if (int(bw) & ~31) == b * 32]
    print(f'box {b} (row {y}, x {xb * 32}): {v[:3]}  visits at {visits} masks {[int(w["boxes"][i]) & 15 for i in visits]}')
print('visit_off', w['visit_off'])
print('bad boxes:', len(bad_boxes), 'of', len(boxes))
