"""Re-creates the inputs of tests/golden/make_golden.py from seeds (no reference needed)."""
import numpy as np

from oracle import synth, masks_gen


def mixed_masks(sy, sx, count, seed):
    """Same mask mix as make_golden.mixed_masks, built from the oracle's generators."""
    out = []
    cy, cx = sy // 2, sx // 2
    for i in range(count):
        kind = i % 4
        if kind == 0:
            m = synth.uniform_f32(0, sy * sx, seed + i).reshape(sy, sx)
        elif kind == 1:
            m = masks_gen.circular(cx, cy, sx, sy, radius=min(sy, sx) / 4 + i).astype(np.float32)
        elif kind == 2:
            m = masks_gen.ring(cx, cy, sx, sy, radius=min(sy, sx) / 3 + i,
                               radius_inner=min(sy, sx) / 6).astype(np.float32)
        else:
            m = (masks_gen.gradient_x(sx, sy) - cx) * 0.5 + (masks_gen.gradient_y(sx, sy) - cy) * 0.25
            m = m.astype(np.float32)
        out.append(m)
    return np.stack(out)


def roi_from_seed(nav_shape, seed):
    n = int(np.prod(nav_shape))
    return (synth.hash_u32(0, n, seed) % 3 != 0).reshape(nav_shape)


def ring_stack(sig_shape, rings, cy, cx):
    sy, sx = sig_shape
    return np.stack([masks_gen.ring(cx, cy, sx, sy, ro, ri) for ri, ro in rings])


def int_detector_inputs(name, meta):
    """data and binary mask stack of the integer-detector goldens (make_golden.g_int_detector)"""
    shape = tuple(meta['shape'])
    if name == 'u16':
        data = synth.dataset(shape, np.uint16, meta['data_seed'])
        stack = np.stack([masks_gen.circular(128, 128, 256, 256, 40),
                          masks_gen.ring(128, 128, 256, 256, 90, 60)])
    else:
        data = (synth.hash_u32(0, int(np.prod(shape)), meta['data_seed']) % 23).astype(
            np.uint8).reshape(shape)
        stack = np.stack([masks_gen.circular(32, 32, 64, 64, 10),
                          masks_gen.ring(32, 32, 64, 64, 30, 20)])
    return data, stack.astype(np.float32)
