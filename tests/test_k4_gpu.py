"""GPU parity of the group-sparse kernels -- K4 (ltb200_group_masks, FFMA2) and K7
(ltb200_group_masks_tc, tcgen05 split-TF32) -- against numpy float64."""
import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def make_stack(n_groups, size, K, seed, empty_group=None):
    rng = np.random.default_rng(seed)
    stack = np.zeros((n_groups * size, K), dtype=np.complex64)
    for g in range(n_groups):
        if g == empty_group:
            continue
        n = int(rng.integers(1, max(2, K // 3)))
        px = np.sort(rng.choice(K, size=n, replace=False))
        vals = (rng.random((size, n)) - 0.5 + 1j * (rng.random((size, n)) - 0.5))
        vals[np.abs(vals) == 0] = 0.25
        stack[g * size:(g + 1) * size, px] = vals.astype(np.complex64)
    return stack


@pytest.mark.parametrize('kernel', ['ffma', 'tc', 'banded', 'auto'])
@pytest.mark.parametrize('F,K,n_groups,size', [(64, 512, 3, 25), (100, 1000, 5, 7), (7, 300, 2, 28),
                                               (200, 4096, 32, 25), (65, 640, 4, 1),
                                               (129, 1000, 5, 4), (300, 2048, 6, 16),
                                               (1000, 4096, 8, 25), (128, 700, 3, 9)])
def test_group_masks_matches_numpy(F, K, n_groups, size, kernel):
    from libertem_b200 import group_masks as gm
    stack = make_stack(n_groups, size, K, seed=F + K, empty_group=1 if n_groups > 3 else None)
    assert gm.find_groups(stack) in (size, None) or size == 1
    plan = gm.build_plan(stack, size, torch.device('cuda'))
    data = synth.uniform_f32(0, F * K, 9).reshape(F, K)
    t = torch.from_numpy(data).cuda()
    out = gm.group_masks(t, plan, kernel=kernel).cpu().numpy()
    ref = data.astype(np.float64) @ stack.astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(stack).astype(np.float64).T).max() + 1e-30
    assert out.shape == ref.shape and out.dtype == np.complex64
    assert np.abs(out - ref).max() / scale <= 2e-6
    # accumulate
    out2 = gm.group_masks(t, plan, out=torch.from_numpy(out).cuda(), accumulate=True,
                          kernel=kernel).cpu().numpy()
    assert np.abs(out2 - 2 * ref).max() / scale <= 4e-6
    # strided tile
    big = torch.zeros((F, K + 24), device='cuda')
    big[:, 8:8 + K] = t
    out3 = gm.group_masks(big[:, 8:8 + K], plan, kernel=kernel).cpu().numpy()
    assert np.array_equal(out3, out)


@pytest.mark.parametrize('chain', [1, 2, 3, 4, 7])
def test_group_masks_tc_chains(chain):
    """every TMEM chain length of K7 drains correctly; the two kernels agree"""
    from libertem_b200 import group_masks as gm
    F, K, n_groups, size = 400, 3000, 4, 25
    stack = make_stack(n_groups, size, K, seed=77)
    plan = gm.build_plan(stack, size, torch.device('cuda'))
    data = synth.uniform_f32(0, F * K, 10).reshape(F, K)
    t = torch.from_numpy(data).cuda()
    ref = data.astype(np.float64) @ stack.astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(stack).astype(np.float64).T).max() + 1e-30
    out = gm.group_masks(t, plan, kernel='tc', chain=chain).cpu().numpy()
    assert np.abs(out - ref).max() / scale <= 3e-6
    ffma = gm.group_masks(t, plan, kernel='ffma').cpu().numpy()
    assert np.abs(out - ffma).max() / scale <= 3e-6


@pytest.mark.parametrize('F,K,n_groups,size,n_bands', [(300, 4096, 6, 25, 4), (1000, 3000, 5, 7, 3),
                                                       (129, 2048, 4, 25, 16), (2000, 8192, 8, 25, 2),
                                                       (64, 1000, 3, 9, 5)])
def test_group_masks_banded(F, K, n_groups, size, n_bands):
    """banded plan of K7 ((pixel band, ring) groups, band-major schedule, fixed-order band
    reduction): same result as the ring-major schedule and as numpy float64; empty (band, ring)
    groups, an empty ring, accumulate and strided tiles"""
    from libertem_b200 import group_masks as gm
    stack = make_stack(n_groups, size, K, seed=F + K + 1, empty_group=1 if n_groups > 3 else None)
    stack[:size, K // 2:] = 0                     # ring 0 lives in the first bands only
    plan = gm.build_plan(stack, size, torch.device('cuda'), n_bands=n_bands)
    assert plan.banded is not None and plan.banded['n_groups'] == n_bands * n_groups
    data = synth.uniform_f32(0, F * K, 9).reshape(F, K)
    t = torch.from_numpy(data).cuda()
    out = gm.group_masks(t, plan, kernel='banded').cpu().numpy()
    ref = data.astype(np.float64) @ stack.astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(stack).astype(np.float64).T).max() + 1e-30
    assert out.shape == ref.shape and out.dtype == np.complex64
    assert np.abs(out - ref).max() / scale <= 2e-6
    tc = gm.group_masks(t, plan, kernel='tc').cpu().numpy()
    assert np.abs(out - tc).max() / scale <= 2e-6
    out2 = gm.group_masks(t, plan, out=torch.from_numpy(out).cuda(), accumulate=True,
                          kernel='banded').cpu().numpy()
    assert np.abs(out2 - 2 * ref).max() / scale <= 4e-6
    big = torch.zeros((F, K + 24), device='cuda')
    big[:, 8:8 + K] = t
    out3 = gm.group_masks(big[:, 8:8 + K], plan, kernel='banded').cpu().numpy()
    assert np.array_equal(out3, out)              # deterministic
    import os
    os.environ['LTB200_K7_FBG'] = '2'
    try:
        out4 = gm.group_masks(t, plan, kernel='banded').cpu().numpy()
    finally:
        del os.environ['LTB200_K7_FBG']
    assert np.array_equal(out4, out)              # schedule does not change the arithmetic


@pytest.mark.parametrize('S,n_bins,max_order,F', [(64, 4, 6, 300), (128, 8, 24, 1000)])
def test_group_masks_sym(S, n_bins, max_order, F):
    """mirror-symmetric plan of K7 on the reference's radial masks: same result as the banded
    quad plan and as numpy float64"""
    from libertem_b200 import group_masks as gm, masks as M
    from libertem_b200.analysis.radialfourier import radial_mask_factory
    ro = M.bounding_radius(S / 2, S / 2, S, S)
    stack = np.asarray(radial_mask_factory(S, S, S / 2, S / 2, 0, ro, n_bins, max_order,
                                           use_sparse=False)()).astype(np.complex64)
    plan = gm.build_plan(stack, max_order + 1, torch.device('cuda'), n_bands=2, sym=True)
    assert plan.sym is not None
    data = synth.uniform_f32(0, F * S * S, 12).reshape(F, S * S)
    t = torch.from_numpy(data).cuda()
    out = gm.group_masks(t, plan, kernel='sym').cpu().numpy()
    flat = stack.reshape(stack.shape[0], -1)
    ref = data.astype(np.float64) @ flat.astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(flat).astype(np.float64).T).max() + 1e-30
    assert np.abs(out - ref).max() / scale <= 3e-6
    banded = gm.group_masks(t, plan, kernel='banded').cpu().numpy()
    assert np.abs(out - banded).max() / scale <= 3e-6
    out2 = gm.group_masks(t, plan, out=torch.from_numpy(out).cuda(), accumulate=True,
                          kernel='sym').cpu().numpy()
    assert np.abs(out2 - 2 * ref).max() / scale <= 6e-6


def test_split_table_layout():
    from libertem_b200 import group_masks as gm
    rng = np.random.default_rng(5)
    t = (rng.random((25, 64, 2)) - 0.5).astype(np.float32)
    s = gm.split_table(t, 112)
    assert s.shape == (112, 64)
    for r in (0, 1, 27, 28, 49):
        h, j = divmod(r, 28)
        hi, lo = s[h * 56 + j], s[h * 56 + 28 + j]
        assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)
        want = t[r // 2, :, r % 2]
        assert np.abs(hi.astype(np.float64) + lo - want).max() <= 2.0 ** -22 * np.abs(want).max()
    assert not s[56 + 22:56 + 28].any() and not s[56 + 28 + 22:].any()


def test_find_groups():
    from libertem_b200 import group_masks as gm
    s = make_stack(4, 5, 200, seed=3)
    assert gm.find_groups(s) == 5
    s[7, :] = 1           # breaks the uniform structure
    assert gm.find_groups(s) is None
    packed = gm.pack_rows(np.arange(2 * 64 * 2, dtype=np.float32).reshape(2, 64, 2))
    # entry 4q+c of block 0, component w -> (c//2)*32 + q*4 + (c%2)*2 + w
    for (q, c, w) in [(0, 0, 0), (3, 1, 1), (7, 2, 0), (5, 3, 1)]:
        assert packed[0, (c // 2) * 32 + q * 4 + (c % 2) * 2 + w] == (4 * q + c) * 2 + w


def _shifted_ref(data, masks, shifts):
    F, sy, sx = data.shape
    out = np.zeros((F, masks.shape[0]))
    for f in range(F):
        dy, dx = shifts[f if len(shifts) > 1 else 0]
        y0, y1 = max(0, dy), min(sy, sy + dy)
        x0, x1 = max(0, dx), min(sx, sx + dx)
        if y1 <= y0 or x1 <= x0:
            continue
        d = data[f, y0:y1, x0:x1].astype(np.float64)
        m = masks[:, y0 - dy:y1 - dy, x0 - dx:x1 - dx].astype(np.float64)
        out[f] = (m * d).sum(axis=(1, 2))
    return out


@pytest.mark.parametrize('dt', [np.float32, np.uint16])
def test_masks_shifted_kernel(dt):
    from libertem_b200 import engine
    F, sy, sx, M = 50, 24, 40, 6
    if dt == np.float32:
        data = synth.uniform_f32(0, F * sy * sx, 5).reshape(F, sy, sx)
        t = torch.from_numpy(data).cuda()
    else:
        data = synth.poisson3_u16(0, F * sy * sx, 5).reshape(F, sy, sx)
        t = torch.from_numpy(data.view(np.int16)).cuda().view(torch.uint16)
    masks = synth.uniform_f32(0, M * sy * sx, 6).reshape(M, sy, sx) - 0.3
    sh = (synth.hash_u32(0, 2 * F, 7) % 31).astype(np.int64).reshape(F, 2) - 15
    sh[0] = (0, 0)
    sh[1] = (100, 0)        # no overlap
    sh[2] = (-23, 39)       # single pixel
    out = engine.masks_shifted(t, torch.from_numpy(masks.reshape(M, -1)).cuda(),
                               torch.from_numpy(sh)).cpu().numpy()
    ref = _shifted_ref(data, masks, sh)
    scale = np.abs(ref).max() + 1e-30
    assert np.abs(out - ref).max() / scale <= 2e-6
    assert np.all(out[1] == 0)
    const = engine.masks_shifted(t, torch.from_numpy(masks.reshape(M, -1)).cuda(),
                                 torch.tensor([[3, -7]])).cpu().numpy()
    assert np.abs(const - _shifted_ref(data, masks, [(3, -7)])).max() / scale <= 2e-6


@pytest.mark.parametrize('dt', [np.float32, np.uint16])
@pytest.mark.parametrize('M,D', [(3, 4), (8, 15), (11, 2), (6, 40)])
def test_masks_shifted_banded_kernel(dt, M, D):
    """K5 banded form (mask row band + |dy| halo in shared memory, band partial sums): same
    numbers as the float64 evaluation of the reference's shifted-mask semantics
    (udf/masks.py:85-124), incl. frames without overlap, accumulate and ragged frame chunks"""
    from libertem_b200 import engine
    F, sy, sx = 300, 48, 40
    if dt == np.float32:
        data = synth.uniform_f32(0, F * sy * sx, 5).reshape(F, sy, sx)
        t = torch.from_numpy(data).cuda()
    else:
        data = synth.poisson3_u16(0, F * sy * sx, 5).reshape(F, sy, sx)
        t = torch.from_numpy(data.view(np.int16)).cuda().view(torch.uint16)
    masks = synth.uniform_f32(0, M * sy * sx, 6).reshape(M, sy, sx) - 0.3
    sh = (synth.hash_u32(0, 2 * F, 7) % (2 * D + 1)).astype(np.int64).reshape(F, 2) - D
    sh[:, 1] = (synth.hash_u32(0, F, 8) % 61).astype(np.int64) - 30      # dx: unrestricted
    sh[0] = (0, 0)
    sh[1] = (D, 100)         # no overlap in x
    sh[2] = (-D, sx - 1)     # single column
    mt = torch.from_numpy(masks.reshape(M, -1)).cuda()
    out = engine.masks_shifted(t, mt, torch.from_numpy(sh)).cpu().numpy()
    assert engine.last_kernel() == 50
    ref = _shifted_ref(data, masks, sh)
    scale = np.abs(ref).max() + 1e-30
    assert np.abs(out - ref).max() / scale <= 2e-6
    assert np.all(out[1] == 0)
    generic = engine.masks_shifted(t, mt, torch.from_numpy(sh), banded=False).cpu().numpy()
    assert engine.last_kernel() == 5
    assert np.abs(out - generic).max() / scale <= 2e-6
    acc = torch.from_numpy(out).cuda()
    engine.masks_shifted(t, mt, torch.from_numpy(sh), out=acc, accumulate=True)
    assert np.abs(acc.cpu().numpy() - 2 * ref).max() / scale <= 4e-6
    const = engine.masks_shifted(t, mt, torch.tensor([[3, -7]])).cpu().numpy()
    assert engine.last_kernel() == 50
    assert np.abs(const - _shifted_ref(data, masks, [(3, -7)])).max() / scale <= 2e-6


def _radial_flat(S, n_bins, max_order, **kw):
    from libertem_b200 import masks as M
    from libertem_b200.analysis.radialfourier import radial_mask_factory
    cx, cy = kw.get('cx', S / 2), kw.get('cy', S / 2)
    ro = kw.get('ro', M.bounding_radius(cx, cy, S, S))
    st = np.asarray(radial_mask_factory(S, S, cx, cy, kw.get('ri', 0), ro, n_bins, max_order,
                                        use_sparse=False)())
    return st.reshape(st.shape[0], -1).astype(np.complex64)


@pytest.mark.parametrize('S,n_bins,max_order,F,kw', [
    (64, 4, 6, 128, {}), (128, 8, 24, 300, {}), (128, 8, 24, 1000, {}), (256, 16, 24, 257, {}),
    (128, 5, 12, 500, dict(cx=60.5, cy=70.25, ri=6.0, ro=50.0)), (128, 8, 7, 2000, {})])
def test_group_masks_walk(S, n_bins, max_order, F, kw):
    """K10 (dense-walk plan, ltb200_group_masks_walk) on the reference's radial masks: numpy
    float64, the banded K7 plan, accumulate, strided tiles, ragged last frame block; the result
    does not depend on the launch (atomics with <= 2 addends per element)"""
    from libertem_b200 import engine, group_masks as gm
    flat = _radial_flat(S, n_bins, max_order, **kw)
    plan = gm.build_plan(flat, max_order + 1, torch.device('cuda'))
    assert plan.walk is not None
    data = synth.uniform_f32(0, F * S * S, 21).reshape(F, S * S)
    t = torch.from_numpy(data).cuda()
    out = gm.group_masks(t, plan, kernel='auto').cpu().numpy()
    assert engine.last_kernel() == (10 if F >= gm.TC_MIN_FRAMES else 4)
    out = gm.group_masks(t, plan, kernel='walk').cpu().numpy()
    assert engine.last_kernel() == 10
    ref = data.astype(np.float64) @ flat.astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(flat).astype(np.float64).T).max() + 1e-30
    assert out.shape == ref.shape and out.dtype == np.complex64
    assert np.abs(out - ref).max() / scale <= 2e-6
    banded = gm.group_masks(t, plan, kernel='banded').cpu().numpy()
    assert np.abs(out - banded).max() / scale <= 3e-6
    out2 = gm.group_masks(t, plan, out=torch.from_numpy(out).cuda(), accumulate=True,
                          kernel='walk').cpu().numpy()
    assert np.abs(out2 - 2 * ref).max() / scale <= 4e-6
    big = torch.zeros((F, S * S + 24), device='cuda')
    big[:, 8:8 + S * S] = t
    for _ in range(3):
        out3 = gm.group_masks(big[:, 8:8 + S * S], plan, kernel='walk').cpu().numpy()
        assert np.array_equal(out3, out)


@pytest.mark.parametrize('n_groups,size,K,F', [(3, 25, 512, 200), (8, 7, 2048, 129),
                                               (12, 4, 4096, 1000), (1, 5, 256, 128)])
def test_group_masks_walk_bands(n_groups, size, K, F):
    """K10 on stacks that are not rings (overlapping pixel bands; empty first group)"""
    from libertem_b200 import engine, group_masks as gm
    rng = np.random.default_rng(K)
    stack = np.zeros((n_groups * size, K), dtype=np.complex64)
    width = 96
    step = (K - width) // max(1, n_groups - 1) if n_groups > 1 else 0
    for g in range(n_groups):
        if g == 0 and n_groups > 4:
            continue
        a = g * step
        b = min(K, a + width)
        stack[g * size:(g + 1) * size, a:b] = (rng.random((size, b - a)) - 0.5 +
                                               1j * (rng.random((size, b - a)) - 0.5))
    plan = gm.build_plan(stack, size, torch.device('cuda'))
    assert plan.walk is not None
    data = synth.uniform_f32(0, F * K, 22).reshape(F, K)
    t = torch.from_numpy(data).cuda()
    out = gm.group_masks(t, plan, kernel='walk').cpu().numpy()
    assert engine.last_kernel() == 10
    ref = data.astype(np.float64) @ stack.astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(stack).astype(np.float64).T).max() + 1e-30
    assert np.abs(out - ref).max() / scale <= 2e-6


@pytest.mark.parametrize('max_order', [6, 7])
def test_group_masks_walk_narrow_rings_repeated(max_order):
    """regression: boxes with a single op hand their shared-memory stage back almost at once;
    the converters used to release it before their loads had returned (an empty asm is no
    scoreboard wait) and the TMA refill corrupted rows of the outermost rings in 5-50 % of the
    launches of exactly this geometry (scripts/k10_stress.py)"""
    from libertem_b200 import engine, group_masks as gm
    S, F = 128, 2000
    flat = _radial_flat(S, 16, max_order)
    plan = gm.build_plan(flat, max_order + 1, torch.device('cuda'), walk_max_dup=1e9)
    assert plan.walk is not None
    data = synth.uniform_f32(0, F * S * S, 23).reshape(F, S * S)
    t = torch.from_numpy(data).cuda()
    ref = data.astype(np.float64) @ flat.astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(flat).astype(np.float64).T).max() + 1e-30
    first = None
    for _ in range(25):
        out = gm.group_masks(t, plan, kernel='walk').cpu().numpy()
        assert engine.last_kernel() == 10
        assert np.abs(out - ref).max() / scale <= 2e-6
        first = out if first is None else first
        assert np.array_equal(out, first)


def test_group_masks_walk_rejects():
    from libertem_b200 import _lib, group_masks as gm
    flat = _radial_flat(128, 16, 6)              # narrow rings: no walk plan by default
    plan = gm.build_plan(flat, 7, torch.device('cuda'))
    assert plan.walk is None and plan.banded is not None
    t = torch.zeros((128, 128 * 128), device='cuda')
    with pytest.raises(_lib.LTB200Error):
        gm.group_masks(t, plan, kernel='walk')
    out = gm.group_masks(t, plan, kernel='auto')
    assert float(out.abs().max()) == 0.0
