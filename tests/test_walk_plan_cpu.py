"""CPU tests of the dense-walk plan of K10 (libertem_b200/walk_plan.py): the numpy emulation of
the kernel's lists reproduces the direct sums, and the invariants the kernel relies on hold."""
import numpy as np
import pytest

from libertem_b200 import masks as M
from libertem_b200 import walk_plan as wp
from libertem_b200.analysis.radialfourier import radial_mask_factory


def radial_stack(S, n_bins, max_order, cx=None, cy=None, ri=0, ro=None):
    cx = S / 2 if cx is None else cx
    cy = S / 2 if cy is None else cy
    ro = M.bounding_radius(cx, cy, S, S) if ro is None else ro
    st = np.asarray(radial_mask_factory(S, S, cx, cy, ri, ro, n_bins, max_order,
                                        use_sparse=False)())
    return st.reshape(st.shape[0], -1).astype(np.complex64)


def band_stack(n_groups, size, K, seed, width=96):
    """groups = overlapping contiguous pixel bands (<= 2 groups on a slice, every box once)"""
    rng = np.random.default_rng(seed)
    stack = np.zeros((n_groups * size, K), dtype=np.complex64)
    step = (K - width) // max(1, n_groups - 1) if n_groups > 1 else 0
    for g in range(n_groups):
        a = g * step
        b = min(K, a + width)
        vals = rng.random((size, b - a)) - 0.5 + 1j * (rng.random((size, b - a)) - 0.5)
        stack[g * size:(g + 1) * size, a:b] = vals.astype(np.complex64)
    return stack


def check(plan, flat, F=3, seed=0):
    rng = np.random.default_rng(seed)
    tile = rng.random((F, flat.shape[1])).astype(np.float32)
    res = wp.emulate(plan, tile)
    ref = tile.astype(np.float64) @ flat.astype(np.complex128).T
    scale = np.abs(tile.astype(np.float64)) @ np.abs(flat).astype(np.float64).T
    assert (np.abs(res - ref) / np.maximum(scale, 1e-30)).max() < 5e-8   # hi + lo of the weights


@pytest.mark.parametrize('S,n_bins,max_order,kw', [
    (64, 4, 6, {}), (128, 8, 24, {}), (128, 5, 12, dict(cx=60.5, cy=70.25, ri=6.0, ro=50.0)),
    (96, 6, 3, {})])
def test_walk_plan_radial(S, n_bins, max_order, kw):
    flat = radial_stack(S, n_bins, max_order, **kw)
    plan = wp.build_walk(flat, max_order + 1)
    assert plan is not None
    # what the C ABI documents (include/ltb200.h, ltb200_group_masks_walk)
    assert np.all(np.diff(plan['visit_off']) % wp.BOX_PAD == 0)
    for c in range(4):
        assert np.all(plan[f'op_off{c}'] % wp.STAGE_OPS == 0)
        assert len(plan[f'ops{c}']) == plan[f'op_off{c}'][-1]
        assert plan[f'table{c}'].shape == (plan[f'tab_off{c}'][-1], wp.STAGE_ROWS, 32)
        assert plan[f'table{c}'].dtype == np.float32
        # the table holds only TF32 numbers (low 13 bits zero)
        assert not np.any(plan[f'table{c}'].view(np.uint32) & np.uint32(0x1FFF))
    assert np.all((plan['boxes'] & ~np.uint32(31)) < flat.shape[1])
    check(plan, flat)


@pytest.mark.parametrize('n_groups,size,K,seed', [(3, 25, 512, 1), (8, 7, 2048, 2), (2, 28, 320, 3),
                                                  (12, 4, 4096, 4), (1, 5, 256, 5)])
def test_walk_plan_bands(n_groups, size, K, seed):
    flat = band_stack(n_groups, size, K, seed)
    plan = wp.build_walk(flat, size)
    assert plan is not None
    check(plan, flat, seed=seed)


@pytest.mark.parametrize('seed', range(4))
def test_walk_plan_generic_stacks(seed):
    """random supports (boxes visited once per window of 4 groups, many groups per slice)"""
    rng = np.random.default_rng(seed)
    n_groups, size, K = int(rng.integers(2, 10)), int(rng.integers(1, 9)), 32 * int(rng.integers(4, 24))
    stack = np.zeros((n_groups * size, K), dtype=np.complex64)
    for g in range(n_groups):
        px = np.sort(rng.choice(K, size=int(rng.integers(1, K // 2)), replace=False))
        stack[g * size:(g + 1) * size, px] = (rng.random((size, len(px))) - 0.5 +
                                              1j * (rng.random((size, len(px))) - 0.5))
    plan = wp.build_walk(stack, size, max_dup=np.inf)
    assert plan is not None
    check(plan, stack, seed=seed)


def test_walk_plan_gate():
    """narrow rings (boxes that span more than 4 groups would be visited > 1.5 times on average)
    are not admitted by default: those stacks stay on K7"""
    flat = radial_stack(128, 16, 6)
    assert wp.build_walk(flat, 7) is None
    assert wp.build_walk(flat, 7, max_per_slice=2, max_dup=np.inf) is None
    plan = wp.build_walk(flat, 7, max_dup=np.inf)
    assert plan is not None
    check(plan, flat)
    assert wp.build_walk(radial_stack(64, 4, 6)[:, :4010], 7) is None      # K % 32
    assert wp.build_walk(np.zeros((58, 64), np.complex64), 29) is None     # > 28 columns
