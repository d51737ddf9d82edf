"""Pins the CPU oracle (oracle/) against outputs of the unmodified reference (tests/golden/).

CPU-only.  Tolerances: bit-exact for integer-valued data; 1e-6 rel for float32 (the oracle
uses the same BLAS call as the reference, differences come only from partition/threading).
"""
import numpy as np
import pytest

from conftest import load_golden
from golden_inputs import mixed_masks, roi_from_seed, ring_stack
from oracle import synth, masks_gen, udf_oracle as O

RTOL = 2e-6


def test_synth_thresholds_frozen():
    assert np.array_equal(synth._poisson_thresholds(3.0)[:20], synth.POISSON3_THRESHOLDS)
    a = synth.uniform_f32(0, 1000, 5)
    assert a.dtype == np.float32 and a.min() >= 0 and a.max() < 1
    # slices are consistent with the whole
    assert np.array_equal(synth.uniform_f32(100, 50, 5), a[100:150])
    big = (1 << 32) + 17
    assert np.array_equal(synth.uniform_f32(big, 8, 5)[3:], synth.uniform_f32(big + 3, 5, 5))


def test_masks_gen():
    _, g = load_golden('masks_gen')
    assert np.array_equal(masks_gen.circular(7.3, 5.1, 20, 16, 4.6), g['circular'])
    assert np.array_equal(masks_gen.ring(9, 8, 20, 16, 7.5, 3.2), g['ring'])
    assert np.array_equal(masks_gen.gradient_x(20, 16), g['gradient_x'])
    assert np.array_equal(masks_gen.gradient_y(20, 16), g['gradient_y'])
    rb = masks_gen.radial_bins(9.5, 8.2, 20, 16, radius=9., radius_inner=0, n_bins=4,
                               dtype=np.float32)
    assert np.array_equal(rb, g['radial_bins'])
    rb2 = masks_gen.radial_bins(10, 8, 20, 16, n_bins=5, dtype=np.float64)
    assert np.array_equal(rb2, g['radial_bins_default'])
    r, phi = masks_gen.polar_map(9.5, 8.2, 20, 16)
    assert np.array_equal(r, g['polar_r']) and np.array_equal(phi, g['polar_phi'])
    assert masks_gen.bounding_radius(9.5, 8.2, 20, 16) == int(g['bounding_radius'])


@pytest.mark.parametrize('nparts', [1, 8])
def test_cfg1(nparts):
    meta, g = load_golden(f'cfg1_p{nparts}')
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed'])
    mask = synth.uniform_f32(0, 64 * 64, meta['mask_seed']).reshape(1, 64, 64)
    res = O.apply_masks(data, mask, num_partitions=nparts)
    assert res.dtype == g['intensity'].dtype and res.shape == g['intensity'].shape
    np.testing.assert_allclose(res, g['intensity'], rtol=RTOL)


def test_cfg2_small():
    meta, g = load_golden('cfg2_small')
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed'])
    stack = mixed_masks(256, 256, meta['n_masks'], meta['mask_seed'])
    P = meta['num_partitions']
    res = O.apply_masks(data, stack, num_partitions=P)
    np.testing.assert_allclose(res, g['intensity'], rtol=RTOL, atol=1e-3)
    com = O.com_udf(data, num_partitions=P)
    np.testing.assert_allclose(com['raw_mask_result'], g['com_raw_mask_result'], rtol=RTOL)
    for k in ('raw_com', 'raw_shifts', 'field', 'field_y', 'field_x', 'magnitude',
              'divergence', 'curl', 'regression'):
        assert com[k].dtype == g['com_' + k].dtype, k
        assert com[k].shape == g['com_' + k].shape, k
        # shifts lose ~7 bits to cancellation: absolute tolerance
        np.testing.assert_allclose(com[k], g['com_' + k], rtol=1e-5, atol=2e-4, err_msg=k)
    np.testing.assert_allclose(O.sum_udf(data, num_partitions=P), g['sum'], rtol=RTOL)
    np.testing.assert_allclose(O.sumsig_udf(data, num_partitions=P), g['sumsig'], rtol=RTOL)


def test_cfg3_small_bit_exact():
    meta, g = load_golden('cfg3_small')
    data = synth.dataset(meta['shape'], np.uint16, meta['data_seed'])
    stack = ring_stack((128, 128), meta['rings'], 64, 64)
    P = meta['num_partitions']
    ts = (32, 64, 128)   # the reference's negotiated tile shape for this case
    res = O.apply_masks(data, stack, num_partitions=P, tileshape=ts, mask_dtype=np.float32,
                        use_sparse=True)
    assert res.dtype == np.float32
    assert np.array_equal(res, g['intensity'])
    assert np.array_equal(O.sum_udf(data, num_partitions=P, tileshape=ts), g['sum'])
    assert np.array_equal(O.sumsig_udf(data, num_partitions=P, tileshape=ts), g['sumsig'])
    # and against plain integer arithmetic
    exact = data.reshape(256, -1).astype(np.int64) @ stack.reshape(4, -1).T.astype(np.int64)
    assert np.array_equal(res.astype(np.int64), exact)


@pytest.mark.parametrize('kind', ['sparse', 'dense'])
def test_cfg4_small(kind):
    meta, g = load_golden('cfg4_small_' + kind)
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed'])
    p = meta['params']
    raw = O.radial_fourier(data, num_partitions=meta['num_partitions'],
                           use_sparse=(kind == 'sparse'),
                           cx=p['cx'], cy=p['cy'], ri=p['ri'], ro=p['ro'], n_bins=p['n_bins'],
                           max_order=p['max_order'])
    assert raw.shape == g['raw_results'].shape and raw.dtype == np.complex64
    scale = np.abs(g['raw_results'][:, 0]).max()
    assert np.abs(raw - g['raw_results']).max() <= 1e-5 * scale
    if kind == 'dense':
        stack = O.radial_mask_stack((64, 64), p['cx'], p['cy'], p['ri'], p['ro'], p['n_bins'],
                                    p['max_order'])
        np.testing.assert_allclose(stack[::37], g['mask_stack_sub'], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('i', [0, 1, 2])
def test_com_params(i):
    meta, g = load_golden(f'com_params_{i}')
    c = meta['com']
    if c['r'] == 'inf':
        c['r'] = float('inf')
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed'])
    com = O.com_udf(data, num_partitions=meta['num_partitions'], **c)
    np.testing.assert_allclose(com['raw_mask_result'], g['raw_mask_result'], rtol=RTOL)
    for k in ('raw_com', 'raw_shifts', 'field', 'field_y', 'field_x', 'magnitude',
              'divergence', 'curl', 'regression'):
        assert com[k].dtype == g[k].dtype and com[k].shape == g[k].shape, k
        np.testing.assert_allclose(com[k], g[k], rtol=1e-5, atol=5e-5, err_msg=k)


def test_roi():
    meta, g = load_golden('roi')
    shape = meta['shape']
    data = synth.dataset(shape, np.float32, meta['data_seed'])
    roi = roi_from_seed(shape[:2], meta['roi_seed'])
    stack = mixed_masks(shape[2], shape[3], meta['n_masks'], meta['mask_seed'])
    P = meta['num_partitions']
    res = O.apply_masks(data, stack, num_partitions=P, roi=roi)
    np.testing.assert_allclose(res, g['intensity'], rtol=RTOL, atol=1e-4)
    com = O.com_udf(data, num_partitions=P, regression=1, roi=roi)
    np.testing.assert_allclose(com['raw_mask_result'], g['com_raw_mask_result'], rtol=RTOL)
    for k in ('raw_com', 'raw_shifts', 'field', 'magnitude', 'divergence', 'curl', 'regression'):
        assert com[k].shape == g['com_' + k].shape, k
        np.testing.assert_allclose(com[k], g['com_' + k], rtol=1e-5, atol=5e-5, err_msg=k,
                                   equal_nan=True)
    np.testing.assert_allclose(O.sum_udf(data, num_partitions=P, roi=roi), g['sum'], rtol=RTOL)
    np.testing.assert_allclose(O.sumsig_udf(data, num_partitions=P, roi=roi), g['sumsig'],
                               rtol=RTOL)


@pytest.mark.parametrize('dt', ['float32', 'uint16'])
def test_odd(dt):
    meta, g = load_golden('odd_' + dt)
    shape = meta['shape']
    data = synth.dataset(shape, np.dtype(dt), meta['data_seed'])
    stack = mixed_masks(shape[2], shape[3], meta['n_masks'], meta['mask_seed'])
    kw = dict(num_partitions=meta['num_partitions'], tileshape=(4, 17, 23))
    np.testing.assert_allclose(O.apply_masks(data, stack, **kw), g['intensity'], rtol=RTOL,
                               atol=1e-4)
    com = O.com_udf(data, **kw)
    np.testing.assert_allclose(com['raw_mask_result'], g['com_raw_mask_result'], rtol=RTOL)
    np.testing.assert_allclose(com['raw_com'], g['com_raw_com'], rtol=1e-5)
    np.testing.assert_allclose(O.sum_udf(data, **kw), g['sum'], rtol=RTOL)
    np.testing.assert_allclose(O.sumsig_udf(data, **kw), g['sumsig'], rtol=RTOL)
    if dt == 'uint16':
        assert np.array_equal(O.sum_udf(data, **kw), g['sum'])
        assert np.array_equal(O.sumsig_udf(data, **kw), g['sumsig'])


def test_subframe_tiles():
    meta, g = load_golden('subframe')
    shape = meta['shape']
    data = synth.dataset(shape, np.float32, meta['data_seed'])
    stack = mixed_masks(shape[2], shape[3], meta['n_masks'], meta['mask_seed'])
    kw = dict(num_partitions=meta['num_partitions'], tileshape=tuple(meta['tileshape']))
    np.testing.assert_allclose(O.apply_masks(data, stack, **kw), g['intensity'], rtol=RTOL,
                               atol=1e-4)
    np.testing.assert_allclose(O.sum_udf(data, **kw), g['sum'], rtol=RTOL)
    np.testing.assert_allclose(O.sumsig_udf(data, **kw), g['sumsig'], rtol=RTOL)


def test_dtype_rules():
    meta, g = load_golden('dtypes')
    shape = meta['shape']
    n = int(np.prod(shape))
    stack32 = mixed_masks(8, 8, 2, meta['mask_seed'])
    d_i32 = (synth.hash_u32(0, n, meta['seeds']['i32']) % 1000).astype(np.int32).reshape(shape)
    r = O.apply_masks(d_i32, stack32, num_partitions=2)
    assert r.dtype == np.float64
    np.testing.assert_allclose(r, g['i32_f32'], rtol=1e-12)
    d_f32 = synth.dataset(shape, np.float32, meta['seeds']['f32'])
    stack64 = stack32.astype(np.float64) * 1.000000123
    r = O.apply_masks(d_f32, stack64, num_partitions=2)
    assert r.dtype == np.float64
    np.testing.assert_allclose(r, g['f32_f64'], rtol=1e-12)
    r = O.apply_masks(d_f32, stack64, num_partitions=2, mask_dtype=np.float32)
    assert r.dtype == np.float32
    np.testing.assert_allclose(r, g['f32_f64_forced32'], rtol=RTOL)
    stackc = (stack32 + 1j * stack32[::-1]).astype(np.complex64)
    r = O.apply_masks(d_f32, stackc, num_partitions=2)
    assert r.dtype == np.complex64
    np.testing.assert_allclose(r, g['f32_c64'], rtol=RTOL)
    d_u8 = (synth.hash_u32(0, n, meta['seeds']['u8']) % 256).astype(np.uint8).reshape(shape)
    assert np.array_equal(O.apply_masks(d_u8, stack32, num_partitions=2), g['u8_f32']) or \
        np.allclose(O.apply_masks(d_u8, stack32, num_partitions=2), g['u8_f32'], rtol=RTOL)
    assert np.array_equal(O.sum_udf(d_u8, num_partitions=2), g['u8_sum'])
    assert np.array_equal(O.sumsig_udf(d_u8, num_partitions=2), g['u8_sumsig'])


def test_rmatmul():
    import scipy.sparse as sp
    meta, g = load_golden('rmatmul')
    left = synth.uniform_f32(0, 37 * 300, meta['left_seed']).reshape(37, 300)
    dense = synth.uniform_f32(0, 300 * 6, meta['right_seed']).reshape(300, 6)
    dense[synth.hash_u32(0, 1800, meta['sel_seed']).reshape(300, 6) % 5 != 0] = 0
    np.testing.assert_allclose(O.rmatmul(left, sp.csr_matrix(dense)), g['csr'], rtol=RTOL)
    np.testing.assert_allclose(O.rmatmul(left, sp.csc_matrix(dense)), g['csc'], rtol=RTOL)
    with pytest.raises(ValueError):
        O.rmatmul(left[:, :10], sp.csr_matrix(dense))


@pytest.mark.parametrize('dt', ['float32', 'uint16'])
def test_corrections(dt):
    """oracle restatement of the detector corrections AND the product's mask folding
    (libertem_b200/corrections.py, pure host math) against the reference's outputs"""
    from oracle import corrections as OC
    from libertem_b200.corrections import CorrectionSet
    meta, g = load_golden('corrections')
    shape = meta['shape']
    stack = mixed_masks(16, 12, 3, meta['mask_seed'])
    dark = synth.uniform_f32(0, 192, meta['dark_seed']).reshape(16, 12) * 0.3
    gain = 0.5 + synth.uniform_f32(0, 192, meta['gain_seed']).reshape(16, 12)
    excl = np.zeros((16, 12), dtype=bool)
    for y, x in meta['excluded']:
        excl[y, x] = True
    data = synth.dataset(shape, np.dtype(dt), meta['seeds'][dt])
    for name, kw in (('dg', dict(dark=dark, gain=gain)),
                     ('dge', dict(dark=dark, gain=gain, excluded_mask=excl)),
                     ('e', dict(excluded_mask=excl))):
        key = f'{dt}_{name}_'
        c = OC.correct(data, **kw)
        np.testing.assert_allclose(O.apply_masks(c, stack, num_partitions=2),
                                   g[key + 'intensity'], rtol=RTOL, atol=1e-4)
        np.testing.assert_allclose(O.sum_udf(c, num_partitions=2), g[key + 'sum'], rtol=1e-5)
        np.testing.assert_allclose(O.sumsig_udf(c, num_partitions=2), g[key + 'sumsig'],
                                   rtol=1e-5)
        np.testing.assert_allclose(O.com_udf(c, num_partitions=2)['raw_com'],
                                   g[key + 'raw_com'], rtol=1e-5)
        cs = CorrectionSet(dark=kw.get('dark'), gain=kw.get('gain'),
                           excluded_pixels=kw.get('excluded_mask'))
        rows, const = cs.fold_masks(stack.reshape(3, -1))
        folded = data.reshape(20, -1).astype(np.float64) @ rows.T + const
        scale = np.abs(g[key + 'intensity']).max(axis=0)
        assert (np.abs(folded - g[key + 'intensity']) / scale).max() <= 1e-6
        fs = cs.correct_frame_sum(data.reshape(20, -1).astype(np.float64).sum(0), 20)
        np.testing.assert_allclose(fs.reshape(16, 12), g[key + 'sum'], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('name', ['u16', 'u8'])
def test_int_detector(name):
    """integer detectors (uint16 / uint8 frames, binary masks, CoM with a disk): the oracle
    reproduces the unmodified reference bit for bit -- every sum is an exact integer in
    float32 -- which is what the int8 tensor-core path of the product is held to"""
    from golden_inputs import int_detector_inputs
    meta, g = load_golden('int_detector_' + name)
    data, stack = int_detector_inputs(name, meta)
    P = meta['num_partitions']
    assert np.array_equal(O.sum_udf(data, num_partitions=P), g['sum'])
    assert np.array_equal(O.sumsig_udf(data, num_partitions=P), g['sumsig'])
    assert np.array_equal(O.apply_masks(data, stack, num_partitions=P), g['intensity'])
    com = O.com_udf(data, num_partitions=P, **meta['com'])
    assert np.array_equal(com['raw_mask_result'], g['com_raw_mask_result'])
    for k in ('raw_com', 'raw_shifts', 'field', 'field_y', 'field_x', 'magnitude',
              'divergence', 'curl', 'regression'):
        assert com[k].dtype == g['com_' + k].dtype and com[k].shape == g['com_' + k].shape, k
        np.testing.assert_allclose(com[k], g['com_' + k], rtol=1e-6, atol=1e-6, err_msg=k)


def _complex_inputs(meta):
    shape = tuple(meta['shape'])
    s0, s1 = meta['seeds']
    data = (synth.dataset(shape, np.float32, s0)
            + 1j * (synth.dataset(shape, np.float32, s1) - 0.5)).astype(np.complex64)
    m0, m1, m2 = meta['mask_seeds']
    real_masks = mixed_masks(shape[2], shape[3], 3, m0)
    cmasks = (mixed_masks(shape[2], shape[3], 2, m1)
              + 1j * mixed_masks(shape[2], shape[3], 2, m2)).astype(np.complex64)
    return data, real_masks, cmasks


def test_complex_input():
    """complex64 frames: result_type(input, mask) is complex64, numpy `@` path
    (udf/masks.py:76-77,360-368)"""
    meta, g = load_golden('complex_input')
    data, real_masks, cmasks = _complex_inputs(meta)
    for name, kw in (('p2', dict(num_partitions=2)),
                     ('tiled', dict(num_partitions=3, tileshape=(5, 8, 32)))):
        for key, masks in (('real_masks_', real_masks), ('complex_masks_', cmasks)):
            got = O.apply_masks(data, masks, **kw)
            ref = g[key + name]
            assert got.dtype == ref.dtype == np.complex64
            assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max()
    raw = O.apply_masks(data, O.com_mask_stack(data.shape[-2:], 0, 0), num_partitions=2)
    assert np.abs(raw - g['com_raw']).max() <= RTOL * np.abs(g['com_raw']).max()
    img = raw.reshape(data.shape[:2] + (3,))
    y, x = O.center_shifts(img[..., 0], img[..., 1], img[..., 2], 0, 0)
    y, x = O.apply_correction(y, x, 0., False)
    np.testing.assert_allclose(x, g['com_x'], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(y, g['com_y'], rtol=1e-4, atol=1e-4)
