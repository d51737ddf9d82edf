"""UDF-level parity on the GPU: the host mirror of the reference API (ApplyMasksUDF, CoMUDF,
SumUDF, SumSigUDF, analyses) against golden vectors produced by the unmodified reference and
against the CPU oracle.  pytest -m gpu."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from golden_inputs import mixed_masks, roi_from_seed, ring_stack
from oracle import synth, udf_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5          # north_star tolerance for float32 mask / CoM results
COM_KEYS = ('raw_com', 'raw_shifts', 'field', 'field_y', 'field_x', 'magnitude', 'divergence',
            'curl', 'regression')


@pytest.fixture(scope='module')
def lt():
    import libertem_b200.udf as udf
    from libertem_b200.io import MemoryDataSet, SyntheticDataSet
    from libertem_b200.runner import run_udf, UDFRunner
    from libertem_b200.api import Context

    class NS:
        pass
    ns = NS()
    ns.udf, ns.MemoryDataSet, ns.SyntheticDataSet = udf, MemoryDataSet, SyntheticDataSet
    ns.run_udf, ns.UDFRunner, ns.Context = run_udf, UDFRunner, Context
    return ns


def close_cols(res, ref, rtol=RTOL):
    """per-column tolerance relative to the column's magnitude (masks may be signed)"""
    res = np.asarray(res, dtype=np.float64).reshape(ref.shape)
    ref64 = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref64).reshape(-1, ref.shape[-1]).max(axis=0) + 1e-30
    err = (np.abs(res - ref64).reshape(-1, ref.shape[-1]) / scale).max()
    assert err <= rtol, f'max rel err {err:.3e}'


def check_com(res, g, prefix='com_', atol=2e-4):
    for k in COM_KEYS:
        ref = g[prefix + k]
        got = res[k].raw_data
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        assert got.dtype == ref.dtype, (k, got.dtype, ref.dtype)
        # shifts/field lose ~7 bits to cancellation -> absolute tolerance
        np.testing.assert_allclose(got, ref, rtol=RTOL, atol=atol, err_msg=k, equal_nan=True)


@pytest.mark.parametrize('nparts', [1, 8])
@pytest.mark.parametrize('where', ['device', 'host'])
def test_cfg1(lt, nparts, where):
    meta, g = load_golden(f'cfg1_p{nparts}')
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed'])
    mask = synth.uniform_f32(0, 64 * 64, meta['mask_seed']).reshape(64, 64)
    src = torch.from_numpy(data).cuda() if where == 'device' else data
    ds = lt.MemoryDataSet(data=src, num_partitions=nparts, sig_dims=2)
    res = lt.run_udf(ds, lt.udf.ApplyMasksUDF(mask_factories=[lambda: mask]))
    inten = res['intensity']
    assert inten.data.shape == (32, 32, 1) and inten.data.dtype == np.float32
    close_cols(inten.raw_data, g['intensity'])


@pytest.mark.parametrize('fuse', [True, False])
def test_cfg2_small(lt, fuse):
    meta, g = load_golden('cfg2_small')
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed'])
    stack = mixed_masks(256, 256, meta['n_masks'], meta['mask_seed'])
    ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(),
                          num_partitions=meta['num_partitions'], sig_dims=2)
    runner = lt.UDFRunner([lt.udf.ApplyMasksUDF(mask_factories=lambda: stack),
                           lt.udf.CoMUDF(), lt.udf.SumUDF(), lt.udf.SumSigUDF()], fuse=fuse)
    from libertem_b200 import engine
    engine.launch_count(reset=True)
    res = runner.run_for_dataset(ds).buffers
    launches = engine.launch_count()
    if fuse:
        # ONE pass of the dense kernel per tile for all four UDFs (12 columns) -- not one pass
        # per UDF; helpers (mask pack, split-K finalize, column sum) are tiny launches
        assert runner.stats['unfused_calls'] == 0
        assert runner.stats['fused_launch_groups'] == runner.stats['tiles'] == \
            meta['num_partitions']
        # (+ the four nav-space launches of CoMUDF.get_results, K9)
        assert launches <= 5 * meta['num_partitions'] + 4, launches
    close_cols(res[0]['intensity'].raw_data, g['intensity'])
    assert 'raw_mask_result' not in res[1]          # private buffer stays private
    check_com(res[1], g)
    np.testing.assert_allclose(res[2]['intensity'].data, g['sum'], rtol=RTOL)
    np.testing.assert_allclose(res[3]['intensity'].raw_data, g['sumsig'], rtol=RTOL)
    assert res[3]['intensity'].data.shape == (16, 16)
    # raw_com to 1e-5 relative (north_star: CoM output within 1e-5 of reference)
    np.testing.assert_allclose(res[1]['raw_com'].raw_data, g['com_raw_com'], rtol=RTOL)


def test_cfg3_small_bit_exact(lt):
    meta, g = load_golden('cfg3_small')
    data = synth.dataset(meta['shape'], np.uint16, meta['data_seed'])
    rings = meta['rings']
    from libertem_b200 import masks as M
    facs = [lambda ri=ri, ro=ro: M.ring(64, 64, 128, 128, ro, ri) for ri, ro in rings]
    for src in (data, torch.from_numpy(data.view(np.int16)).cuda().view(torch.uint16)):
        ds = lt.MemoryDataSet(data=src, num_partitions=meta['num_partitions'], sig_dims=2)
        res = lt.run_udf(ds, [lt.udf.SumUDF(), lt.udf.SumSigUDF(),
                              lt.udf.ApplyMasksUDF(mask_factories=facs, use_sparse=True,
                                                   mask_dtype=np.float32)])
        assert res[0]['intensity'].data.dtype == np.float32
        assert np.array_equal(res[0]['intensity'].data, g['sum'])
        assert np.array_equal(res[1]['intensity'].raw_data, g['sumsig'])
        assert np.array_equal(res[2]['intensity'].raw_data, g['intensity'])


def test_cfg3_int8_tensor_path(lt):
    """cfg3 at a tile size that takes the integer fast path (uint16 tiles x binary masks on
    the int8 tensor cores, K8): SumUDF + SumSigUDF + 4 sparse ring masks in one pass, bit-exact
    against the oracle's restatement of the reference loops, with and without a ROI, and
    identical to the float kernels' results"""
    from libertem_b200 import masks as M, runner as R, engine
    shape = (48, 64, 128, 128)
    rings = [(8, 16), (20, 28), (32, 40), (44, 52)]
    data = synth.dataset(shape, np.uint16, 107)
    stack = ring_stack((128, 128), rings, 64, 64)
    facs = [lambda ri=ri, ro=ro: M.ring(64, 64, 128, 128, ro, ri) for ri, ro in rings]
    src = torch.from_numpy(data.view(np.int16)).cuda().view(torch.uint16)

    def run(roi=None):
        ds = lt.MemoryDataSet(data=src, num_partitions=2, sig_dims=2)
        runner = lt.UDFRunner([lt.udf.SumUDF(), lt.udf.SumSigUDF(),
                               lt.udf.ApplyMasksUDF(mask_factories=facs, use_sparse=True,
                                                    mask_dtype=np.float32)])
        res = runner.run_for_dataset(ds, roi=roi).buffers
        return runner, res

    runner, res = run()
    assert runner.stats.get('int8_passes', 0) == 2 and runner.stats['unfused_calls'] == 0
    assert engine.last_kernel() == 8
    assert np.array_equal(res[0]['intensity'].data, O.sum_udf(data, num_partitions=2))
    assert np.array_equal(res[1]['intensity'].raw_data, O.sumsig_udf(data, num_partitions=2))
    assert np.array_equal(res[2]['intensity'].raw_data,
                          O.apply_masks(data, stack, num_partitions=2, use_sparse=True,
                                        mask_dtype=np.float32))
    old = R.INT8_PATH
    R.INT8_PATH = False
    try:
        runner2, res2 = run()
    finally:
        R.INT8_PATH = old
    assert runner2.stats.get('int8_passes', 0) == 0
    for a, b in zip(res, res2):
        assert np.array_equal(a['intensity'].raw_data, b['intensity'].raw_data)
    roi = roi_from_seed(shape[:2], 108)
    _, res3 = run(roi)
    assert np.array_equal(res3[0]['intensity'].data, O.sum_udf(data, num_partitions=2, roi=roi))
    assert np.array_equal(res3[2]['intensity'].raw_data,
                          O.apply_masks(data, stack, num_partitions=2, use_sparse=True,
                                        mask_dtype=np.float32, roi=roi))


def test_com_u16_int8_tensor_path(lt):
    """CoMUDF + binary ApplyMasksUDF + SumSigUDF + SumUDF on a uint16 detector (256x256): the
    coordinate masks (weights up to 255) are split into two base-128 int8 digits and the whole
    group runs on the int8 tensor cores; sums below 2^24 are bit-exact, the rest within 1e-5"""
    from libertem_b200 import masks as M
    shape = (16, 64, 256, 256)
    data = synth.dataset(shape, np.uint16, 109)
    stack = np.stack([M.circular(128, 128, 256, 256, 40), M.ring(128, 128, 256, 256, 90, 60)])
    src = torch.from_numpy(data.view(np.int16)).cuda().view(torch.uint16)
    ds = lt.MemoryDataSet(data=src, num_partitions=2, sig_dims=2)
    runner = lt.UDFRunner([lt.udf.CoMUDF.with_params(cy=120, cx=131, r=100), lt.udf.SumSigUDF(),
                           lt.udf.ApplyMasksUDF(mask_factories=lambda: stack.astype(np.float32)),
                           lt.udf.SumUDF()])
    res = runner.run_for_dataset(ds).buffers
    assert runner.stats.get('int8_passes', 0) == 2 and runner.stats['unfused_calls'] == 0
    com = O.com_udf(data, cy=120, cx=131, r=100, num_partitions=2)
    raw = runner._udfs[0].results.get_buffer('raw_mask_result').raw_data
    assert np.array_equal(raw[:, 0], com['raw_mask_result'][:, 0])      # m00 < 2^24: exact
    np.testing.assert_allclose(raw, com['raw_mask_result'], rtol=1e-6)
    np.testing.assert_allclose(res[0]['raw_com'].raw_data, com['raw_com'], rtol=RTOL)
    np.testing.assert_allclose(res[0]['field'].raw_data, com['field'], rtol=RTOL, atol=2e-4)
    assert np.array_equal(res[1]['intensity'].raw_data, O.sumsig_udf(data, num_partitions=2))
    assert np.array_equal(res[2]['intensity'].raw_data,
                          O.apply_masks(data, stack.astype(np.float32), num_partitions=2))
    assert np.array_equal(res[3]['intensity'].data, O.sum_udf(data, num_partitions=2))
    # and against the unmodified reference (tests/golden/int_detector_u16.npz, same inputs)
    meta, g = load_golden('int_detector_u16')
    assert tuple(meta['shape']) == shape and meta['com'] == dict(cy=120, cx=131, r=100)
    assert np.array_equal(raw[:, 0], g['com_raw_mask_result'][:, 0])
    np.testing.assert_allclose(raw, g['com_raw_mask_result'], rtol=1e-6)
    np.testing.assert_allclose(res[0]['raw_com'].raw_data, g['com_raw_com'], rtol=RTOL)
    assert np.array_equal(res[1]['intensity'].raw_data, g['sumsig'])
    assert np.array_equal(res[2]['intensity'].raw_data, g['intensity'])
    assert np.array_equal(res[3]['intensity'].data, g['sum'])


def test_u8_int8_tensor_path(lt):
    """uint8 detector data: SumUDF + SumSigUDF + binary masks + CoM in one pass on the int8
    tensor cores, bit-exact against the oracle"""
    from libertem_b200 import masks as M
    shape = (16, 32, 64, 64)
    data = (synth.hash_u32(0, int(np.prod(shape)), 111) % 23).astype(np.uint8).reshape(shape)
    stack = np.stack([M.circular(32, 32, 64, 64, 10), M.ring(32, 32, 64, 64, 30, 20)])
    ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(), num_partitions=2, sig_dims=2)
    runner = lt.UDFRunner([lt.udf.SumUDF(), lt.udf.SumSigUDF(), lt.udf.CoMUDF(),
                           lt.udf.ApplyMasksUDF(mask_factories=lambda: stack.astype(np.float32))])
    res = runner.run_for_dataset(ds).buffers
    assert runner.stats.get('int8_passes', 0) == 2 and runner.stats['unfused_calls'] == 0
    assert np.array_equal(res[0]['intensity'].data, O.sum_udf(data, num_partitions=2))
    assert np.array_equal(res[1]['intensity'].raw_data, O.sumsig_udf(data, num_partitions=2))
    com = O.com_udf(data, num_partitions=2)
    raw = runner._udfs[2].results.get_buffer('raw_mask_result').raw_data
    assert np.array_equal(raw, com['raw_mask_result'])             # all sums < 2^24: exact
    np.testing.assert_allclose(res[2]['raw_com'].raw_data, com['raw_com'], rtol=RTOL)
    assert np.array_equal(res[3]['intensity'].raw_data,
                          O.apply_masks(data, stack.astype(np.float32), num_partitions=2))
    # and against the unmodified reference (tests/golden/int_detector_u8.npz, same inputs)
    meta, g = load_golden('int_detector_u8')
    assert tuple(meta['shape']) == shape
    assert np.array_equal(res[0]['intensity'].data, g['sum'])
    assert np.array_equal(res[1]['intensity'].raw_data, g['sumsig'])
    assert np.array_equal(raw, g['com_raw_mask_result'])
    np.testing.assert_allclose(res[2]['raw_com'].raw_data, g['com_raw_com'], rtol=RTOL)
    assert np.array_equal(res[3]['intensity'].raw_data, g['intensity'])


@pytest.mark.parametrize('kind', ['sparse', 'dense'])
def test_cfg4_small_radial_fourier(lt, kind):
    meta, g = load_golden('cfg4_small_' + kind)
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed'])
    ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(),
                          num_partitions=meta['num_partitions'], sig_dims=2)
    ctx = lt.Context()
    a = ctx.create_radial_fourier_analysis(ds, n_bins=8, use_sparse=(kind == 'sparse'))
    for k in ('cx', 'cy', 'ri', 'ro', 'n_bins', 'max_order', 'mask_count'):
        assert a.parameters[k] == meta['params'][k], k
    res = ctx.run(a)
    raw = res.raw_results
    assert raw.shape == g['raw_results'].shape and raw.dtype == np.complex64
    scale = np.abs(g['raw_results'][:, 0]).max()
    assert np.abs(raw - g['raw_results']).max() <= RTOL * scale
    assert res['absolute_0_0'].raw_data.shape == (6, 6)
    assert np.allclose(res['phase_3_2'].raw_data, np.angle(g['raw_results'][3, 2]), atol=1e-3)


@pytest.mark.parametrize('i', [0, 1, 2])
def test_com_params(lt, i):
    meta, g = load_golden(f'com_params_{i}')
    c = meta['com']
    if c['r'] == 'inf':
        c['r'] = float('inf')
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed'])
    ds = lt.MemoryDataSet(data=data, num_partitions=meta['num_partitions'], sig_dims=2)
    res = lt.run_udf(ds, lt.udf.CoMUDF.with_params(**c))
    check_com(res, g, prefix='', atol=5e-5)


def test_com_analysis_matches_udf(lt):
    # tests/udf/test_com.py:16-103: CoMUDF == COMAnalysis (field is (x, y) in the analysis)
    data = synth.dataset((6, 7, 16, 16), np.float32, 31)
    ds = lt.MemoryDataSet(data=data, num_partitions=2, sig_dims=2)
    ctx = lt.Context()
    a = ctx.create_com_analysis(ds, cx=8, cy=8, mask_radius=6, scan_rotation=12.)
    ares = ctx.run(a)
    ures = ctx.run_udf(ds, lt.udf.CoMUDF.with_params(cy=8, cx=8, r=6, scan_rotation=12.))
    fx, fy = ares['field'].raw_data
    np.testing.assert_allclose(ures['field'].data[..., 0], fy, atol=1e-6)
    np.testing.assert_allclose(ures['field'].data[..., 1], fx, atol=1e-6)
    np.testing.assert_allclose(ures['divergence'].data, ares['divergence'].raw_data, atol=1e-6)
    np.testing.assert_allclose(ures['curl'].data, ares['curl'].raw_data, atol=1e-6)
    np.testing.assert_allclose(ures['magnitude'].data, ares['magnitude'].raw_data, atol=1e-6)


def test_com_zero_frames_no_nan(lt):
    # tests/analysis/test_analysis_com.py:36-52
    data = np.zeros((4, 4, 8, 8), dtype=np.float32)
    ds = lt.MemoryDataSet(data=data, num_partitions=2, sig_dims=2)
    res = lt.run_udf(ds, lt.udf.CoMUDF())
    for k in ('raw_com', 'raw_shifts', 'field', 'magnitude', 'divergence', 'curl'):
        assert not np.any(np.isnan(res[k].data)), k
    assert np.allclose(res['raw_shifts'].data, 0)
    assert np.allclose(res['raw_com'].data, 4)


def test_com_known_answer(lt):
    # hand-crafted: single bright pixel per frame -> CoM is that pixel
    data = np.zeros((3, 3, 16, 16), dtype=np.float32)
    ys = np.arange(9).reshape(3, 3) + 2
    xs = 12 - np.arange(9).reshape(3, 3)
    for i in range(3):
        for j in range(3):
            data[i, j, ys[i, j], xs[i, j]] = 5.
    ds = lt.MemoryDataSet(data=data, num_partitions=3, sig_dims=2)
    res = lt.run_udf(ds, lt.udf.CoMUDF())
    assert np.array_equal(res['raw_com'].data[..., 0], ys.astype(np.float32))
    assert np.array_equal(res['raw_com'].data[..., 1], xs.astype(np.float32))
    assert np.array_equal(res['raw_shifts'].data[..., 0], (ys - 8).astype(np.float32))
    # linear field -> constant divergence, zero curl
    np.testing.assert_allclose(res['divergence'].data, 3 - 1, atol=1e-6)


def test_roi(lt):
    meta, g = load_golden('roi')
    shape = meta['shape']
    data = synth.dataset(shape, np.float32, meta['data_seed'])
    roi = roi_from_seed(shape[:2], meta['roi_seed'])
    stack = mixed_masks(shape[2], shape[3], meta['n_masks'], meta['mask_seed'])
    for src in (data, torch.from_numpy(data).cuda()):
        ds = lt.MemoryDataSet(data=src, num_partitions=meta['num_partitions'], sig_dims=2)
        res = lt.run_udf(ds, [lt.udf.ApplyMasksUDF(mask_factories=lambda: stack),
                              lt.udf.CoMUDF.with_params(regression=1), lt.udf.SumUDF(),
                              lt.udf.SumSigUDF()], roi=roi)
        close_cols(res[0]['intensity'].raw_data, g['intensity'])
        full = res[0]['intensity'].data
        assert full.shape == g['intensity_full'].shape
        assert np.array_equal(np.isnan(full), np.isnan(g['intensity_full']))
        check_com(res[1], g, atol=5e-5)
        np.testing.assert_allclose(res[1]['divergence'].data, g['com_divergence_full'],
                                   rtol=RTOL, atol=5e-5, equal_nan=True)
        np.testing.assert_allclose(res[2]['intensity'].data, g['sum'], rtol=RTOL)
        np.testing.assert_allclose(res[3]['intensity'].raw_data, g['sumsig'], rtol=RTOL)


@pytest.mark.parametrize('dt', ['float32', 'uint16'])
def test_odd_shapes(lt, dt):
    meta, g = load_golden('odd_' + dt)
    shape = meta['shape']
    data = synth.dataset(shape, np.dtype(dt), meta['data_seed'])
    stack = mixed_masks(shape[2], shape[3], meta['n_masks'], meta['mask_seed'])
    ds = lt.MemoryDataSet(data=data, num_partitions=meta['num_partitions'], sig_dims=2,
                          tileshape=(4, 17, 23))
    res = lt.run_udf(ds, [lt.udf.ApplyMasksUDF(mask_factories=lambda: stack), lt.udf.CoMUDF(),
                          lt.udf.SumUDF(), lt.udf.SumSigUDF()])
    close_cols(res[0]['intensity'].raw_data, g['intensity'])
    np.testing.assert_allclose(res[1]['raw_com'].raw_data, g['com_raw_com'], rtol=RTOL)
    np.testing.assert_allclose(res[2]['intensity'].data, g['sum'], rtol=RTOL)
    np.testing.assert_allclose(res[3]['intensity'].raw_data, g['sumsig'], rtol=RTOL)
    if dt == 'uint16':
        assert np.array_equal(res[2]['intensity'].data, g['sum'])
        assert np.array_equal(res[3]['intensity'].raw_data, g['sumsig'])


def test_subframe_tiles_accumulate(lt):
    meta, g = load_golden('subframe')
    shape = meta['shape']
    data = synth.dataset(shape, np.float32, meta['data_seed'])
    stack = mixed_masks(shape[2], shape[3], meta['n_masks'], meta['mask_seed'])
    ds = lt.MemoryDataSet(data=data, num_partitions=meta['num_partitions'], sig_dims=2,
                          tileshape=tuple(meta['tileshape']))
    res = lt.run_udf(ds, [lt.udf.ApplyMasksUDF(mask_factories=lambda: stack), lt.udf.SumUDF(),
                          lt.udf.SumSigUDF()])
    close_cols(res[0]['intensity'].raw_data, g['intensity'])
    np.testing.assert_allclose(res[1]['intensity'].data, g['sum'], rtol=RTOL)
    np.testing.assert_allclose(res[2]['intensity'].raw_data, g['sumsig'], rtol=RTOL)


def test_dtype_rules(lt):
    meta, g = load_golden('dtypes')
    shape = meta['shape']
    n = int(np.prod(shape))
    stack32 = mixed_masks(8, 8, 2, meta['mask_seed'])

    def run(data, stack, **kw):
        ds = lt.MemoryDataSet(data=data, num_partitions=2, sig_dims=2)
        return lt.run_udf(ds, lt.udf.ApplyMasksUDF(mask_factories=lambda: stack, **kw))

    d_i32 = (synth.hash_u32(0, n, meta['seeds']['i32']) % 1000).astype(np.int32).reshape(shape)
    r = run(d_i32, stack32)['intensity'].raw_data
    assert r.dtype == np.float64
    np.testing.assert_allclose(r, g['i32_f32'], rtol=1e-12)
    d_f32 = synth.dataset(shape, np.float32, meta['seeds']['f32'])
    stack64 = stack32.astype(np.float64) * 1.000000123
    r = run(d_f32, stack64)['intensity'].raw_data
    assert r.dtype == np.float64
    np.testing.assert_allclose(r, g['f32_f64'], rtol=1e-12)
    r = run(d_f32, stack64, mask_dtype=np.float32)['intensity'].raw_data
    assert r.dtype == np.float32
    close_cols(r, g['f32_f64_forced32'])
    stackc = (stack32 + 1j * stack32[::-1]).astype(np.complex64)
    r = run(d_f32, stackc)['intensity'].raw_data
    assert r.dtype == np.complex64
    np.testing.assert_allclose(r, g['f32_c64'], rtol=RTOL)
    d_u8 = (synth.hash_u32(0, n, meta['seeds']['u8']) % 256).astype(np.uint8).reshape(shape)
    ds = lt.MemoryDataSet(data=d_u8, num_partitions=2, sig_dims=2)
    res = lt.run_udf(ds, [lt.udf.ApplyMasksUDF(mask_factories=lambda: stack32), lt.udf.SumUDF(),
                          lt.udf.SumSigUDF()])
    close_cols(res[0]['intensity'].raw_data, g['u8_f32'])
    assert np.array_equal(res[1]['intensity'].data, g['u8_sum'])
    assert np.array_equal(res[2]['intensity'].raw_data, g['u8_sumsig'])


def test_shifts(lt):
    meta, g = load_golden('shifts')
    shape = meta['shape']
    data = synth.dataset(shape, np.float32, meta['data_seed'])
    stack = mixed_masks(shape[2], shape[3], meta['n_masks'], meta['mask_seed'])
    ds = lt.MemoryDataSet(data=data, num_partitions=2, sig_dims=2)
    r = lt.run_udf(ds, lt.udf.ApplyMasksUDF(mask_factories=lambda: stack,
                                            shifts=tuple(meta['const'])))
    close_cols(r['intensity'].raw_data, g['const'])
    sh = g['shifts']
    udf = lt.udf.ApplyMasksUDF(
        mask_factories=lambda: stack,
        shifts=lt.udf.ApplyMasksUDF.aux_data(sh.ravel(), kind='nav', extra_shape=(2,),
                                             dtype=sh.dtype))
    r = lt.run_udf(ds, udf)
    close_cols(r['intensity'].raw_data, g['perframe'])
    assert np.all(r['intensity'].raw_data[3] == 0)      # no overlap -> 0
    with pytest.raises(ValueError):
        lt.udf.ApplyMasksUDF(mask_factories=lambda: stack, shifts=(1, 1),
                             use_sparse='scipy.sparse')
    # the whole tile goes through ONE launch of the shifted-mask kernel, not a frame loop
    from libertem_b200 import engine
    engine.launch_count(reset=True)
    lt.run_udf(ds, udf)
    assert engine.launch_count() == 2          # one per partition
    # float64 masks (float64 result, the reference's dtype rule) and complex masks go through the
    # same kernel -- one launch per partition, no frame loop -- with the same numbers
    stack64 = stack.astype(np.float64)
    engine.launch_count(reset=True)
    r64 = lt.run_udf(ds, lt.udf.ApplyMasksUDF(mask_factories=lambda: stack64,
                                              shifts=tuple(meta['const'])))
    assert engine.launch_count() == 2
    assert r64['intensity'].raw_data.dtype == np.float64
    close_cols(r64['intensity'].raw_data, g['const'], rtol=1e-6)
    stackc = (stack[:1] + 1j * stack[1:2]).astype(np.complex64)
    udfc = lt.udf.ApplyMasksUDF(
        mask_factories=lambda: stackc,
        shifts=lt.udf.ApplyMasksUDF.aux_data(sh.ravel(), kind='nav', extra_shape=(2,),
                                             dtype=sh.dtype))
    engine.launch_count(reset=True)
    rc = lt.run_udf(ds, udfc)['intensity'].raw_data
    assert engine.launch_count() == 2 and rc.dtype == np.complex64
    close_cols(rc.real, g['perframe'][:, :1])
    close_cols(rc.imag, g['perframe'][:, 1:2])


def test_process_tile_seam_with_numpy_tile(lt):
    """drop-in seam: UDF.process_tile(tile) with a host numpy tile, as the reference's
    BACKEND_CUDA runtime would call it"""
    from libertem_b200.udf.base import UDFMeta, UDFData
    from libertem_b200.common import Shape
    data = synth.dataset((5, 4, 32, 32), np.float32, 41)
    mask = synth.uniform_f32(0, 1024, 42).reshape(32, 32)
    udf = lt.udf.ApplyMasksUDF(mask_factories=[lambda: mask])
    dsh = Shape(data.shape, sig_dims=2)
    udf.set_meta(UDFMeta(dataset_shape=dsh, dataset_dtype=np.float32, input_dtype=np.float32,
                         device=torch.device('cuda')))
    decl = udf.get_result_buffers()
    for b in decl.values():
        b.set_shape_partition(dsh, 20)
        b.allocate(torch.device('cuda'))
    udf.results = UDFData(decl)
    udf.task_data = udf.get_task_data()
    udf.results.set_view('intensity', decl['intensity'].rows(0, 20))
    udf.process_tile(data.reshape(20, 32, 32))
    udf.process_tile(data.reshape(20, 32, 32))         # += semantics
    ref = O.apply_masks(data, mask[None])
    close_cols(decl['intensity'].raw_data, 2 * ref)


def test_synthetic_dataset_matches_oracle(lt):
    ds = lt.SyntheticDataSet((8, 8, 64, 64), np.float32, seed=55, num_partitions=4)
    mask = synth.uniform_f32(0, 4096, 56).reshape(64, 64)
    res = lt.run_udf(ds, [lt.udf.ApplyMasksUDF(mask_factories=[lambda: mask]), lt.udf.CoMUDF()])
    data = synth.dataset((8, 8, 64, 64), np.float32, 55)
    close_cols(res[0]['intensity'].raw_data, O.apply_masks(data, mask[None], num_partitions=4))
    ref = O.com_udf(data, num_partitions=4)
    np.testing.assert_allclose(res[1]['raw_com'].raw_data, ref['raw_com'], rtol=RTOL)


def _corr_inputs(meta):
    dark = synth.uniform_f32(0, 192, meta['dark_seed']).reshape(16, 12) * 0.3
    gain = 0.5 + synth.uniform_f32(0, 192, meta['gain_seed']).reshape(16, 12)
    excl = np.zeros((16, 12), dtype=bool)
    for y, x in meta['excluded']:
        excl[y, x] = True
    return dark, gain, excl


@pytest.mark.parametrize('dt', ['float32', 'uint16'])
@pytest.mark.parametrize('name', ['dg', 'dge', 'e'])
def test_corrections_folded_into_masks(lt, dt, name):
    """dark / gain / dead-pixel corrections (io/corrections/corrset.py) folded into the masks:
    same results as the reference's per-tile correction, frames still read once"""
    from libertem_b200.corrections import CorrectionSet
    meta, g = load_golden('corrections')
    dark, gain, excl = _corr_inputs(meta)
    kw = {'dg': dict(dark=dark, gain=gain), 'dge': dict(dark=dark, gain=gain, excluded_pixels=excl),
          'e': dict(excluded_pixels=excl)}[name]
    data = synth.dataset(meta['shape'], np.dtype(dt), meta['seeds'][dt])
    stack = mixed_masks(16, 12, 3, meta['mask_seed'])
    ds = lt.MemoryDataSet(data=data.copy(), num_partitions=2, sig_dims=2)
    runner = lt.UDFRunner([lt.udf.ApplyMasksUDF(mask_factories=lambda: stack), lt.udf.CoMUDF(),
                           lt.udf.SumUDF(), lt.udf.SumSigUDF()])
    res = runner.run_for_dataset(ds, corrections=CorrectionSet(**kw)).buffers
    assert runner.stats['unfused_calls'] == 0          # fused: no corrected copy of the tile
    key = f'{dt}_{name}_'
    close_cols(res[0]['intensity'].raw_data, g[key + 'intensity'])
    np.testing.assert_allclose(res[1]['raw_com'].raw_data, g[key + 'raw_com'], rtol=RTOL)
    np.testing.assert_allclose(res[2]['intensity'].data, g[key + 'sum'], rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(res[3]['intensity'].raw_data, g[key + 'sumsig'], rtol=RTOL)
    assert np.array_equal(ds.data, data)               # the input is never modified


def test_corrections_explicit_path_subframe_tiles(lt):
    from libertem_b200.corrections import CorrectionSet, RepairValueError
    from oracle import corrections as OC
    meta, g = load_golden('corrections')
    dark, gain, excl = _corr_inputs(meta)
    data = synth.dataset(meta['shape'], np.float32, meta['seeds']['float32'])
    stack = mixed_masks(16, 12, 3, meta['mask_seed'])
    ds = lt.MemoryDataSet(data=data, num_partitions=2, sig_dims=2, tileshape=(4, 8, 12))
    res = lt.run_udf(ds, [lt.udf.ApplyMasksUDF(mask_factories=lambda: stack), lt.udf.SumUDF()],
                     corrections=CorrectionSet(dark=dark, gain=gain))
    c = OC.correct(data, dark=dark, gain=gain)
    close_cols(res[0]['intensity'].raw_data, O.apply_masks(c, stack))
    np.testing.assert_allclose(res[1]['intensity'].data, c.reshape(20, 16, 12).sum(0), rtol=RTOL)
    bad = np.ones((16, 12), dtype=bool)
    with pytest.raises(RepairValueError):
        CorrectionSet(excluded_pixels=bad)


def test_user_defined_udfs_through_the_plugin_api(lt):
    """the UDF protocol is generic: user UDFs with process_tile / process_frame /
    process_partition run next to the fused hot-path UDFs (tiles are CUDA tensors)"""
    UDF = lt.udf.UDF

    class MaxTileUDF(UDF):
        def get_result_buffers(self):
            return {'maxval': self.buffer(kind='nav', dtype=np.float32, where='device')}

        def process_tile(self, tile):
            self.results.maxval[:] = tile.reshape(tile.shape[0], -1).amax(dim=1)

    class FrameStdUDF(UDF):
        def get_result_buffers(self):
            return {'std': self.buffer(kind='nav', dtype=np.float32, where='device')}

        def process_frame(self, frame):
            self.results.std[...] = frame.std(unbiased=False)

    class PartSumUDF(UDF):
        def get_result_buffers(self):
            return {'total': self.buffer(kind='single', dtype=np.float64, where='device')}

        def process_partition(self, partition):
            self.results.total[:] += partition.double().sum()

        def merge(self, dest, src):
            dest.total[:] += src.total

    data = synth.dataset((6, 5, 16, 16), np.float32, 71)
    ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(), num_partitions=3, sig_dims=2)
    res = lt.run_udf(ds, [MaxTileUDF(), FrameStdUDF(), PartSumUDF(), lt.udf.SumSigUDF()])
    np.testing.assert_allclose(res[0]['maxval'].data, data.max(axis=(2, 3)))
    np.testing.assert_allclose(res[1]['std'].data, data.std(axis=(2, 3)), rtol=1e-5)
    np.testing.assert_allclose(res[2]['total'].data[0], data.astype(np.float64).sum(), rtol=1e-10)
    np.testing.assert_allclose(res[3]['intensity'].data, data.sum(axis=(2, 3)), rtol=RTOL)
    assert MaxTileUDF().get_method() == 'tile' and FrameStdUDF().get_method() == 'frame'
    assert PartSumUDF().get_method() == 'partition'


def test_host_streaming_many_partitions_no_buffer_reuse_race(lt):
    """host data, several partitions, several depth blocks per partition --
    the staging buffers persist on the dataset and every copy waits for the last reader, so
    the last tile of a partition cannot be overwritten by the next partition's first copy.
    Run repeatedly (the race was timing dependent) on a stream kept busy."""
    shape = (16, 24, 64, 64)
    data = synth.dataset(shape, np.float32, 91)
    masks = mixed_masks(64, 64, 4, 5)
    ref = O.apply_masks(data, masks, num_partitions=6)
    ref_sum = data.reshape(-1, 64 * 64).sum(axis=0, dtype=np.float64)
    ds = lt.MemoryDataSet(data=data, num_partitions=6, sig_dims=2, tile_depth=11)
    busy = torch.empty(1 << 26, device='cuda')
    for _ in range(5):
        for _ in range(4):
            busy.normal_()          # keep the main stream well ahead of the host
        res = lt.run_udf(ds, [lt.udf.ApplyMasksUDF(mask_factories=lambda: masks),
                              lt.udf.SumUDF()])
        close_cols(res[0]['intensity'].raw_data, ref)
        np.testing.assert_allclose(res[1]['intensity'].raw_data.reshape(-1), ref_sum, rtol=1e-5)
    ds.release()


def test_merge_all_matches_merge(lt):
    """use_merge_all=True: the reference's merge_all contract (udf/base.py:944-1002,
    udf/sum.py:54-58) gives the same buffers as the sequential merge"""
    shape = (6, 8, 32, 32)
    data = synth.dataset(shape, np.float32, 17)
    masks = mixed_masks(32, 32, 3, 2)
    src = torch.from_numpy(data).cuda()
    out = []
    for use in (False, True):
        ds = lt.MemoryDataSet(data=src, num_partitions=4, sig_dims=2)
        udfs = [lt.udf.ApplyMasksUDF(mask_factories=lambda: masks), lt.udf.SumUDF(),
                lt.udf.SumSigUDF()]
        # fuse=False: partition buffers are real copies, so merge / merge_all do the work
        r = lt.UDFRunner(udfs, fuse=False).run_for_dataset(ds, use_merge_all=use).buffers
        out.append(r)
    for a, b in zip(*out):
        for k in a:
            np.testing.assert_array_equal(a[k].raw_data, b[k].raw_data)
    close_cols(out[1][0]['intensity'].raw_data, O.apply_masks(data, masks, num_partitions=4))
    np.testing.assert_allclose(out[1][1]['intensity'].raw_data.reshape(-1),
                               data.reshape(-1, 1024).sum(axis=0), rtol=1e-5)


def test_user_udf_gets_input_dtype_tiles_and_frame_views(lt):
    """a generic UDF on a uint16 dataset receives float32 tiles (tile.dtype ==
    meta.input_dtype, reference udf/base.py:106-123), and process_frame's nav view of a buffer
    without extra_shape has shape (1,), so ``self.results.x[:] = value`` works
    (common/buffers.py:761-790)"""
    UDF = lt.udf.UDF
    seen = []

    class TileMaxUDF(UDF):
        def get_result_buffers(self):
            return {'m': self.buffer(kind='nav', dtype=np.float32, where='device')}

        def process_tile(self, tile):
            seen.append(tile.dtype)
            self.results.m[:] = tile.reshape(tile.shape[0], -1).amax(dim=1)

    class FrameMeanUDF(UDF):
        def get_result_buffers(self):
            return {'mean': self.buffer(kind='nav', dtype=np.float32, where='device'),
                    'mm': self.buffer(kind='nav', extra_shape=(2,), dtype=np.float32,
                                      where='device')}

        def process_frame(self, frame):
            seen.append(frame.dtype)
            assert self.results.mean.shape == (1,)
            assert self.results.mm.shape == (2,)
            self.results.mean[:] = frame.mean()
            self.results.mm[:] = torch.stack([frame.amin(), frame.amax()])

    data = synth.dataset((4, 5, 16, 16), np.uint16, 5)
    ds = lt.MemoryDataSet(data=torch.from_numpy(data.view(np.int16)).view(torch.uint16).cuda(),
                          num_partitions=2, sig_dims=2)
    res = lt.run_udf(ds, [TileMaxUDF(), FrameMeanUDF(), lt.udf.SumSigUDF()])
    assert seen and all(d == torch.float32 for d in seen)
    np.testing.assert_array_equal(res[0]['m'].data, data.max(axis=(2, 3)).astype(np.float32))
    np.testing.assert_allclose(res[1]['mean'].data, data.mean(axis=(2, 3)), rtol=1e-6)
    np.testing.assert_array_equal(res[1]['mm'].data[..., 1], data.max(axis=(2, 3)))
    np.testing.assert_array_equal(res[2]['intensity'].data, data.sum(axis=(2, 3)))


def test_corrections_do_not_leak_between_runs(lt):
    """a UDFRunner reused with a different CorrectionSet on a sub-frame
    tiling must not apply the previous run's dark / gain"""
    from libertem_b200.corrections import CorrectionSet
    data = synth.dataset((4, 4, 16, 16), np.float32, 9)
    masks = mixed_masks(16, 16, 2, 1)
    ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(), num_partitions=2, sig_dims=2,
                          tileshape=(4, 8, 16))
    runner = lt.UDFRunner([lt.udf.ApplyMasksUDF(mask_factories=lambda: masks)])
    outs = []
    for dark in (np.full((16, 16), 0.25, np.float32), np.full((16, 16), -0.5, np.float32)):
        cs = CorrectionSet(dark=dark)
        got = runner.run_for_dataset(ds, corrections=cs).buffers[0]['intensity'].raw_data
        want = O.apply_masks(data - dark, masks, num_partitions=2)
        close_cols(got, want)
        outs.append(got)
    assert not np.allclose(outs[0], outs[1])


@pytest.mark.parametrize('name', ['centre', 'offcentre'])
@pytest.mark.parametrize('kernel', ['auto', 'walk', 'banded', 'ffma'])
def test_cfg4_k7_radial_fourier_golden(lt, name, kernel, monkeypatch):
    """RadialFourierAnalysis vs the unmodified reference at a size where the DEFAULT kernels
    (tcgen05 group-sparse, >= 96 frames per tile) run: 256 frames, one partition
    (reference analysis/radialfourier.py:106-146,184-194).  'auto' must pick K10 (dense walk)
    when the stack has such a plan, else K7; 'walk' / 'banded' force K10 / K7."""
    from libertem_b200 import engine, group_masks as gm
    meta, g = load_golden('cfg4_k7_' + name)
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed'])
    ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(), num_partitions=1, sig_dims=2)
    ctx = lt.Context()
    a = ctx.create_radial_fourier_analysis(ds, **meta['call'])
    for k in ('cx', 'cy', 'ri', 'ro', 'n_bins', 'max_order', 'mask_count'):
        assert a.parameters[k] == pytest.approx(meta['params'][k]), k
    if kernel != 'auto':
        orig = gm.group_masks
        monkeypatch.setattr(gm, 'group_masks',
                            lambda *args, **kw: orig(*args, **{**kw, 'kernel': kernel}))
    res = ctx.run(a)
    want_kernel = {'auto': (10, 70, 71), 'walk': (10,), 'banded': (70,), 'ffma': (4,)}[kernel]
    assert engine.last_kernel() in want_kernel, engine.last_kernel()
    raw = res.raw_results
    ref = g['raw_results']
    assert raw.shape == ref.shape and raw.dtype == np.complex64
    # per (bin, order) scale, like close_cols: every output column within 1e-5 of its own
    # magnitude scale (the o = 0 column of its bin bounds |sum| of every order of that bin)
    scale = np.abs(ref[:, :1]).max(axis=(2, 3), keepdims=True)
    err = (np.abs(raw - ref) / scale).max()
    assert err <= RTOL, err


@pytest.mark.parametrize('kernel', ['auto', 'ffma'])
def test_radial_fourier_symmetries_known_answer(lt, kernel, monkeypatch):
    """the reference's known-answer test tests/analysis/test_analysis_radialfourier.py:78-188
    (CBED frames with 1- / 2- / 4-fold spot symmetry), tiled to 256 frames so that it runs
    through K7; inputs from the reference's cbed_frame generator (golden fixture), checked
    against the reference's outputs AND against the analytic statements of that test."""
    from libertem_b200 import engine, group_masks as gm
    meta, g = load_golden('radial_symmetries')
    frames = g['frames']
    data = np.zeros((16, 16) + frames.shape[1:], dtype=np.float32)
    for i in range(16):
        for j in range(16):
            data[i, j] = frames[(i % 2) * 2 + (j % 2)]
    ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(), num_partitions=1, sig_dims=2)
    ctx = lt.Context()
    a = ctx.create_radial_fourier_analysis(ds, **meta['call'])
    if kernel != 'auto':
        orig = gm.group_masks
        monkeypatch.setattr(gm, 'group_masks',
                            lambda *args, **kw: orig(*args, **{**kw, 'kernel': kernel}))
    res = ctx.run(a)
    assert engine.last_kernel() in ((10, 70, 71, 7) if kernel == 'auto' else (4,))
    raw = res.raw_results
    ref = g['raw_results']
    tot = float(g['frame_sums'].max())
    assert np.abs(raw - ref).max() <= RTOL * tot
    sums = data.sum(axis=(2, 3), dtype=np.float64)
    # bin 0 holds (numerically) nothing, so its channels are not normalised (normal = max(1, .));
    # bin 1's higher orders are divided by |c_1_0| = the frame sum
    tol = dict(atol=2e-5 * tot, rtol=1e-5)
    tol1 = dict(atol=2e-5, rtol=1e-5)
    c = lambda b, o: res[f'complex_{b}_{o}'].raw_data[:2, :2]      # the original 2x2 scan
    np.testing.assert_allclose(np.abs(c(0, 0)), 0, **tol)
    np.testing.assert_allclose(np.abs(c(1, 0)), sums[:2, :2], rtol=1e-5)
    for o in (1, 2, 3, 4, 5):
        np.testing.assert_allclose(np.abs(c(0, o)), 0, **tol)
    # odd harmonics suppressed in 2-fold symmetry, 2-fold suppressed in 4-fold symmetry
    for o in (1, 3, 5, 7):
        np.testing.assert_allclose(np.abs(c(1, o)[1]), 0, **tol1)
    np.testing.assert_allclose(np.abs(c(1, 2)[1, 1]), 0, **tol1)
    assert np.all(np.abs(c(1, 1)[0]) > 0.1)
    np.testing.assert_allclose(np.angle(c(1, 1)[0, 0]), np.pi / 2, atol=1e-4)
    np.testing.assert_allclose(np.angle(c(1, 1)[0, 1]), -np.pi / 2, atol=1e-4)
    np.testing.assert_allclose(np.abs(np.angle(c(1, 2)[0])), np.pi, atol=1e-4)
    np.testing.assert_allclose(np.angle(c(1, 3)[0, 0]), -np.pi / 2, atol=1e-4)
    np.testing.assert_allclose(np.angle(c(1, 3)[0, 1]), np.pi / 2, atol=1e-4)
    np.testing.assert_allclose(np.angle(c(1, 4)), 0, atol=1e-4)
    np.testing.assert_allclose(np.angle(c(1, 8)), 0, atol=1e-4)


def test_udf_level_k6_tensor_path(lt):
    """the runner -> result slab -> K6 (tcgen05 dense kernel) combination that produces the
    headline number: one partition of 2048 frames, 8 masks + CoM + SumSig fused in one pass,
    against the oracle (the bench path was only covered piecewise)"""
    from libertem_b200 import engine
    shape = (32, 64, 64, 64)
    data = synth.dataset(shape, np.float32, 33)
    masks = mixed_masks(64, 64, 8, 4)
    ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(), num_partitions=1, sig_dims=2)
    runner = lt.UDFRunner([lt.udf.ApplyMasksUDF(mask_factories=lambda: masks), lt.udf.CoMUDF(),
                           lt.udf.SumSigUDF()])
    res = runner.run_for_dataset(ds).buffers
    assert engine.last_kernel() == 6, engine.last_kernel()
    assert runner.stats['unfused_calls'] == 0 and runner.stats['fused_launch_groups'] == 1
    close_cols(res[0]['intensity'].raw_data, O.apply_masks(data, masks, num_partitions=1))
    com = O.com_udf(data, num_partitions=1)
    close_cols(res[1]['raw_com'].raw_data, com['raw_com'])
    np.testing.assert_allclose(res[1]['field'].raw_data, com['field'], rtol=RTOL, atol=2e-4)
    np.testing.assert_allclose(res[2]['intensity'].raw_data, O.sumsig_udf(data), rtol=RTOL)
    # and the same through 4 partitions of 512 frames (FFMA2 kernel): same results within tol
    ds4 = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(), num_partitions=4, sig_dims=2)
    res4 = lt.run_udf(ds4, [lt.udf.ApplyMasksUDF(mask_factories=lambda: masks)])
    close_cols(res4[0]['intensity'].raw_data, res[0]['intensity'].raw_data)


def test_udf_level_k8_large_signal_and_wide_stack(lt):
    """uint16 frames x integer masks beyond the round-1 routing limits go to the int8
    tensor-core kernel: a 288x256 detector (73728 px > 65536: K-split with int64
    recombination) with CoM (coordinate masks up to 287 -> base-128 digit rows) + SumUDF +
    SumSigUDF + 20 binary masks (> 16 int8 rows per pass) -- bit-exact"""
    from libertem_b200 import engine
    sy, sx = 288, 256
    shape = (16, 20, sy, sx)
    data = synth.dataset(shape, np.uint16, 44)
    yy, xx = np.mgrid[:sy, :sx]
    masks = np.stack([((yy // 12 + xx // 16 + i) % 3 == 0) for i in range(20)]).astype(np.float32)
    ds = lt.MemoryDataSet(
        data=torch.from_numpy(data.view(np.int16)).view(torch.uint16).cuda(),
        num_partitions=1, sig_dims=2)
    runner = lt.UDFRunner([lt.udf.SumUDF(), lt.udf.SumSigUDF(), lt.udf.CoMUDF(),
                           lt.udf.ApplyMasksUDF(mask_factories=lambda: masks)])
    res = runner.run_for_dataset(ds).buffers
    assert engine.last_kernel() == 8 and runner.stats.get('int8_passes', 0) >= 1
    assert runner.stats['unfused_calls'] == 0
    f32 = data.astype(np.float32)
    flat64 = data.reshape(-1, sy * sx).astype(np.float64)
    assert np.array_equal(res[0]['intensity'].raw_data.reshape(-1),
                          flat64.sum(axis=0).astype(np.float32))
    assert np.array_equal(res[1]['intensity'].raw_data, flat64.sum(axis=1).astype(np.float32))
    exact = (flat64 @ masks.reshape(20, -1).T.astype(np.float64)).astype(np.float32)
    assert np.array_equal(res[3]['intensity'].raw_data, exact)
    com = O.com_udf(f32, num_partitions=1)
    close_cols(res[2]['raw_com'].raw_data, com['raw_com'])


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


@pytest.mark.parametrize('where', ['device', 'host'])
def test_complex_input_golden(lt, where):
    """complex64 frames vs the unmodified reference (udf/masks.py:360-368: the result dtype is
    result_type(input, mask) = complex64): real masks, complex masks, full-frame and sub-frame
    tiles.  The tile is read once as its interleaved (re, im) float view."""
    meta, g = load_golden('complex_input')
    data, real_masks, cmasks = _complex_inputs(meta)
    src = torch.from_numpy(data).cuda() if where == 'device' else data
    for name, kw in (('p2', dict(num_partitions=2)),
                     ('tiled', dict(num_partitions=3, tileshape=(5, 8, 32)))):
        ds = lt.MemoryDataSet(data=src, sig_dims=2, **kw)
        for key, masks, dt in (('real_masks_', real_masks, np.float32),
                               ('complex_masks_', cmasks, np.complex64)):
            res = lt.run_udf(ds, lt.udf.ApplyMasksUDF(mask_factories=lambda: masks,
                                                      mask_dtype=dt, use_sparse=False))
            got = res['intensity'].raw_data
            ref = g[key + name]
            assert got.dtype == np.complex64 and got.shape == ref.shape
            scale = np.abs(ref).max(axis=0)
            assert (np.abs(got - ref) / scale).max() <= RTOL, key + name


def test_com_analysis_complex_known_answers(lt):
    """tests/analysis/test_analysis_com.py:132-231: COMAnalysis on complex frames returns
    x_real / y_real / x_imag / y_imag; handcrafted frames with exact answers, and the
    reference's outputs on random complex data"""
    ctx = lt.Context()
    cases = [
        ([[0, 0, 0, 0], [0, 1 + 2j, 1 - 2j, 0], [0, 1 - 2j, 1 + 2j, 0], [0, 0, 0, 0]], 1.5, 1.5),
        ([[0, 0, 0, 0], [0, 0, 1 - 2j, 0], [0, 1 - 2j, 0, 0], [0, 0, 0, 0]], 1.5, 1.5),
        ([[0, 0, 0, 0], [0, 0, 1 - 2j, 0], [0, 0, 0, 0], [0, 0, 0, 0]], 2, 1),
    ]
    for frame, want_x, want_y in cases:
        data = np.ones((3, 3, 4, 4), dtype=np.complex64)
        data[0, 0] = np.array(frame, dtype=np.complex64)
        ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(), num_partitions=9, sig_dims=2)
        res = ctx.run(ctx.create_com_analysis(dataset=ds, cx=0, cy=0, mask_radius=None))
        fx = res['x_real'].raw_data + 1j * res['x_imag'].raw_data
        fy = res['y_real'].raw_data + 1j * res['y_imag'].raw_data
        assert fx[0, 0] == want_x and fy[0, 0] == want_y
    meta, g = load_golden('complex_input')
    data, _, _ = _complex_inputs(meta)
    ds = lt.MemoryDataSet(data=torch.from_numpy(data).cuda(), num_partitions=2, sig_dims=2)
    res = ctx.run(ctx.create_com_analysis(dataset=ds, cx=0, cy=0, mask_radius=None))
    fx = res['x_real'].raw_data + 1j * res['x_imag'].raw_data
    fy = res['y_real'].raw_data + 1j * res['y_imag'].raw_data
    np.testing.assert_allclose(fx, g['com_x'], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(fy, g['com_y'], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('reg', [-1, 0, 1, 'given'])
@pytest.mark.parametrize('use_roi', [False, True])
def test_com_postprocess_device_matches_host(lt, reg, use_roi):
    """CoMUDF.get_results through the nav-space kernels (K9) vs the numpy pipeline on the same
    moments: all nine result buffers, every regression mode, with and without a roi"""
    from libertem_b200.udf.com import CoMUDF
    from libertem_b200.udf.base import UDFMeta, UDFData
    from libertem_b200.common import Shape
    nav, sig = (9, 13), (16, 16)
    ds_shape = Shape(nav + sig, sig_dims=2)
    n = nav[0] * nav[1]
    roi = roi_from_seed(nav, 5) if use_roi else None
    n_rows = n if roi is None else int(roi.sum())
    rng = np.random.default_rng(3)
    raw = np.empty((n_rows, 3), dtype=np.float32)
    raw[:, 0] = rng.uniform(50, 100, n_rows)
    raw[:, 1] = raw[:, 0] * rng.uniform(6, 10, n_rows)
    raw[:, 2] = raw[:, 0] * rng.uniform(5, 9, n_rows)
    raw[3] = 0                                    # empty frame: shifts fall back to (cy, cx)
    regression = np.array([[0.3, -0.2], [0.01, 0.02], [-0.03, 0.005]]) if reg == 'given' else reg
    out = []
    for device in ('cuda', 'cpu'):
        udf = CoMUDF.with_params(cy=7.3, cx=8.1, scan_rotation=27., flip_y=True,
                                 regression=regression)
        udf.set_meta(UDFMeta(dataset_shape=ds_shape, roi=roi, dataset_dtype=np.float32,
                             input_dtype=np.float32, device=torch.device('cuda')))
        decl = udf.get_result_buffers()
        for b in decl.values():
            b.set_shape_ds(ds_shape, roi)
        decl['raw_mask_result'].replace_array(torch.from_numpy(raw).to(device))
        udf.results = UDFData(decl)
        out.append(udf.get_results())
    dev, host = out
    assert set(dev) == set(host)
    for k in host:
        a, b = np.asarray(dev[k]), np.asarray(host[k], dtype=np.float64)
        assert a.reshape(-1).shape == b.reshape(-1).shape, k
        np.testing.assert_allclose(a.reshape(b.shape), b, rtol=2e-6, atol=2e-6, err_msg=k,
                                   equal_nan=True)


def test_guess_corrections_device(lt):
    """libertem_b200.nav.guess_corrections (one Gram-matrix pass + closed-form sweep) returns
    what the reference's 720-pass search returns (udf/com.py:207-295)"""
    from libertem_b200 import nav
    from libertem_b200.udf.com import guess_corrections, apply_correction
    ny, nx = 24, 31
    yy, xx = np.mgrid[:ny, :nx].astype(np.float64)
    # an anisotropic, mostly divergent field seen through a rotated, flipped scan + an offset
    fy, fx = (yy - 11.5) * 0.3 + 0.05 * np.sin(xx / 3), (xx - 15.0) * 0.12
    rng = np.random.default_rng(1)
    fy += rng.normal(0, 0.01, fy.shape)
    fx += rng.normal(0, 0.01, fx.shape)
    for rot, flip in ((33, False), (-71, True), (170, False)):
        y, x = apply_correction(fy, fx, scan_rotation=rot, flip_y=flip, forward=False)
        y = (y + 2.5).astype(np.float32)
        x = (x - 1.25).astype(np.float32)
        want = guess_corrections(y, x)
        got = nav.guess_corrections(y, x)
        assert got.scan_rotation == want.scan_rotation and got.flip_y == want.flip_y
        assert got.cy == pytest.approx(float(want.cy), rel=1e-5)
        assert got.cx == pytest.approx(float(want.cx), rel=1e-5)
        sub = (slice(2, 20), slice(3, 25))
        assert nav.guess_corrections(y, x, roi=sub)[:2] == guess_corrections(y, x, roi=sub)[:2]


def test_u16_float_masks_fixed_point_int8_path(lt):
    """uint16 frames x NON-integer masks: the masks become 28-bit fixed-point int8 digit rows
    (runner.int8_digit_plan) and the pass runs on the integer tensor cores (K8) with exact
    integer sums -- the result is closer to the float64 sums than the reference's float32 GEMM
    and within the north-star tolerance of it; SumUDF / SumSigUDF stay bit-exact"""
    from libertem_b200 import engine
    shape = (24, 32, 64, 64)
    data = synth.dataset(shape, np.uint16, 77)
    yy, xx = np.mgrid[:64, :64]
    r = np.hypot(yy - 31.5, xx - 30.2)
    masks = np.stack([np.exp(-((r - r0) / 4.0) ** 2) * w
                      for r0, w in ((8, 1.0), (16, 0.37), (24, 2.5e-3), (28, -41.7))]
                     ).astype(np.float32)
    ds = lt.MemoryDataSet(
        data=torch.from_numpy(data.view(np.int16)).view(torch.uint16).cuda(),
        num_partitions=1, sig_dims=2)
    runner = lt.UDFRunner([lt.udf.SumUDF(), lt.udf.SumSigUDF(),
                           lt.udf.ApplyMasksUDF(mask_factories=lambda: masks)])
    res = runner.run_for_dataset(ds).buffers
    assert engine.last_kernel() == 8 and runner.stats.get('int8_passes', 0) == 1
    flat64 = data.reshape(-1, 4096).astype(np.float64)
    assert np.array_equal(res[0]['intensity'].raw_data.reshape(-1),
                          flat64.sum(axis=0).astype(np.float32))
    assert np.array_equal(res[1]['intensity'].raw_data, flat64.sum(axis=1).astype(np.float32))
    exact = flat64 @ masks.reshape(4, -1).T.astype(np.float64)
    got = res[2]['intensity'].raw_data
    scale = np.abs(exact).max(axis=0)
    assert (np.abs(got - exact) / scale).max() <= 2e-7          # one float32 rounding
    close_cols(got, O.apply_masks(data, masks, num_partitions=1))
