"""GPU parity tests of the C-ABI kernels (K1 dense, K2 sparse, synth) against the CPU oracle
and the committed golden vectors.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from golden_inputs import mixed_masks, ring_stack
from oracle import synth, udf_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: <= 1e-5 rel for float32 mask/CoM results


@pytest.fixture(scope='module', params=['auto', 'eo', 'pair', 'tc'])
def eng(request):
    """every kernel-level test runs against both FFMA2 register tiles of the TMA kernel and
    against the tcgen05 tensor-core kernel (K6) forced on for every shape it takes"""
    from libertem_b200 import engine
    assert torch.cuda.is_available()
    engine.set_k1_variant({'auto': 0, 'eo': 1, 'pair': 2, 'tc': 3}[request.param])
    yield engine
    engine.set_k1_variant(0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def f64_truth(tile, masks):
    return tile.astype(np.float64) @ masks.astype(np.float64).T


def abs_scale(tile, masks):
    """sum_k |tile||mask| per column (max over frames): the scale fp32 rounding errors live on
    (signed masks cancel, so |result| alone under-estimates it)."""
    return (np.abs(tile).astype(np.float64) @ np.abs(masks).astype(np.float64).T).max(
        axis=0, keepdims=True)


def assert_close_rel(res, truth, rtol=RTOL, scale=None):
    if scale is None:
        scale = np.abs(truth).max(axis=0, keepdims=True)
    err = np.abs(res - truth) / (scale + 1e-30)
    assert err.max() <= rtol, f'max rel err {err.max():.3e}'


def test_synth_bit_exact(eng):
    for start in (0, 12345, (1 << 32) - 100, (1 << 33) + 7):
        n = 100003
        a = eng.synth_fill((n,), np.float32, 17, 'cuda', start=start).cpu().numpy()
        assert np.array_equal(a, synth.uniform_f32(start, n, 17))
        b = eng.synth_fill((n,), np.uint16, 99, 'cuda', start=start).cpu().numpy()
        assert np.array_equal(b, synth.poisson3_u16(start, n, 99))


@pytest.mark.parametrize('nparts', [1, 8])
def test_cfg1_golden(eng, nparts):
    meta, g = load_golden(f'cfg1_p{nparts}')
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed']).reshape(1024, 4096)
    mask = synth.uniform_f32(0, 4096, meta['mask_seed']).reshape(1, 4096)
    out = eng.masks_dense(dev(data), dev(mask)).cpu().numpy()
    assert eng.last_kernel() in (1, 3, 6)
    np.testing.assert_allclose(out, g['intensity'], rtol=RTOL)
    assert_close_rel(out, f64_truth(data, mask), 2e-6)


def test_cfg2_small_golden(eng):
    meta, g = load_golden('cfg2_small')
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed']).reshape(256, 65536)
    stack = mixed_masks(256, 256, 8, meta['mask_seed']).reshape(8, -1)
    com = O.com_mask_stack((256, 256), 128, 128).reshape(3, -1)
    ones = np.ones((1, 65536), dtype=np.float32)
    allm = np.concatenate([stack, com, ones])          # 12 columns in ONE pass
    sig_sum = torch.zeros(65536, dtype=torch.float32, device='cuda')
    out = eng.masks_dense(dev(data), dev(allm), sig_sum=sig_sum).cpu().numpy()
    assert eng.last_kernel() in (1, 3, 6)
    truth = f64_truth(data, allm)
    assert_close_rel(out, truth, 2e-6, abs_scale(data, allm))
    assert_close_rel(out[:, :8], g['intensity'])
    assert_close_rel(out[:, 8:11], g['com_raw_mask_result'])
    np.testing.assert_allclose(out[:, 11], g['sumsig'], rtol=RTOL)
    np.testing.assert_allclose(sig_sum.cpu().numpy().reshape(256, 256), g['sum'], rtol=RTOL)


@pytest.mark.parametrize('n_masks', [1, 2, 3, 5, 7, 8, 11, 12, 13, 16, 19, 24, 25, 40])
def test_dense_tma_mask_counts(eng, n_masks):
    F, K = 200, 2048 + 256
    data = synth.uniform_f32(0, F * K, 1).reshape(F, K)
    masks = synth.uniform_f32(0, n_masks * K, 2).reshape(n_masks, K) - 0.25
    out = eng.masks_dense(dev(data), dev(masks)).cpu().numpy()
    assert eng.last_kernel() in (1, 3, 6)
    assert_close_rel(out, f64_truth(data, masks), 2e-6, abs_scale(data, masks))


@pytest.mark.parametrize('F,K', [(8, 128), (9, 132), (63, 1000), (64, 4096), (65, 4100),
                                 (129, 65536), (1000, 516), (300, 16384)])
def test_dense_tma_shapes(eng, F, K):
    data = synth.uniform_f32(0, F * K, 3).reshape(F, K)
    masks = synth.uniform_f32(0, 11 * K, 4).reshape(11, K)
    out = eng.masks_dense(dev(data), dev(masks)).cpu().numpy()
    assert eng.last_kernel() in (1, 3, 6)
    assert_close_rel(out, f64_truth(data, masks), 2e-6)


def test_dense_tma_strided_accumulate(eng):
    F, K, M = 130, 1024, 6
    big = synth.uniform_f32(0, F * (K + 64), 5).reshape(F, K + 64)
    masks = synth.uniform_f32(0, M * K, 6).reshape(M, K)
    tile = dev(big)[:, 32:32 + K]            # row stride K+64, 128 B aligned offset
    out = torch.full((F, M + 2), 1.5, dtype=torch.float32, device='cuda')
    view = out[:, 1:1 + M]
    eng.masks_dense(tile, dev(masks), out=view, accumulate=True)
    assert eng.last_kernel() in (1, 3, 6)
    res = out.cpu().numpy()
    assert np.all(res[:, 0] == 1.5) and np.all(res[:, -1] == 1.5)
    assert_close_rel(res[:, 1:1 + M] - 1.5, f64_truth(big[:, 32:32 + K], masks), 2e-6)
    eng.masks_dense(tile, dev(masks), out=view, accumulate=False)
    assert_close_rel(out.cpu().numpy()[:, 1:1 + M], f64_truth(big[:, 32:32 + K], masks), 2e-6)


def test_dense_exact_integers(eng):
    # integer-valued fp32 data and binary masks: every partial sum is exact -> bit-exact
    F, K = 500, 16384
    data = synth.poisson3_u16(0, F * K, 7).reshape(F, K).astype(np.float32)
    masks = ring_stack((128, 128), [(8, 16), (20, 28), (32, 40), (44, 52)], 64, 64)
    masks = masks.reshape(4, -1).astype(np.float32)
    out = eng.masks_dense(dev(data), dev(masks)).cpu().numpy()
    exact = data.astype(np.int64) @ masks.astype(np.int64).T
    assert np.array_equal(out.astype(np.int64), exact)


@pytest.mark.parametrize('dt', [np.float32, np.uint16, np.uint8, np.int16])
def test_dense_generic_odd(eng, dt):
    F, K, M = 37, 17 * 23, 5
    if dt == np.float32:
        data = synth.uniform_f32(0, F * K, 8).reshape(F, K)
    else:
        data = (synth.hash_u32(0, F * K, 8) % 200).astype(dt).reshape(F, K)
    masks = mixed_masks(17, 23, M, 9).reshape(M, K)
    sig_sum = torch.zeros(K, dtype=torch.float32, device='cuda')
    t = torch.from_numpy(data.view(np.int16) if dt == np.uint16 else data).cuda()
    if dt == np.uint16:
        t = t.view(torch.uint16)
    out = eng.masks_dense(t, dev(masks), sig_sum=sig_sum).cpu().numpy()
    assert eng.last_kernel() == 2
    assert_close_rel(out, f64_truth(data.astype(np.float32), masks), 2e-6)
    np.testing.assert_allclose(sig_sum.cpu().numpy(), data.astype(np.float64).sum(0), rtol=2e-6)


def test_dense_f64(eng):
    _, g = load_golden('dtypes')
    F, K, M = 21, 300, 3
    data = (synth.hash_u32(0, F * K, 10) % 100000).astype(np.int32).reshape(F, K)
    masks = (synth.uniform_f32(0, M * K, 11).astype(np.float64) * 1.0000001).reshape(M, K)
    out = eng.masks_dense(dev(data), dev(masks)).cpu().numpy()
    assert out.dtype == np.float64
    np.testing.assert_allclose(out, data.astype(np.float64) @ masks.T, rtol=1e-13)


def test_empty_and_tiny(eng):
    masks = dev(np.ones((3, 256), dtype=np.float32))
    out = eng.masks_dense(torch.zeros((0, 256), device='cuda'), masks)
    assert out.shape == (0, 3)
    one = eng.masks_dense(torch.ones((1, 256), device='cuda'), masks).cpu().numpy()
    assert np.array_equal(one, np.full((1, 3), 256, dtype=np.float32))
    z = eng.masks_dense(torch.ones((5, 0), device='cuda'), torch.ones((2, 0), device='cuda'))
    assert z.shape == (5, 2) and float(z.abs().sum()) == 0


def test_errors(eng):
    from libertem_b200._lib import LTB200Error
    with pytest.raises(LTB200Error):
        eng.masks_dense(torch.ones((4, 8)), torch.ones((2, 8)))      # CPU tensors: no fallback
    with pytest.raises(ValueError):
        eng.masks_dense(torch.ones((4, 8), device='cuda'), torch.ones((2, 9), device='cuda'))


def test_csc_cfg3_small_golden(eng):
    import scipy.sparse as sp
    meta, g = load_golden('cfg3_small')
    data = synth.dataset(meta['shape'], np.uint16, meta['data_seed']).reshape(256, 16384)
    stack = ring_stack((128, 128), meta['rings'], 64, 64).reshape(4, -1).astype(np.float32)
    csc = sp.csc_matrix(stack.T)
    t = torch.from_numpy(data.view(np.int16)).cuda().view(torch.uint16)
    out = eng.masks_csc(t, dev(csc.indptr.astype(np.int32)), dev(csc.indices.astype(np.int32)),
                        dev(csc.data.astype(np.float32)), 4).cpu().numpy()
    assert np.array_equal(out, g['intensity'])          # bit-exact integer sums
    # dense path on the u16 tile gives the same
    out2 = eng.masks_dense(t, dev(stack)).cpu().numpy()
    assert np.array_equal(out2, g['intensity'])


@pytest.mark.parametrize('n_masks', [1, 4, 5, 6, 7, 12, 13, 24])
def test_dense_u16_tma(eng, n_masks):
    """uint16 ingest through TMA (converted in registers) + fused SumUDF where it applies:
    integer data, binary/small-integer masks -> bit-exact"""
    F, K = 300, 128 * 128
    data = synth.poisson3_u16(0, F * K, 61).reshape(F, K)
    masks = (synth.hash_u32(0, n_masks * K, 62) % 3).astype(np.float32).reshape(n_masks, K)
    t = torch.from_numpy(data.view(np.int16)).cuda().view(torch.uint16)
    sig_sum = torch.zeros(K, dtype=torch.float32, device='cuda')
    out = eng.masks_dense(t, dev(masks), sig_sum=sig_sum).cpu().numpy()
    assert eng.last_kernel() == 3
    exact = data.astype(np.int64) @ masks.astype(np.int64).T
    assert np.array_equal(out.astype(np.int64), exact)
    assert np.array_equal(sig_sum.cpu().numpy().astype(np.int64), data.astype(np.int64).sum(0))
    # float weights too
    fm = synth.uniform_f32(0, n_masks * K, 63).reshape(n_masks, K)
    out = eng.masks_dense(t, dev(fm)).cpu().numpy()
    assert_close_rel(out, f64_truth(data.astype(np.float32), fm), 2e-6)


def test_dense_u16_values_full_range(eng):
    F, K = 64, 1024
    data = (synth.hash_u32(0, F * K, 64) & 0xFFFF).astype(np.uint16).reshape(F, K)
    data[0, :4] = [0, 1, 65535, 32768]
    ones = np.ones((1, K), dtype=np.float32)
    t = torch.from_numpy(data.view(np.int16)).cuda().view(torch.uint16)
    out = eng.masks_dense(t, dev(ones)).cpu().numpy()
    assert eng.last_kernel() == 3
    want = data.astype(np.float64).sum(1)
    np.testing.assert_allclose(out[:, 0], want, rtol=1e-6)
    e = np.zeros((4, K), dtype=np.float32)
    e[np.arange(4), np.arange(4)] = 1
    out = eng.masks_dense(t, dev(e)).cpu().numpy()
    assert np.array_equal(out[0], [0, 1, 65535, 32768])


def test_full_size_properties(eng):
    """BASELINE cfg2 sig size at a nav size the test box handles quickly: linearity and a
    checksum against an independent (torch fp64 on device) evaluation of a frame subsample."""
    F, K, M = 4096, 65536, 11
    data = eng.synth_fill((F, K), np.float32, 21, 'cuda')
    masks = dev(np.concatenate([mixed_masks(256, 256, 8, 22).reshape(8, -1),
                                O.com_mask_stack((256, 256), 128, 128).reshape(3, -1)]))
    out = eng.masks_dense(data, masks)
    assert eng.last_kernel() in (1, 3, 6)
    sel = torch.arange(0, F, 97, device='cuda')
    truth = data[sel].double() @ masks.double().T
    scale = (data[sel].double().abs() @ masks.double().abs().T).amax(0, keepdim=True)
    err = ((out[sel].double() - truth).abs() / scale).max().item()
    assert err <= 2e-6, err
    # linearity: masks scaled by 2 (exact in fp32) -> results exactly doubled
    out2 = eng.masks_dense(data, masks * 2)
    assert torch.equal(out2, out * 2)
    # host twin of a frame sample agrees with the oracle BLAS path
    hf = synth.uniform_f32(5 * K, K, 21).reshape(1, K)
    ref = O.process_flat(hf, masks.cpu().numpy().T.copy())
    np.testing.assert_allclose(out[5:6].cpu().numpy(), ref, rtol=RTOL)
