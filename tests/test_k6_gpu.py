"""GPU parity tests of K6, the tcgen05 tensor-core form of the dense masked reduction
(ltb200_masks_dense_tc: split-TF32 MMAs with TMEM accumulators), against float64 truth, the CPU
oracle and the committed golden vectors.  Same seam as ApplyMasksEngine.process_flat (reference
udf/masks.py:31-83).  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from golden_inputs import mixed_masks, ring_stack
from oracle import synth, udf_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: <= 1e-5 rel for float32 mask/CoM results
TIGHT = 2e-6  # what the kernel is expected to hold against float64 on the sum|x||m| scale


@pytest.fixture(scope='module')
def eng():
    from libertem_b200 import engine
    assert torch.cuda.is_available()
    engine.set_k1_variant(0)
    return engine


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def f64_truth(tile, masks):
    return tile.astype(np.float64) @ masks.astype(np.float64).T


def abs_scale(tile, masks):
    return (np.abs(tile).astype(np.float64) @ np.abs(masks).astype(np.float64).T).max(
        axis=0, keepdims=True)


def assert_close_rel(res, truth, rtol, scale=None):
    if scale is None:
        scale = np.abs(truth).max(axis=0, keepdims=True)
    err = np.abs(res - truth) / (scale + 1e-30)
    assert err.max() <= rtol, f'max rel err {err.max():.3e}'


@pytest.mark.parametrize('n_masks', [1, 3, 8, 9, 11, 12, 13, 16, 17, 19, 24, 25, 32, 33, 40, 70])
def test_tc_mask_counts(eng, n_masks):
    F, K = 600, 4096 + 256
    data = synth.uniform_f32(0, F * K, 1).reshape(F, K)
    masks = synth.uniform_f32(0, n_masks * K, 2).reshape(n_masks, K) - 0.25
    out = eng.masks_dense_tc(dev(data), dev(masks)).cpu().numpy()
    assert eng.last_kernel() == 6
    assert_close_rel(out, f64_truth(data, masks), TIGHT, abs_scale(data, masks))


@pytest.mark.parametrize('F,K', [(1, 128), (8, 128), (255, 132), (256, 4096), (257, 4100),
                                 (1000, 516), (513, 65536), (300, 16384), (5000, 1024)])
def test_tc_shapes(eng, F, K):
    # ragged frame counts (TMA zero-fills the rows past the end) and signal sizes that are not
    # multiples of the 32-pixel sub-stage
    data = synth.uniform_f32(0, F * K, 3).reshape(F, K)
    masks = synth.uniform_f32(0, 19 * K, 4).reshape(19, K)
    out = eng.masks_dense_tc(dev(data), dev(masks)).cpu().numpy()
    assert_close_rel(out, f64_truth(data, masks), TIGHT)


@pytest.mark.parametrize('chain', [0, 1, 2, 3, 5, 8])
def test_tc_chain_lengths(eng, chain):
    # every chain length drains its TMEM accumulators correctly (ring / double-buffer logic);
    # the longer the chain, the larger the truncation bias of the tensor-core accumulate
    F, K, M = 700, 8192 + 32 * 3, 19
    data = synth.uniform_f32(0, F * K, 5).reshape(F, K)
    masks = synth.uniform_f32(0, M * K, 6).reshape(M, K)
    out = eng.masks_dense_tc(dev(data), dev(masks), chain=chain).cpu().numpy()
    assert_close_rel(out, f64_truth(data, masks), 5e-6)


@pytest.mark.parametrize('n_masks', [25, 28, 32])
@pytest.mark.parametrize('chain', [0, 2, 5])
def test_tc_drain_warps(eng, n_masks, chain, monkeypatch):
    # 25-32 columns: the accumulators are drained by eight extra warps (two-level sums of the
    # chain totals); LTB200_K6_DW=0 is the 12-warp form with the converters draining.  Ragged
    # frame count, a K split and a signal size that is not a multiple of the chain length.
    F, K = 1100, 8192 + 96
    data = synth.uniform_f32(0, F * K, 21).reshape(F, K)
    masks = synth.uniform_f32(0, n_masks * K, 22).reshape(n_masks, K) - 0.25
    truth, scale = f64_truth(data, masks), abs_scale(data, masks)
    monkeypatch.setenv('LTB200_K6_DW', '0')
    out_cv = eng.masks_dense_tc(dev(data), dev(masks), chain=chain).cpu().numpy()
    tol = TIGHT if chain in (0, 1) else 5e-6
    assert_close_rel(out_cv, truth, tol, scale)
    outs = []
    for mode in ('1', '2'):                  # 8 / 16 converter warps in front of the drain warps
        monkeypatch.setenv('LTB200_K6_DW', mode)
        out_dw = eng.masks_dense_tc(dev(data), dev(masks), chain=chain).cpu().numpy()
        assert eng.last_kernel() == 6
        assert_close_rel(out_dw, truth, tol, scale)
        assert_close_rel(out_dw, out_cv, 2e-6, scale)
        outs.append(out_dw)
    assert np.array_equal(outs[0], outs[1])  # same arithmetic, different warp layout


@pytest.mark.parametrize('n_masks', [11, 32])
def test_tc_three_products(eng, n_masks, monkeypatch):
    # 9-16 and 25-32 columns drop the lo(x) * lo(mask) products (LTB200_K6_THREE=0 keeps them):
    # below 2^-21 |x||m| per term; constant masks are the worst case (lo(mask) of one sign)
    F, K = 700, 16384
    data = synth.uniform_f32(0, F * K, 31).reshape(F, K)
    masks = synth.uniform_f32(0, n_masks * K, 32).reshape(n_masks, K) - 0.25
    masks[0] = 0.3
    masks[1] = -1.0 / 3.0
    masks[2] = 1.0 + 2.0 ** -12 - 2.0 ** -23       # largest lo(mask) relative to the value
    truth, scale = f64_truth(data, masks), abs_scale(data, masks)
    outs = {}
    for three in ('0', '1'):
        monkeypatch.setenv('LTB200_K6_THREE', three)
        outs[three] = eng.masks_dense_tc(dev(data), dev(masks)).cpu().numpy()
        assert_close_rel(outs[three], truth, TIGHT, scale)
    assert not np.array_equal(outs['0'], outs['1'])
    # element-wise the two differ by their (independent) accumulation rounding; the dropped term
    # itself shows as a shift of the column mean
    assert_close_rel(outs['1'], outs['0'], 8e-7, scale)
    shift = np.abs((outs['1'].astype(np.float64) - outs['0']).mean(axis=0)) / scale[0]
    assert shift.max() <= 1.5e-7, f'mean shift {shift.max():.3e}'


def test_tc_drain_warps_many_items(eng, monkeypatch):
    # many items per CTA (the chain counter and its mbarrier parities run across items), then a
    # long signal with few frames; accumulate into a strided view
    monkeypatch.delenv('LTB200_K6_DW', raising=False)       # the default form
    for F, K, M in ((148 * 256 * 3 + 77, 256, 32), (300, 65536, 29)):
        data = synth.uniform_f32(0, F * K, 23).reshape(F, K)
        masks = synth.uniform_f32(0, M * K, 24).reshape(M, K) - 0.5
        out = torch.full((F, M + 3), 2.0, dtype=torch.float32, device='cuda')
        view = out[:, 2:2 + M]
        eng.masks_dense_tc(dev(data), dev(masks), out=view, accumulate=True)
        res = out.cpu().numpy()
        assert np.all(res[:, :2] == 2.0) and np.all(res[:, -1] == 2.0)
        assert_close_rel(res[:, 2:2 + M] - 2.0, f64_truth(data, masks), TIGHT,
                         abs_scale(data, masks))
        again = eng.masks_dense_tc(dev(data), dev(masks)).cpu().numpy()
        assert np.array_equal(again, eng.masks_dense_tc(dev(data), dev(masks)).cpu().numpy())


def test_tc_strided_accumulate(eng):
    F, K, M = 530, 1024, 6
    big = synth.uniform_f32(0, F * (K + 64), 5).reshape(F, K + 64)
    masks = synth.uniform_f32(0, M * K, 6).reshape(M, K)
    tile = dev(big)[:, 32:32 + K]            # row stride K+64, 128 B aligned offset
    out = torch.full((F, M + 2), 1.5, dtype=torch.float32, device='cuda')
    view = out[:, 1:1 + M]
    eng.masks_dense_tc(tile, dev(masks), out=view, accumulate=True)
    res = out.cpu().numpy()
    assert np.all(res[:, 0] == 1.5) and np.all(res[:, -1] == 1.5)
    assert_close_rel(res[:, 1:1 + M] - 1.5, f64_truth(big[:, 32:32 + K], masks), TIGHT)
    eng.masks_dense_tc(tile, dev(masks), out=view, accumulate=False)
    assert_close_rel(out.cpu().numpy()[:, 1:1 + M], f64_truth(big[:, 32:32 + K], masks), TIGHT)


def test_tc_exact_integers(eng):
    # integer-valued fp32 data (16-bit counts need the hi AND the lo part) and binary masks:
    # every product and every partial sum is exact -> bit-exact
    F, K = 500, 16384
    data = synth.poisson3_u16(0, F * K, 7).reshape(F, K).astype(np.float32)
    data[:, ::97] += 40000.0
    masks = ring_stack((128, 128), [(8, 16), (20, 28), (32, 40), (44, 52)], 64, 64)
    masks = masks.reshape(4, -1).astype(np.float32)
    out = eng.masks_dense_tc(dev(data), dev(masks)).cpu().numpy()
    exact = data.astype(np.int64) @ masks.astype(np.int64).T
    assert np.array_equal(out.astype(np.int64), exact)


def test_tc_special_values(eng):
    # signed zeros, tiny and huge magnitudes, negative data; delta masks pick single pixels, so
    # this checks the hi/lo split value by value (|error| <= 2^-21 |x|)
    F, K = 256, 256
    data = synth.uniform_f32(0, F * K, 8).reshape(F, K) - 0.5
    data[0, :8] = [0.0, -0.0, 1e-30, -1e-30, 3e38, -3e38, 1.0, -1.0]
    data[1, :4] = [16777215.0, 1.0000001, 0.99999994, 123456.789]
    e = np.zeros((12, K), dtype=np.float32)
    e[np.arange(12), np.arange(12)] = 1
    out = eng.masks_dense_tc(dev(data), dev(e)).cpu().numpy()
    want = data[:, :12]
    assert np.all(np.abs(out - want) <= np.abs(want) * 2.0 ** -21)


def test_tc_cfg2_small_golden(eng):
    meta, g = load_golden('cfg2_small')
    data = synth.dataset(meta['shape'], np.float32, meta['data_seed']).reshape(256, 65536)
    stack = mixed_masks(256, 256, 8, meta['mask_seed']).reshape(8, -1)
    com = O.com_mask_stack((256, 256), 128, 128).reshape(3, -1)
    ones = np.ones((1, 65536), dtype=np.float32)
    allm = np.concatenate([stack, com, ones])          # 12 columns in ONE pass
    out = eng.masks_dense_tc(dev(data), dev(allm)).cpu().numpy()
    truth = f64_truth(data, allm)
    assert_close_rel(out, truth, TIGHT, abs_scale(data, allm))
    assert_close_rel(out[:, :8], g['intensity'], RTOL)
    assert_close_rel(out[:, 8:11], g['com_raw_mask_result'], RTOL)
    np.testing.assert_allclose(out[:, 11], g['sumsig'], rtol=RTOL)


def test_tc_matches_oracle_process_flat(eng):
    F, K, M = 64, 4096, 5
    data = synth.uniform_f32(0, F * K, 9).reshape(F, K)
    masks = mixed_masks(64, 64, M, 10).reshape(M, K)
    ref = O.process_flat(data, masks.T.copy())
    out = eng.masks_dense_tc(dev(data), dev(masks)).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=RTOL, atol=1e-5 * np.abs(ref).max())


def test_tc_routing(eng):
    """ltb200_masks_dense sends wide float32 stacks to K6 by itself; variant 3 forces it"""
    F, K = 2048, 4096
    data = eng.synth_fill((F, K), np.float32, 11, 'cuda')
    wide = eng.synth_fill((19, K), np.float32, 12, 'cuda')
    out = eng.masks_dense(data, wide)
    assert eng.last_kernel() == 6
    eng.set_k1_variant(2)
    try:
        ref = eng.masks_dense(data, wide)
        assert eng.last_kernel() == 3
    finally:
        eng.set_k1_variant(0)
    err = ((out - ref).abs().amax(0) / ref.abs().amax(0)).max().item()
    assert err <= 3e-6, err
    # fused frame sum still delivered when K6 takes the mask columns
    sig = torch.zeros(K, dtype=torch.float32, device='cuda')
    eng.masks_dense(data, wide, sig_sum=sig)
    assert eng.last_kernel() == 6
    np.testing.assert_allclose(sig.cpu().numpy(), data.double().sum(0).cpu().numpy(), rtol=2e-6)
    # uint16 tiles with more than 16 columns stay on the FFMA2 kernel
    t16 = eng.synth_fill((F, K), np.uint16, 13, 'cuda')
    eng.set_k1_variant(3)
    try:
        eng.masks_dense(t16, wide)
        assert eng.last_kernel() == 3
        eng.masks_dense(data, wide[:3])
        assert eng.last_kernel() == 6
    finally:
        eng.set_k1_variant(0)


def test_tc_unsupported_shapes(eng):
    from libertem_b200._lib import LTB200Error
    with pytest.raises(LTB200Error):
        eng.masks_dense_tc(torch.ones((300, 130), device='cuda'), torch.ones((2, 130), device='cuda'))
    with pytest.raises(LTB200Error):
        eng.masks_dense_tc(torch.ones((300, 64), device='cuda'), torch.ones((2, 64), device='cuda'))
    with pytest.raises(TypeError):
        eng.masks_dense_tc(torch.ones((300, 256), device='cuda', dtype=torch.float64),
                           torch.ones((2, 256), device='cuda'))
    # the generic entry point takes all of them
    out = eng.masks_dense(torch.ones((300, 130), device='cuda'), torch.ones((2, 130), device='cuda'))
    assert torch.equal(out, torch.full((300, 2), 130.0, device='cuda'))


def test_tc_full_size_properties(eng):
    """BASELINE cfg5 geometry (256x256 signal, 16 masks + CoM = 19 columns) at a frame count the
    test box handles quickly: checksum against torch float64 on a frame subsample, exact
    linearity, agreement with the FFMA2 kernel, and the host twin of one frame vs the oracle."""
    F, K = 8192, 65536
    data = eng.synth_fill((F, K), np.float32, 21, 'cuda')
    masks = dev(np.concatenate([mixed_masks(256, 256, 16, 22).reshape(16, -1),
                                O.com_mask_stack((256, 256), 128, 128).reshape(3, -1)]))
    out = eng.masks_dense(data, masks)
    assert eng.last_kernel() == 6
    sel = torch.arange(0, F, 97, device='cuda')
    truth = data[sel].double() @ masks.double().T
    scale = (data[sel].double().abs() @ masks.double().abs().T).amax(0, keepdim=True)
    err = ((out[sel].double() - truth).abs() / scale).max().item()
    assert err <= TIGHT, err
    out2 = eng.masks_dense(data, masks * 2)
    assert torch.equal(out2, out * 2)
    eng.set_k1_variant(2)
    try:
        ref = eng.masks_dense(data, masks)
    finally:
        eng.set_k1_variant(0)
    d = ((out - ref).abs().amax(0) / scale.float().squeeze(0)).max().item()
    assert d <= 3e-6, d
    hf = synth.uniform_f32(5 * K, K, 21).reshape(1, K)
    ref1 = O.process_flat(hf, masks.cpu().numpy().T.copy())
    # (the signed gradient masks cancel to ~1e-3 of their sum|x||m| scale, where neither BLAS
    # nor any float32 kernel holds 1e-5 of the value itself: compare on the column scale)
    np.testing.assert_allclose(out[5:6].cpu().numpy(), ref1, rtol=RTOL,
                               atol=RTOL * float(scale.max()) * 0.3)


# ---- uint16 tiles on the tensor cores (ltb200_masks_dense_tc_u16) + fused frame sum ----------

def dev_u16(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).cuda().view(torch.uint16)


def as_f64(t16):
    """uint16 CUDA tensor -> float64 (through int16 views: torch has few uint16 kernels)"""
    return (t16.view(torch.int16).to(torch.int64) & 0xFFFF).double()


def full_range_u16(F, K, seed):
    return (synth.hash_u32(0, F * K, seed) & 0xFFFF).astype(np.uint16).reshape(F, K)


@pytest.mark.parametrize('n_masks', [1, 5, 8, 9, 16])
def test_tcu16_mask_counts(eng, n_masks):
    # all 16-bit values (hi and lo parts both in use), float weights, a signal size that is
    # not a multiple of the 64-pixel stage
    F, K = 600, 4096 + 8 * 5
    data = full_range_u16(F, K, 31)
    masks = synth.uniform_f32(0, n_masks * K, 32).reshape(n_masks, K) - 0.25
    sig = torch.zeros(K, dtype=torch.float32, device='cuda')
    out = eng.masks_dense_tc_u16(dev_u16(data), dev(masks), sig_sum=sig).cpu().numpy()
    assert eng.last_kernel() == 6
    f = data.astype(np.float32)
    assert_close_rel(out, f64_truth(f, masks), TIGHT, abs_scale(f, masks))
    assert np.array_equal(sig.cpu().numpy(), data.astype(np.int64).sum(0).astype(np.float32))


@pytest.mark.parametrize('F,K', [(1, 256), (8, 256), (255, 264), (257, 4104), (1000, 520),
                                 (300, 16384), (40960, 1024), (5000, 65536)])
def test_tcu16_shapes_exact(eng, F, K):
    # Poisson counts with a few large outliers x small-integer masks: bit-exact results and an
    # exact fused frame sum, for ragged frame counts, several items per CTA and split-K shapes
    data = synth.poisson3_u16(0, F * K, 33).reshape(F, K).copy()
    data[:, ::1013] += 40000
    masks = (synth.hash_u32(0, 5 * K, 34) % 3).astype(np.float32).reshape(5, K)
    masks[4] = 1
    t = dev_u16(data)
    sig = torch.zeros(K, dtype=torch.float32, device='cuda')
    out = eng.masks_dense_tc_u16(t, dev(masks), sig_sum=sig)
    tt = as_f64(t)
    exact = tt @ dev(masks).double().T          # integers < 2^24: exact in every precision
    assert exact.max().item() < 2 ** 24
    assert torch.equal(out.double(), exact)
    assert torch.equal(sig, tt.sum(0).float())
    # without the frame sum (no summing warps on the stage barrier)
    out2 = eng.masks_dense_tc_u16(t, dev(masks))
    assert torch.equal(out2, out)


@pytest.mark.parametrize('chain', [0, 1, 2, 3, 5, 8])
def test_tcu16_chain_lengths(eng, chain):
    F, K, M = 700, 8192 + 64 * 3 + 8, 11
    data = full_range_u16(F, K, 35)
    masks = synth.uniform_f32(0, M * K, 36).reshape(M, K)
    out = eng.masks_dense_tc_u16(dev_u16(data), dev(masks), chain=chain).cpu().numpy()
    assert_close_rel(out, f64_truth(data.astype(np.float32), masks), 5e-6)


def test_tcu16_strided_accumulate(eng):
    F, K, M = 530, 1024, 6
    big = full_range_u16(F, K + 128, 37)
    masks = synth.uniform_f32(0, M * K, 38).reshape(M, K)
    tile = dev_u16(big)[:, 64:64 + K]          # row stride K+128, 128 B aligned offset
    out = torch.full((F, M + 2), 1.5, dtype=torch.float32, device='cuda')
    view = out[:, 1:1 + M]
    eng.masks_dense_tc_u16(tile, dev(masks), out=view, accumulate=True)
    res = out.cpu().numpy()
    truth = f64_truth(big[:, 64:64 + K].astype(np.float32), masks)
    assert np.all(res[:, 0] == 1.5) and np.all(res[:, -1] == 1.5)
    assert_close_rel(res[:, 1:1 + M] - 1.5, truth, TIGHT)
    eng.masks_dense_tc_u16(tile, dev(masks), out=view, accumulate=False)
    assert_close_rel(out.cpu().numpy()[:, 1:1 + M], truth, TIGHT)


def test_tcu16_frame_sum_beyond_32_bits(eng):
    # 70 000 frames of 65535 at some pixels: the per-pixel sum (4.6e9) exceeds 32 bits
    F, K = 70000, 256
    data = np.zeros((F, K), dtype=np.uint16)
    data[:, 3] = 65535
    data[:, 200] = 65535
    data[::2, 77] = 1
    sig = torch.zeros(K, dtype=torch.float32, device='cuda')
    ones = torch.ones((1, K), dtype=torch.float32, device='cuda')
    out = eng.masks_dense_tc_u16(dev_u16(data), ones, sig_sum=sig).cpu().numpy()
    assert np.array_equal(out[:, 0], data.astype(np.int64).sum(1).astype(np.float32))
    want = data.astype(np.int64).sum(0)
    assert want[3] == 65535 * F > 2 ** 32
    assert np.array_equal(sig.cpu().numpy(), want.astype(np.float32))


def test_tcu16_routing(eng):
    """ltb200_masks_dense sends uint16 tiles of >= 1024 frames and 7..16 columns to the
    tensor-core kernel, frame sum included; results are identical to the FFMA2 kernel's.
    Narrower stacks stay on the FFMA2 kernel (faster there) unless variant 3 forces K6."""
    F, K = 4096, 128 * 128
    t = eng.synth_fill((F, K), np.uint16, 39, 'cuda')
    rings = ring_stack((128, 128), [(8, 16), (20, 28), (32, 40), (44, 52), (4, 60), (0, 9),
                                    (50, 64), (30, 31)], 64, 64)
    masks = dev(np.concatenate([np.ones((1, K), np.float32),
                                rings.reshape(8, -1).astype(np.float32)]))
    sig = torch.zeros(K, dtype=torch.float32, device='cuda')
    out = eng.masks_dense(t, masks, sig_sum=sig)
    assert eng.last_kernel() == 6
    eng.set_k1_variant(2)
    try:
        sig2 = torch.zeros(K, dtype=torch.float32, device='cuda')
        ref = eng.masks_dense(t, masks, sig_sum=sig2)
        assert eng.last_kernel() == 3
    finally:
        eng.set_k1_variant(0)
    assert torch.equal(out, ref) and torch.equal(sig, sig2)
    tt = as_f64(t)
    assert torch.equal(out.double(), tt @ masks.double().T)
    assert torch.equal(sig.double(), tt.sum(0))
    # <= 6 columns and small tiles stay on the FFMA2 kernel
    eng.masks_dense(t, masks[:5], sig_sum=sig)
    assert eng.last_kernel() == 3
    eng.masks_dense(t[:512], masks)
    assert eng.last_kernel() == 3
    eng.set_k1_variant(3)
    try:
        out5 = eng.masks_dense(t, masks[:5])
        assert eng.last_kernel() == 6
    finally:
        eng.set_k1_variant(0)
    assert torch.equal(out5, out[:, :5])


def test_tcu16_unsupported(eng):
    from libertem_b200._lib import LTB200Error
    t = torch.zeros((300, 1024), dtype=torch.uint16, device='cuda')
    with pytest.raises(LTB200Error):
        eng.masks_dense_tc_u16(t, torch.ones((17, 1024), device='cuda'))
    with pytest.raises(LTB200Error):
        eng.masks_dense_tc_u16(t[:, :128], torch.ones((2, 128), device='cuda'))
    with pytest.raises(TypeError):
        eng.masks_dense_tc_u16(t.float(), torch.ones((2, 1024), device='cuda'))
