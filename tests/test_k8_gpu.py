"""GPU parity tests of K8, the integer fast path (ltb200_masks_dense_i8): uint16 tiles x int8
masks on the int8 tensor cores (tcgen05.mma kind::i8) with the frame sum of SumUDF as a second
MMA.  Integer data x integer masks is exact in the reference's float32 arithmetic (udf/masks.py:
59-77, udf/sum.py:44-49) as long as the sums stay below 2^24, so every result here is compared
bit for bit with the exact integer answer.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

from golden_inputs import ring_stack
from oracle import synth, udf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from libertem_b200 import engine
    assert torch.cuda.is_available()
    engine.set_k1_variant(0)
    return engine


def dev_u16(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).cuda().view(torch.uint16)


def as_f64(t16):
    return (t16.view(torch.int16).to(torch.int64) & 0xFFFF).double()


def int_masks(M, K, seed, lo=-3, hi=3):
    m = (synth.hash_u32(0, M * K, seed) % (hi - lo + 1)).astype(np.int64) + lo
    return m.astype(np.int8).reshape(M, K)


def check_exact(eng, data, masks, with_sum=True):
    t = dev_u16(data)
    m = torch.from_numpy(masks).cuda()
    K = data.shape[1]
    sig = torch.zeros(K, dtype=torch.float32, device='cuda') if with_sum else None
    out = eng.masks_dense_i8(t, m, sig_sum=sig)
    assert eng.last_kernel() == 8
    tt = as_f64(t)
    exact = tt @ m.double().T                      # < 2^53: exact
    assert torch.equal(out, exact.float()), (out.double() - exact).abs().max().item()
    if with_sum:
        assert torch.equal(sig, tt.sum(0).float())
    return out


@pytest.mark.parametrize('n_masks', [1, 2, 5, 8, 9, 13, 16])
def test_i8_mask_counts(eng, n_masks):
    # all 16-bit values (both bytes in use), signed weights over the whole int8 range
    F, K = 600, 4096 + 8 * 5
    data = (synth.hash_u32(0, F * K, 41) & 0xFFFF).astype(np.uint16).reshape(F, K)
    masks = int_masks(n_masks, K, 42, -128, 127)
    masks[0] = 1
    check_exact(eng, data, masks)


@pytest.mark.parametrize('F,K', [(1, 256), (8, 256), (255, 264), (257, 4104), (1000, 520),
                                 (300, 16384), (40960, 1024), (3000, 65536)])
def test_i8_shapes(eng, F, K):
    # ragged frame counts, signal sizes that are not multiples of the 64-pixel stage, several
    # items per CTA, split-K shapes; with and without the fused frame sum
    data = synth.poisson3_u16(0, F * K, 43).reshape(F, K).copy()
    data[:, ::1013] += 40000
    masks = int_masks(5, K, 44, 0, 2)
    masks[4] = 1
    out = check_exact(eng, data, masks)
    out2 = check_exact(eng, data, masks, with_sum=False)
    assert torch.equal(out, out2)


def test_i8_matches_reference_formulation(eng):
    # the reference's own float32 GEMM on the same integers (oracle process_flat) is exact
    # below 2^24, so it agrees bit for bit
    F, K = 512, 128 * 128
    data = synth.poisson3_u16(0, F * K, 45).reshape(F, K)
    rings = ring_stack((128, 128), [(8, 16), (20, 28), (32, 40), (44, 52)], 64, 64)
    masks = np.concatenate([np.ones((1, K), np.int8), rings.reshape(4, -1).astype(np.int8)])
    out = check_exact(eng, data, masks).cpu().numpy()
    ref = O.process_flat(data.astype(np.float32), masks.astype(np.float32).T.copy())
    assert np.array_equal(out, ref)
    # and with the float-mask kernels of this library
    f = eng.masks_dense(dev_u16(data), torch.from_numpy(masks.astype(np.float32)).cuda())
    assert np.array_equal(out, f.cpu().numpy())


def test_i8_strided_accumulate(eng):
    F, K, M = 530, 1024, 6
    big = (synth.hash_u32(0, F * (K + 128), 46) & 0xFFFF).astype(np.uint16).reshape(F, K + 128)
    masks = int_masks(M, K, 47, -5, 5)
    tile = dev_u16(big)[:, 64:64 + K]          # row stride K+128, 128 B aligned offset
    m = torch.from_numpy(masks).cuda()
    out = torch.full((F, M + 2), 1.5, dtype=torch.float32, device='cuda')
    view = out[:, 1:1 + M]
    eng.masks_dense_i8(tile, m, out=view, accumulate=True)
    exact = (as_f64(dev_u16(big))[:, 64:64 + K] @ m.double().T)
    assert torch.all(out[:, 0] == 1.5) and torch.all(out[:, -1] == 1.5)
    assert torch.equal(out[:, 1:1 + M], exact.float() + 1.5)
    eng.masks_dense_i8(tile, m, out=view, accumulate=False)
    assert torch.equal(out[:, 1:1 + M], exact.float())


def test_i8_frame_sum_beyond_32_bits(eng):
    F, K = 70000, 256
    data = np.zeros((F, K), dtype=np.uint16)
    data[:, 3] = 65535
    data[:, 200] = 65535
    data[::2, 77] = 1
    sig = torch.zeros(K, dtype=torch.float32, device='cuda')
    ones = torch.ones((1, K), dtype=torch.int8, device='cuda')
    out = eng.masks_dense_i8(dev_u16(data), ones, sig_sum=sig).cpu().numpy()
    assert np.array_equal(out[:, 0], data.astype(np.int64).sum(1).astype(np.float32))
    want = data.astype(np.int64).sum(0)
    assert want[3] == 65535 * F > 2 ** 32
    assert np.array_equal(sig.cpu().numpy(), want.astype(np.float32))
    # accumulates into sig_sum
    eng.masks_dense_i8(dev_u16(data), ones, sig_sum=sig)
    assert np.array_equal(sig.cpu().numpy(), want.astype(np.float32) * 2)


def test_i8_extreme_sums(eng):
    # the largest accumulators the shape check admits: 65536 pixels of 65535 x 127
    F, K = 260, 65536
    data = np.full((F, K), 65535, dtype=np.uint16)
    masks = np.stack([np.full(K, 127, np.int8), np.full(K, -128, np.int8)])
    check_exact(eng, data, masks)


def test_i8_unsupported(eng):
    from libertem_b200._lib import LTB200Error
    t = torch.zeros((300, 1024), dtype=torch.uint16, device='cuda')
    with pytest.raises(LTB200Error):
        eng.masks_dense_i8(t, torch.ones((33, 1024), dtype=torch.int8, device='cuda'))
    with pytest.raises(LTB200Error):
        eng.masks_dense_i8(t[:, :128], torch.ones((2, 128), dtype=torch.int8, device='cuda'))
    with pytest.raises(TypeError):
        eng.masks_dense_i8(t, torch.ones((2, 1024), device='cuda'))
    K = 64 * 65536 + 64                          # beyond the 64 forced K splits
    big = torch.zeros((2, K), dtype=torch.uint16, device='cuda')
    with pytest.raises(LTB200Error):
        eng.masks_dense_i8(big, torch.ones((1, K), dtype=torch.int8, device='cuda'))


@pytest.mark.parametrize('n_masks', [17, 24, 32])
@pytest.mark.parametrize('dt', ['u16', 'u8'])
def test_i8_wide_stacks(eng, n_masks, dt):
    """17..32 int8 rows per pass (N = 64 for uint16 tiles with one accumulator buffer, N = 32
    for uint8 tiles), several items per CTA, with the fused frame sum"""
    F, K = 40960, 1024
    masks = int_masks(n_masks, K, 72, -128, 127)
    masks[0] = 1
    if dt == 'u16':
        data = (synth.hash_u32(0, F * K, 71) & 0xFFFF).astype(np.uint16).reshape(F, K)
        check_exact(eng, data, masks)
        check_exact(eng, data[:300], masks, with_sum=False)
    else:
        data = (synth.hash_u32(0, F * K, 71) & 0xFF).astype(np.uint8).reshape(F, K)
        check_exact_u8(eng, data, masks)
        check_exact_u8(eng, data[:300], masks, with_sum=False)


@pytest.mark.parametrize('K', [65536 + 64, 512 * 512, 3 * 65536 + 8 * 5])
def test_i8_large_signals(eng, K):
    """signals beyond 65536 pixels are K-split so that every int32 accumulator stays exact; the
    int64 recombination gives the correctly rounded exact sum"""
    F = 300
    data = np.full((F, K), 65535, dtype=np.uint16)
    data[:, ::7] = (synth.hash_u32(0, F * len(range(0, K, 7)), 61) & 0xFFFF).astype(
        np.uint16).reshape(F, -1)
    masks = np.stack([np.full(K, 127, np.int8), np.full(K, -128, np.int8),
                      int_masks(1, K, 62, -3, 3)[0]])
    check_exact(eng, data, masks)


# ---- uint8 tiles (one byte per pixel) -----------------------------------------------------------

def check_exact_u8(eng, data, masks, with_sum=True):
    t = torch.from_numpy(np.ascontiguousarray(data)).cuda()
    m = torch.from_numpy(masks).cuda()
    K = data.shape[1]
    sig = torch.zeros(K, dtype=torch.float32, device='cuda') if with_sum else None
    out = eng.masks_dense_i8(t, m, sig_sum=sig)
    assert eng.last_kernel() == 8
    tt = t.double()
    exact = tt @ m.double().T
    assert torch.equal(out, exact.float()), (out.double() - exact).abs().max().item()
    if with_sum:
        assert torch.equal(sig, tt.sum(0).float())
    return out


@pytest.mark.parametrize('n_masks', [1, 5, 8, 9, 16])
def test_i8_u8_mask_counts(eng, n_masks):
    F, K = 600, 4096 + 16 * 3
    data = (synth.hash_u32(0, F * K, 51) & 0xFF).astype(np.uint8).reshape(F, K)
    masks = int_masks(n_masks, K, 52, -128, 127)
    masks[0] = 1
    check_exact_u8(eng, data, masks)


@pytest.mark.parametrize('F,K', [(1, 512), (255, 528), (257, 4112), (1000, 1040), (300, 16384),
                                 (40960, 1024), (3000, 65536)])
def test_i8_u8_shapes(eng, F, K):
    data = (synth.hash_u32(0, F * K, 53) % 7).astype(np.uint8).reshape(F, K)
    data[:, ::1013] = 255
    masks = int_masks(5, K, 54, 0, 2)
    masks[4] = 1
    out = check_exact_u8(eng, data, masks)
    out2 = check_exact_u8(eng, data, masks, with_sum=False)
    assert torch.equal(out, out2)
    # the float kernels of this library agree (uint8 ingest of the generic path)
    f = eng.masks_dense(torch.from_numpy(data).cuda(),
                        torch.from_numpy(masks.astype(np.float32)).cuda())
    assert torch.equal(out, f)


def test_i8_u8_unsupported(eng):
    from libertem_b200._lib import LTB200Error
    t = torch.zeros((300, 1024), dtype=torch.uint8, device='cuda')
    with pytest.raises(LTB200Error):
        eng.masks_dense_i8(t[:, :256], torch.ones((2, 256), dtype=torch.int8, device='cuda'))
    with pytest.raises(LTB200Error):
        eng.masks_dense_i8(t[:, :1000], torch.ones((2, 1000), dtype=torch.int8, device='cuda'))
    with pytest.raises(TypeError):
        eng.masks_dense_i8(t.to(torch.int16), torch.ones((2, 1024), dtype=torch.int8,
                                                          device='cuda'))
