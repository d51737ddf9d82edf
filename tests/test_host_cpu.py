"""CPU-only tests (pytest -m "not gpu"): host logic of the reference-API mirror, the C-ABI
library's exported symbols, and the no-CPU-fallback guarantee.  No compute kernels run here."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest
import torch

from conftest import load_golden, ROOT
from golden_inputs import mixed_masks
from oracle import udf_oracle as O

from libertem_b200 import _lib, masks as M
from libertem_b200.common import Shape, Slice
from libertem_b200.common.container import MaskContainer, full_sig_slice
from libertem_b200.common.buffers import BufferWrapper
from libertem_b200.io.memory import partition_boundaries, MemoryDataSet
from libertem_b200.udf import ApplyMasksUDF, CoMUDF, SumUDF, SumSigUDF, CoMParams
from libertem_b200.udf.base import UDFMeta, UDFData, UDF
from libertem_b200.udf import com as com_mod
from libertem_b200.runner import _get_dtype, UDFRunner


# ---- C ABI ------------------------------------------------------------------------------------

def declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'ltb200.h')).read()
    return sorted(set(re.findall(r'^LTB_API[^;(]*?\b(ltb200_\w+)\s*\(', hdr, flags=re.M)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'run __graft_entry__.build() first'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 9
    for name in syms:
        assert hasattr(lib, name), f'{name} declared in include/ltb200.h but not exported'
    # the python binding covers exactly the declared surface
    assert sorted(_lib.SIGNATURES) == syms
    assert _lib.get_lib().ltb200_abi_version() == 1


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, 'include', 'ltb200.h')).read()
    code = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)        # declarations only, no comments
    assert 'torch' not in code.lower() and 'at::' not in code and 'Tensor' not in code


def test_no_cpu_fallback():
    from libertem_b200 import engine
    with pytest.raises(_lib.LTB200Error):
        engine.masks_dense(torch.ones((4, 8)), torch.ones((2, 8)))
    # product code never imports the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'libertem_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f


# ---- Shape / Slice ----------------------------------------------------------------------------

def test_shape():
    s = Shape((5, 6, 16, 12), sig_dims=2)
    assert tuple(s.nav) == (5, 6) and tuple(s.sig) == (16, 12)
    assert s.size == 5 * 6 * 16 * 12 and s.nav.size == 30 and s.sig.size == 192
    assert tuple(s.flatten_nav()) == (30, 16, 12)
    assert tuple(s.flatten_sig()) == (5, 6, 192)
    assert pickle.loads(pickle.dumps(s)).sig_dims == 2


def test_slice_ops():
    sh = Shape((4, 16, 16), sig_dims=2)
    s = Slice(origin=(2, 0, 0), shape=sh)
    assert s.get() == (slice(2, 6), slice(0, 16), slice(0, 16))
    assert s.get(sig_only=True) == (slice(0, 16), slice(0, 16))
    sig = s.discard_nav()
    shifted = sig.shift_by((3, -5))
    inter = sig.intersection_with(shifted)
    assert inter.origin == (3, 0) and tuple(inter.shape) == (13, 11)
    assert sig.intersection_with(sig.shift_by((20, 0))).is_null()
    arr = np.arange(16 * 16).reshape(16, 16)
    assert np.array_equal(inter.get(arr), arr[3:16, 0:11])
    assert hash(sig) == hash(s.discard_nav()) and sig == s.discard_nav()


def test_partition_boundaries_match_reference_rule():
    for n, p in [(1024, 8), (15, 2), (63, 3), (7, 16), (65536, 8)]:
        assert partition_boundaries(n, p) == O.partition_boundaries(n, p)
    assert partition_boundaries(15, 2) == [(0, 7), (7, 15)]


# ---- MaskContainer (tests/common/test_mask_container.py in the reference) ---------------------

def test_mask_container_dense():
    stack = mixed_masks(16, 12, 3, 5)
    mc = MaskContainer(mask_factories=lambda: stack, dtype=np.float32)
    assert len(mc) == 3 and mc.use_sparse is False and mc.dtype == np.float32
    sl = full_sig_slice((16, 12))
    m = mc.get_for_sig_slice(sl)
    assert m.shape == (192, 3) and m.flags['F_CONTIGUOUS']     # container.py:86-91 layout
    assert np.array_equal(m, stack.reshape(3, -1).T)
    assert mc.get_for_sig_slice(sl) is m                        # cached
    sub = Slice(origin=(4, 0), shape=Shape((8, 12), sig_dims=2))
    ms = mc.get_for_sig_slice(sub, transpose=False)
    assert np.array_equal(ms, stack[:, 4:12, :].reshape(3, -1))
    # list of factories, count without computing
    mc2 = MaskContainer(mask_factories=[lambda: stack[0], lambda: stack[1]])
    assert len(mc2) == 2 and mc2._computed_masks is None
    assert mc2.computed_masks.shape == (2, 16, 12)
    # caches never travel (container.py:181-185)
    mc3 = pickle.loads(pickle.dumps(MaskContainer(mask_factories=_factory, dtype=np.float32)))
    assert mc3._computed_masks is None and mc3._slice_cache == {}


def _factory():
    return np.ones((2, 4, 4), dtype=np.float32)


def test_mask_container_sparse_modes():
    import scipy.sparse as sp
    dense = M.ring(8, 8, 16, 16, 6, 3)
    # use_sparse=None: sparse only if ALL factories return sparse (container.py:245-258)
    mc = MaskContainer(mask_factories=[lambda: sp.csr_matrix(dense), lambda: sp.coo_matrix(dense)])
    assert mc.use_sparse == 'scipy.sparse'
    mc = MaskContainer(mask_factories=[lambda: sp.csr_matrix(dense), lambda: dense])
    assert mc.use_sparse is False
    mc = MaskContainer(mask_factories=[lambda: dense], use_sparse=True, dtype=np.float32)
    assert mc.use_sparse == 'scipy.sparse'
    m = mc.get_for_sig_slice(full_sig_slice((16, 16)))
    assert sp.issparse(m) and m.format == 'csr' and m.shape == (256, 1)
    assert np.array_equal(m.toarray()[:, 0], dense.reshape(-1).astype(np.float32))
    mc = MaskContainer(mask_factories=[lambda: dense], use_sparse='scipy.sparse.csc')
    assert mc.get_for_sig_slice(full_sig_slice((16, 16))).format == 'csc'
    with pytest.raises(ValueError):
        MaskContainer(mask_factories=[lambda: dense], use_sparse='bogus')
    with pytest.raises(TypeError):
        mc.get((slice(0, 1),))


def test_radial_bins_partition_of_unity():
    rb = M.radial_bins(32, 32, 64, 64, n_bins=8, use_sparse=False, dtype=np.float64)
    assert np.allclose(rb.sum(axis=0), 1)          # tests/test_masks.py in the reference
    sparse = M.radial_bins(32, 32, 64, 64, n_bins=8, use_sparse=True, dtype=np.float64)
    assert M.is_sparse(sparse) and np.array_equal(sparse.todense(), rb)


# ---- dtype rules / buffers --------------------------------------------------------------------

def test_input_dtype_rule():
    assert _get_dtype([ApplyMasksUDF(mask_factories=_factory)], np.uint16) == np.float32
    assert _get_dtype([ApplyMasksUDF(mask_factories=_factory)], np.int32) == np.float64
    assert _get_dtype([SumUDF(), SumSigUDF()], np.uint8) == np.float32
    assert _get_dtype([SumUDF(dtype=np.float64)], np.float32) == np.float64
    assert _get_dtype([ApplyMasksUDF(mask_factories=_factory, preferred_dtype=np.int64)],
                      np.uint16) == np.int64


def _meta(shape, dtype=np.float32, roi=None):
    return UDFMeta(dataset_shape=Shape(shape, sig_dims=2), dataset_dtype=dtype,
                   input_dtype=dtype, roi=roi, device=torch.device('cpu'))


def test_result_buffer_declarations():
    udf = ApplyMasksUDF(mask_factories=_factory, mask_count=2, mask_dtype=np.complex64)
    udf.set_meta(_meta((3, 3, 4, 4)))
    b = udf.get_result_buffers()['intensity']
    assert (b.kind, b.extra_shape, b.dtype, b.where) == ('nav', (2,), np.complex64, 'device')
    com = CoMUDF()
    com.set_meta(_meta((3, 3, 4, 4)))
    decl = com.get_result_buffers()
    assert decl['raw_mask_result'].use == 'private' and decl['raw_mask_result'].extra_shape == (3,)
    assert decl['regression'].kind == 'single' and decl['regression'].dtype == np.float64
    assert all(decl[k].use == 'result_only' for k in ('raw_com', 'field', 'curl'))
    s = SumUDF()
    s.set_meta(_meta((3, 3, 4, 4), np.uint16))
    assert s.get_result_buffers()['intensity'].kind == 'sig'
    assert com.get_params() == CoMParams(cy=2, cx=2)
    with pytest.raises(ValueError):
        CoMUDF.with_params(r=3., ri=4.)


def test_buffer_wrapper_roi_view():
    roi = np.array([[True, False, True], [False, True, True]])
    b = BufferWrapper('nav', extra_shape=(2,), dtype=np.float32)
    b.set_shape_ds(Shape((2, 3, 4, 4), sig_dims=2), roi)
    b.allocate()
    assert b.raw_data.shape == (4, 2)
    b.tensor[:] = torch.arange(8, dtype=torch.float32).reshape(4, 2)
    d = b.data
    assert d.shape == (2, 3, 2) and np.isnan(d[0, 1]).all() and d[1, 2, 1] == 7
    bi = BufferWrapper('nav', dtype=np.int32)
    bi.set_shape_ds(Shape((2, 3, 4, 4), sig_dims=2), roi)
    bi.allocate()
    assert bi.data[0, 1] == 0


def test_default_merge_only_for_nav():
    class SigUDF(UDF):
        def get_result_buffers(self):
            return {'x': self.buffer(kind='sig', dtype=np.float32)}
    u = SigUDF()
    u.set_meta(_meta((2, 2, 4, 4)))
    with pytest.raises(NotImplementedError):
        u.merge(dest={}, src={})


def test_merge_all_contract():
    """reference udf/base.py:944-1002,1208-1224: default merge_all = concatenation of nav
    buffers in partition order; SumUDF.merge_all = stack-sum (udf/sum.py:54-58); sig buffers
    without a custom merge_all raise; unknown names raise ValueError"""
    import torch
    from collections import OrderedDict
    from libertem_b200.udf.base import MergeAttrMapping, UDFData
    from libertem_b200.udf import SumUDF, SumSigUDF

    def prepared(u, shape):
        u.set_meta(_meta(shape))
        decl = u.get_result_buffers()
        for b in decl.values():
            b.set_shape_ds(Shape(shape, sig_dims=2), None)
            b.allocate()
        u.results = UDFData(decl)
        return u

    u = prepared(SumSigUDF(), (2, 3, 4, 4))
    parts = OrderedDict([('p0', MergeAttrMapping({'intensity': torch.tensor([1., 2.])})),
                         ('p1', MergeAttrMapping({'intensity': torch.tensor([3., 4., 5., 6.])}))])
    u._do_merge_all(parts)
    assert u.results.get_buffer('intensity').raw_data.tolist() == [1, 2, 3, 4, 5, 6]

    s = prepared(SumUDF(), (2, 3, 4, 4))
    assert s.requires_custom_merge_all
    parts = OrderedDict([(i, MergeAttrMapping({'intensity': torch.full((16,), float(i + 1))}))
                         for i in range(3)])
    s._do_merge_all(parts)
    assert np.all(s.results.get_buffer('intensity').raw_data == 6.0)

    class SigUDF(UDF):
        def get_result_buffers(self):
            return {'x': self.buffer(kind='sig', dtype=np.float32)}
    g = prepared(SigUDF(), (2, 3, 4, 4))
    with pytest.raises(NotImplementedError):
        g._do_merge_all(OrderedDict([(0, MergeAttrMapping({'x': torch.zeros(16)}))]))

    class BadUDF(SigUDF):
        def merge_all(self, ordered_results):
            return {'nope': torch.zeros(16)}
    b = prepared(BadUDF(), (2, 3, 4, 4))
    with pytest.raises(ValueError):
        b._do_merge_all(OrderedDict([(0, MergeAttrMapping({'x': torch.zeros(16)}))]))


def test_apply_masks_argument_errors():
    with pytest.raises(ValueError):
        ApplyMasksUDF(mask_factories=_factory, backends=('nonsense',))
    with pytest.raises(ValueError):
        ApplyMasksUDF(mask_factories=_factory, shifts=(1, 2), use_sparse='scipy.sparse')
    assert ApplyMasksUDF(mask_factories=_factory, shifts=(1, 2)).get_method() == 'frame'
    assert ApplyMasksUDF(mask_factories=_factory).get_method() == 'tile'


# ---- CoM post-processing on the host, against the reference's golden outputs --------------------

def _com_from_raw(raw, nav_shape, sig_shape, roi=None, **params):
    udf = CoMUDF.with_params(**params)
    udf.set_meta(_meta(tuple(nav_shape) + tuple(sig_shape), roi=roi))
    decl = udf.get_result_buffers()
    for b in decl.values():
        b.set_shape_ds(udf.meta.dataset_shape, roi)
    decl['raw_mask_result'].replace_array(torch.from_numpy(np.ascontiguousarray(raw)))
    udf.results = UDFData(decl)
    return udf.get_results()


@pytest.mark.parametrize('i', [0, 1, 2])
def test_com_get_results_matches_reference(i):
    meta, g = load_golden(f'com_params_{i}')
    c = meta['com']
    if c['r'] == 'inf':
        c['r'] = float('inf')
    res = _com_from_raw(g['raw_mask_result'], meta['shape'][:2], meta['shape'][2:], **c)
    for k in ('raw_com', 'raw_shifts', 'field', 'field_y', 'field_x', 'magnitude', 'divergence',
              'curl', 'regression'):
        assert res[k].dtype == g[k].dtype and res[k].shape == g[k].shape, k
        np.testing.assert_allclose(res[k], g[k], rtol=1e-6, atol=1e-6, err_msg=k)


def test_com_get_results_roi_matches_reference():
    from golden_inputs import roi_from_seed
    meta, g = load_golden('roi')
    roi = roi_from_seed(meta['shape'][:2], meta['roi_seed'])
    res = _com_from_raw(g['com_raw_mask_result'], meta['shape'][:2], meta['shape'][2:], roi=roi,
                        regression=1)
    for k in ('raw_com', 'raw_shifts', 'field', 'magnitude', 'divergence', 'curl', 'regression'):
        np.testing.assert_allclose(res[k], g['com_' + k], rtol=1e-6, atol=1e-6, err_msg=k,
                                   equal_nan=True)


def test_guess_corrections():
    # synthetic rotated curl-free field: grad of a radial potential, rotated by 30 deg + flip
    y, x = np.mgrid[-16:16, -16:16].astype(np.float64)
    pot = np.exp(-(x ** 2 + y ** 2) / 60.)
    fy, fx = np.gradient(pot)
    ry, rx = com_mod.apply_correction(fy, fx, scan_rotation=30., flip_y=True, forward=False)
    g = com_mod.guess_corrections(ry + 0.25, rx - 0.5)
    assert g.flip_y is True and abs(((g.scan_rotation - 30 + 90) % 180) - 90) <= 1
    assert abs(g.cy - 0.25) < 0.05 and abs(g.cx + 0.5) < 0.05
    ref_y, ref_x = O.apply_correction(fy, fx, 30., True, forward=False)
    assert np.allclose(ry, ref_y) and np.allclose(rx, ref_x)


# ---- dataset tiling (host side only: CPU tensors) -----------------------------------------------

def test_memory_dataset_partitions_and_roi_tiles():
    data = torch.arange(6 * 5 * 4 * 4, dtype=torch.float32).reshape(6, 5, 4, 4)
    ds = MemoryDataSet(data=data, num_partitions=4, sig_dims=2)
    parts = list(ds.get_partitions())
    assert [(p.start, p.stop) for p in parts] == O.partition_boundaries(30, 4)
    assert parts[1].slice.origin == (parts[1].start, 0, 0)
    assert UDFRunner.my_partitions(parts, 1, 2) == parts[2:4]
    # weighted shares: contiguous, disjoint, complete
    blocks = [UDFRunner.my_partitions(parts, r, 3, weights=[1, 2, 1]) for r in range(3)]
    assert sum(blocks, []) == parts and [len(b) for b in blocks] == [1, 2, 1]
    # sub-frame tiles in the reference's order: depth blocks outer, sig slices inner
    ds2 = MemoryDataSet(data=data, num_partitions=1, sig_dims=2, tileshape=(8, 2, 4))
    seen = [(f0, f1, t.origin, tuple(t.shape)) for _, f0, f1, t in
            _cpu_tiles(ds2)]
    want = [(f0, f1, (f0,) + tuple(s.start for s in sl), (f1 - f0,) +
             tuple(s.stop - s.start for s in sl)) for f0, f1, sl in O.iter_tiles(0, 30, (4, 4), (8, 2, 4))]
    assert seen == want


def _cpu_tiles(ds):
    # the device-resident branch is pure indexing; drive it with CPU tensors by faking is_cuda
    part = list(ds.get_partitions())[0]
    flat = ds._flat()
    depth = ds.tileshape[0]
    for f0 in range(part.start, part.stop, depth):
        f1 = min(f0 + depth, part.stop)
        yield from ds._emit(flat[f0:f1], f0, f1, None, None, ds.tileshape)


def test_banded_quad_plan_reconstructs_the_masks():
    """host plan of the tensor-core group-sparse kernel (group_masks.build_banded): quads are 4
    consecutive pixels at a multiple of 4 inside their band, groups are (band, ring) pairs
    padded to 64 entries, and hi + lo of the split weight table reproduce every mask value"""
    from libertem_b200 import group_masks as gm
    rng = np.random.default_rng(0)
    K, size, n_rings, n_bands = 4096, 5, 3, 4
    stack = np.zeros((n_rings * size, K), np.complex64)
    for g in range(n_rings):
        px = np.sort(rng.choice(K, 300, replace=False))
        stack[g * size:(g + 1) * size, px] = (rng.random((size, 300)) - 0.5
                                              + 1j * rng.random((size, 300))).astype(np.complex64)
    n_cols = 32                                   # 4 * size = 20 -> 32 accumulator columns
    b = gm.build_banded(stack, size, n_bands, n_cols)
    off, quad_px, table = b['group_off'], b['entry_px'], b['table_split']
    assert b['n_groups'] == n_bands * n_rings and len(off) == b['n_groups'] + 1
    assert np.all(off % gm.TC_KT == 0) and off[-1] == 4 * len(quad_px) == table.shape[1]
    assert np.all(quad_px % 4 == 0)
    nq = n_cols // 4
    dense = np.zeros_like(stack, dtype=np.complex128)
    for gidx in range(b['n_groups']):
        band, ring = divmod(gidx, n_rings)
        for e in range(off[gidx], off[gidx + 1]):
            px = quad_px[e // 4] + e % 4
            in_band = band * K // n_bands <= px < (band + 1) * K // n_bands
            assert in_band or not table[:, e].any()
            for r in range(2 * size):
                h, j = divmod(r, nq)
                val = float(table[h * 2 * nq + j, e]) + float(table[h * 2 * nq + nq + j, e])
                dense[ring * size + r // 2, px] += val * (1j if r % 2 else 1.0)
    assert np.abs(dense - stack).max() <= 2.0 ** -21
    assert gm.default_bands(512 * 512) == 4 and gm.default_bands(64 * 64) == 1
    assert gm.build_banded(stack[:, :4090], size, 4, n_cols) is None     # K % 16 != 0


@pytest.mark.parametrize('seed,K,size,n_rings,n_bands', [(1, 1024, 3, 2, 1), (2, 2048, 7, 4, 8),
                                                        (3, 960, 1, 3, 5), (4, 4096, 25, 2, 16)])
def test_banded_quad_plan_random(seed, K, size, n_rings, n_bands):
    """random supports (incl. an empty ring and rings confined to a few bands): every support
    pixel appears exactly once with its weight, padding carries weight zero"""
    from libertem_b200 import group_masks as gm
    rng = np.random.default_rng(seed)
    stack = np.zeros((n_rings * size, K), np.complex64)
    for g in range(n_rings):
        if g == 1:
            continue                                   # an empty ring
        lo = int(rng.integers(0, K // 2))
        px = np.sort(rng.choice(np.arange(lo, min(K, lo + K // 3)), 150, replace=False))
        stack[g * size:(g + 1) * size, px] = (rng.random((size, 150)) + 0.1
                                              + 1j * rng.random((size, 150))).astype(np.complex64)
    n_cols = _lib.get_lib().ltb200_group_masks_tc_columns(size)
    b = gm.build_banded(stack, size, n_bands, n_cols)
    assert b['n_groups'] == n_bands * n_rings
    x = rng.random((2, K)).astype(np.float32)
    out = _emulate_quad_plan(b, x, n_rings, size, n_cols)
    ref = x.astype(np.float64) @ stack.astype(np.complex128).T
    assert np.abs(out - ref).max() <= 2.0 ** -20 * np.abs(ref).max()
    # every (ring, pixel) of the support is listed exactly once with a non-zero weight
    off, quad_px, table = b['group_off'], b['entry_px'], b['table_split']
    for ring in range(n_rings):
        seen = []
        for band in range(n_bands):
            gidx = band * n_rings + ring
            e = np.arange(off[gidx], off[gidx + 1])
            px = quad_px[e // 4] + e % 4
            seen.append(px[np.any(table[:, e] != 0, axis=0)])
        seen = np.concatenate(seen) if seen else np.zeros(0, int)
        want = np.nonzero(np.any(stack[ring * size:(ring + 1) * size] != 0, axis=0))[0]
        assert np.array_equal(np.sort(seen), want)


def test_int8_digit_plan():
    """host side of the integer fast path: which mask stacks K8 takes, how wide integer weights
    are split into base-128 int8 digits and how float weights become 28-bit fixed point"""
    from libertem_b200.runner import int8_digit_plan
    K = 64
    binary = torch.zeros((3, K))
    binary[0] = 1
    binary[1, ::3] = 1
    binary[2, 5:9] = -127
    i8, comb = int8_digit_plan(binary)
    assert comb is None and i8.dtype == torch.int8
    assert torch.equal(i8.float(), binary)
    # CoM coordinate rows of a 256-wide detector: 0..255 needs a second digit
    grad = torch.arange(K, dtype=torch.float32).repeat(2, 1) * 4.0          # 0 .. 252
    grad[1] = -grad[1] - 16000 + 3                                          # down to -16249
    stack = torch.cat([binary[:1], grad])
    i8, comb = int8_digit_plan(stack)
    assert i8.shape == (5, K) and comb.shape == (3, 5)
    assert int(i8.abs().max()) <= 127
    assert torch.equal((comb @ i8.double()).float(), stack)                 # exact
    # integer weights beyond two digits / too many rows: not representable as integer digits
    assert int8_digit_plan(stack * 2, float_digits=0) is None
    assert int8_digit_plan(stack * 0.5 + 0.25, float_digits=0) is None
    assert int8_digit_plan(grad[:1].repeat(9, 1)) is None                   # 9 + 9 rows > 16
    assert int8_digit_plan(grad[:1].repeat(8, 1)) is not None
    # float weights: four balanced base-128 digits of round(m 2^(27 - e)); |m - s q| <= 2^-27 max|m|
    rng = np.random.default_rng(7)
    fl = torch.from_numpy(np.stack([
        rng.uniform(0, 1, K), rng.normal(0, 3e-4, K), rng.uniform(-5e6, 5e6, K),
        np.exp(rng.uniform(-30, 3, K))]).astype(np.float32))
    both = torch.cat([binary[:1], fl, grad[:1]])
    i8, comb = int8_digit_plan(both, max_rows=32)
    assert i8.shape[0] == 1 + 4 * 4 + 2 and int(i8.abs().max()) <= 127
    rebuilt = comb @ i8.double()
    err = (rebuilt - both.double()).abs().amax(dim=1)
    bound = both.abs().amax(dim=1).double() * 2.0 ** -27
    assert bool((err <= bound).all()), (err, bound)
    assert torch.equal(rebuilt[0].float(), both[0]) and torch.equal(rebuilt[5].float(), both[5])
    assert int8_digit_plan(fl.repeat(3, 1), max_rows=32) is None            # 48 rows > 32
    z = int8_digit_plan(torch.zeros((2, K)))
    assert z[1] is None and not z[0].any()


def _emulate_quad_plan(b, data, n_rings, size, n_cols):
    """what K7 computes from a banded quad plan (gather 4 px per quad, weights = hi + lo)"""
    off, quad_px, table = b['group_off'], b['entry_px'], b['table_split']
    nq = n_cols // 4
    out = np.zeros((data.shape[0], n_rings * size), dtype=np.complex128)
    for gidx in range(b['n_groups']):
        ring = gidx % n_rings
        for e in range(off[gidx], off[gidx + 1]):
            x = data[:, quad_px[e // 4] + e % 4].astype(np.float64)
            for r in range(2 * size):
                h, j = divmod(r, nq)
                w = float(table[h * 2 * nq + j, e]) + float(table[h * 2 * nq + nq + j, e])
                out[:, ring * size + r // 2] += x * w * (1j if r % 2 else 1.0)
    return out


def test_mirror_symmetric_plan_emulation():
    """host plan of the opt-in mirror-symmetric kernel (group_masks.build_sym), emulated
    stage by stage exactly as the kernel walks it: 8 upper quads + their 8 mirror images per
    stage, float32 butterflies, real weights on the sums and imaginary weights on the
    differences, one table column per orbit; plus the rows without a partner through the
    ordinary quad plan.  Must reproduce the direct sum over the reference's radial masks."""
    from libertem_b200 import group_masks as gm
    from libertem_b200.analysis.radialfourier import radial_mask_factory
    S, n_bins, max_order = 32, 4, 6
    size = max_order + 1
    ro = M.bounding_radius(S / 2, S / 2, S, S)
    stack = np.asarray(radial_mask_factory(S, S, S / 2, S / 2, 0, ro, n_bins, max_order,
                                           use_sparse=False)()).astype(np.complex64)
    assert stack.shape == (n_bins * size, S, S)
    b = gm.build_sym(stack, size, n_bands=3)
    assert b is not None and b['n_groups'] == 3 * n_bins
    off, quads, T = b['group_off'], b['entry_px'], b['table_sym']
    assert np.all(off % gm.TC_KT == 0) and off[-1] == 4 * len(quads) and T.shape == (128, off[-1] // 2)
    assert np.all(quads % 4 == 0)
    F = 5
    data = np.random.default_rng(3).random((F, S * S), dtype=np.float32)
    out = np.zeros((F, n_bins * size), dtype=np.complex128)
    for gidx in range(b['n_groups']):
        ring = gidx % n_bins
        for k in range((off[gidx + 1] - off[gidx]) // 64):
            q0 = off[gidx] // 4 + 16 * k
            upper, lower = quads[q0:q0 + 8], quads[q0 + 8:q0 + 16]
            assert np.all(lower // S == S - upper // S) and np.all(lower % S == upper % S)
            px_u = (upper[:, None] + np.arange(4)[None, :]).reshape(-1)
            px_l = (lower[:, None] + np.arange(4)[None, :]).reshape(-1)
            s = (data[:, px_u] + data[:, px_l]).astype(np.float64)        # float32 butterflies
            d = (data[:, px_u] - data[:, px_l]).astype(np.float64)
            col0 = off[gidx] // 2 + 32 * k
            t = T[:, col0:col0 + 32].astype(np.float64)
            for c in range(size):
                out[:, ring * size + c] += s @ (t[c] + t[32 + c]) + 1j * (d @ (t[64 + c] + t[96 + c]))
    rest = gm.build_banded(np.where(b['residual'].reshape(-1)[None, :], stack.reshape(-1, S * S), 0),
                           size, 1, 32)
    out += _emulate_quad_plan(rest, data, n_bins, size, 32)
    ref = data.astype(np.float64) @ stack.reshape(-1, S * S).astype(np.complex128).T
    scale = (np.abs(data).astype(np.float64) @ np.abs(stack.reshape(-1, S * S)).astype(np.float64).T).max()
    assert np.abs(out - ref).max() / scale <= 2e-6
    # a stack without the symmetry has no such plan
    broken = stack.copy()
    broken[3, 5, 7] += 0.5
    assert gm.build_sym(broken, size, 3) is None
    assert gm.build_sym(stack[:, :31, :], size, 3) is None
