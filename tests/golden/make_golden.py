"""Generate golden vectors by running the UNMODIFIED reference (LiberTEM) on CPU.

Run in the build container only (needs /root/reference):

    PYTHONPATH=tests/golden/shims:/root/reference/src python tests/golden/make_golden.py

Inputs are regenerated from the counter-based generator in oracle/synth.py (seeded), so only
the reference's *outputs* are stored (small .npz files under tests/golden/).  Third-party
packages missing from this image (sparseconverter, sparse, matplotlib, ...) are replaced by the
stand-ins in tests/golden/shims; no reference code is modified or copied.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))

from oracle import synth  # noqa: E402

from libertem.udf.masks import ApplyMasksUDF  # noqa: E402
from libertem.udf.com import CoMUDF  # noqa: E402
from libertem.udf.sum import SumUDF  # noqa: E402
from libertem.udf.sumsigudf import SumSigUDF  # noqa: E402
from libertem.udf.base import UDFRunner  # noqa: E402
from libertem.executor.inline import InlineJobExecutor  # noqa: E402
from libertem.io.dataset.memory import MemoryDataSet  # noqa: E402
from libertem import masks as M  # noqa: E402
from libertem.analysis.radialfourier import RadialFourierAnalysis  # noqa: E402
from libertem.common.numba import rmatmul  # noqa: E402
import scipy.sparse as sp  # noqa: E402


def run(ds_kwargs, udfs, roi=None):
    ex = InlineJobExecutor()
    ds = MemoryDataSet(**ds_kwargs)
    ds.initialize(ex)
    res = UDFRunner(udfs).run_for_dataset(ds, ex, roi=roi)
    return ds, res.buffers


def com_raw(ds_kwargs, cy=None, cx=None, r=float('inf'), ri=0., roi=None):
    """CoMUDF's private 'raw_mask_result' buffer, re-derived through the reference's own
    com mask factories + ApplyMasksUDF (the COMAnalysis formulation, analysis/com.py:286-334)."""
    from libertem.udf.com import com_masks_factory, com_masks_generic
    sy, sx = ds_kwargs['data'].shape[-2:]
    cy = sy // 2 if cy is None else cy
    cx = sx // 2 if cx is None else cx
    if ri is None or np.isclose(ri, 0.):
        fac = com_masks_factory(detector_y=sy, detector_x=sx, cy=cy, cx=cx, r=r)
    else:
        fac = com_masks_generic(sy, sx, lambda: M.ring(
            imageSizeY=sy, imageSizeX=sx, centerY=cy, centerX=cx, radius=r, radius_inner=ri))
    _, bufs = run(ds_kwargs, [ApplyMasksUDF(mask_factories=fac, mask_count=3,
                                            mask_dtype=np.float32, use_sparse=False)], roi=roi)
    return bufs[0]['intensity'].raw_data


COM_KEYS = ('raw_com', 'raw_shifts', 'field', 'field_y', 'field_x', 'magnitude', 'divergence',
            'curl', 'regression')


def mixed_masks(sy, sx, count, seed):
    """The dense mask mix used by cfg2/cfg5 (also built by libertem_b200.bench_masks)."""
    out = []
    cy, cx = sy // 2, sx // 2
    for i in range(count):
        kind = i % 4
        if kind == 0:
            m = synth.uniform_f32(0, sy * sx, seed + i).reshape(sy, sx)
        elif kind == 1:
            m = M.circular(cx, cy, sx, sy, radius=min(sy, sx) / 4 + i).astype(np.float32)
        elif kind == 2:
            m = M.ring(cx, cy, sx, sy, radius=min(sy, sx) / 3 + i,
                       radius_inner=min(sy, sx) / 6).astype(np.float32)
        else:
            m = (M.gradient_x(sx, sy) - cx) * 0.5 + (M.gradient_y(sx, sy) - cy) * 0.25
            m = m.astype(np.float32)
        out.append(m)
    return np.stack(out)


def save(name, meta, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, meta=json.dumps(meta), **arrays)
    print(name, {k: (v.shape, str(v.dtype)) for k, v in arrays.items()})


def g_cfg1():
    # BASELINE.json configs[0]: 32x32 nav x 64x64 sig f32, 1 dense mask, inline executor
    shape = (32, 32, 64, 64)
    data = synth.dataset(shape, np.float32, seed=101)
    mask = synth.uniform_f32(0, 64 * 64, 201).reshape(64, 64)
    for nparts in (1, 8):
        _, bufs = run(dict(data=data, num_partitions=nparts, sig_dims=2),
                      [ApplyMasksUDF(mask_factories=[lambda: mask])])
        save(f'cfg1_p{nparts}', dict(shape=shape, data_seed=101, mask_seed=201,
                                     num_partitions=nparts),
             intensity=bufs[0]['intensity'].raw_data)


def g_cfg2_small():
    # cfg2 scaled: 16x16 nav x 256x256 sig f32, 8 dense masks + CoM (+SumSig, Sum fused)
    shape = (16, 16, 256, 256)
    data = synth.dataset(shape, np.float32, seed=102)
    stack = mixed_masks(256, 256, 8, seed=202)
    _, bufs = run(dict(data=data, num_partitions=4, sig_dims=2),
                  [ApplyMasksUDF(mask_factories=lambda: stack), CoMUDF(), SumUDF(), SumSigUDF()])
    arrays = dict(intensity=bufs[0]['intensity'].raw_data,
                  sum=bufs[2]['intensity'].raw_data, sumsig=bufs[3]['intensity'].raw_data)
    for k in COM_KEYS:
        arrays['com_' + k] = bufs[1][k].raw_data
    arrays['com_raw_mask_result'] = com_raw(dict(data=data, num_partitions=4, sig_dims=2))
    save('cfg2_small', dict(shape=shape, data_seed=102, mask_seed=202, n_masks=8,
                            num_partitions=4), **arrays)


def g_cfg3_small():
    # cfg3 scaled: 16x16 nav x 128x128 sig u16 Poisson(3), Sum + SumSig + 4 sparse ring masks
    shape = (16, 16, 128, 128)
    data = synth.dataset(shape, np.uint16, seed=103)
    rings = [(8, 16), (20, 28), (32, 40), (44, 52)]
    facs = [lambda ri=ri, ro=ro: M.ring(64, 64, 128, 128, ro, ri) for ri, ro in rings]
    ds, bufs = run(dict(data=data, num_partitions=4, sig_dims=2),
                   [SumUDF(), SumSigUDF(),
                    ApplyMasksUDF(mask_factories=facs, use_sparse=True, mask_dtype=np.float32)])
    save('cfg3_small', dict(shape=shape, data_seed=103, rings=rings, num_partitions=4),
         sum=bufs[0]['intensity'].raw_data, sumsig=bufs[1]['intensity'].raw_data,
         intensity=bufs[2]['intensity'].raw_data)


def g_cfg4_small():
    # cfg4 scaled: 6x6 nav x 64x64 sig f32, radial Fourier 8 bins, default max_order=24
    shape = (6, 6, 64, 64)
    data = synth.dataset(shape, np.float32, seed=104)
    ex = InlineJobExecutor()
    for use_sparse in (True, False):
        ds = MemoryDataSet(data=data, num_partitions=2, sig_dims=2)
        ds.initialize(ex)
        params = {'n_bins': 8}
        params['use_sparse'] = 'scipy.sparse' if use_sparse else False
        a = RadialFourierAnalysis(dataset=ds, parameters=params)
        udf = a.get_udf()
        res = UDFRunner([udf]).run_for_dataset(ds, ex)
        inten = res.buffers[0]['intensity'].data
        p = a.parameters
        raw = inten.reshape((36, -1)).T.reshape((p['n_bins'], p['max_order'] + 1, 6, 6))
        stack = np.asarray(udf.masks.computed_masks) if not use_sparse else None
        meta = dict(shape=shape, data_seed=104,
                    params={k: (v if not isinstance(v, (np.generic,)) else v.item())
                            for k, v in p.items() if k not in ('mask_dtype',)},
                    num_partitions=2)
        meta['params']['use_sparse'] = bool(use_sparse)
        arrays = dict(raw_results=raw)
        if stack is not None:
            arrays['mask_stack_sub'] = stack[::37]   # sample of the complex mask stack
        save('cfg4_small_' + ('sparse' if use_sparse else 'dense'), meta, **arrays)


def g_cfg4_k7():
    # RadialFourierAnalysis at a size where this repo's DEFAULT kernel (K7, tcgen05 group-sparse,
    # >= 96 frames per tile) runs: 16x16 nav x 128x128 sig f32, 8 bins, max_order 24, ONE
    # partition (256 frames).  'centre': default cx/cy (mirror-symmetric plan applies);
    # 'offcentre': cx/cy off the pixel grid with an inner radius (banded plan only).
    shape = (16, 16, 128, 128)
    data = synth.dataset(shape, np.float32, seed=114)
    ex = InlineJobExecutor()
    cases = {'centre': {'n_bins': 8},
             'offcentre': {'n_bins': 5, 'cx': 60.5, 'cy': 70.25, 'ri': 6., 'ro': 50.,
                           'max_order': 12}}
    for name, params in cases.items():
        ds = MemoryDataSet(data=data, num_partitions=1, sig_dims=2)
        ds.initialize(ex)
        a = RadialFourierAnalysis(dataset=ds, parameters=dict(params))
        udf = a.get_udf()
        res = UDFRunner([udf]).run_for_dataset(ds, ex)
        inten = res.buffers[0]['intensity'].data
        p = a.parameters
        raw = inten.reshape((256, -1)).T.reshape((p['n_bins'], p['max_order'] + 1, 16, 16))
        meta = dict(shape=shape, data_seed=114, num_partitions=1,
                    params={k: (v if not isinstance(v, (np.generic,)) else v.item())
                            for k, v in p.items() if k not in ('mask_dtype', 'use_sparse')},
                    call={k: v for k, v in params.items()})
        save('cfg4_k7_' + name, meta, raw_results=raw.astype(np.complex64))


def g_radial_symmetries():
    # the reference's known-answer test tests/analysis/test_analysis_radialfourier.py:78-188:
    # four CBED frames with 1- / 2- / 4-fold symmetric spot arrangements; here the 2x2 scan is
    # tiled to 16x16 (256 frames, one partition) so that the run goes through K7.  Inputs come
    # from the reference's generator (libertem.utils.generate.cbed_frame) and are stored (they
    # are mostly zeros and compress to a few KB).
    from libertem.utils.generate import cbed_frame
    (d1, i1, p1) = cbed_frame(all_equal=True, radius=3, indices=np.array([(1, 0)]))
    (d2, i2, p2) = cbed_frame(all_equal=True, radius=3, indices=np.array([(-1, 0)]))
    (d3, i3, p3) = cbed_frame(all_equal=True, radius=3, indices=np.array([(1, 0), (-1, 0)]))
    (d4, i4, p4) = cbed_frame(
        all_equal=True, radius=3, indices=np.array([(1, 0), (-1, 0), (0, 1), (0, -1)]))
    frames = np.stack([d1[0], d2[0], d3[0], d4[0]]).astype(np.float32)
    data = np.zeros((16, 16) + frames.shape[1:], dtype=np.float32)
    for i in range(16):
        for j in range(16):
            data[i, j] = frames[(i % 2) * 2 + (j % 2)]
    r = np.linalg.norm(p2[0] - p1[0]) / 2
    cy, cx = (p2[0] + p1[0]) / 2
    ex = InlineJobExecutor()
    ds = MemoryDataSet(data=data, num_partitions=1, sig_dims=2)
    ds.initialize(ex)
    params = dict(cy=float(cy), cx=float(cx), ri=0, ro=float(r + 4), n_bins=2, max_order=8)
    a = RadialFourierAnalysis(dataset=ds, parameters=dict(params))
    res = UDFRunner([a.get_udf()]).run_for_dataset(ds, ex)
    inten = res.buffers[0]['intensity'].data
    p = a.parameters
    raw = inten.reshape((256, -1)).T.reshape((p['n_bins'], p['max_order'] + 1, 16, 16))
    save('radial_symmetries', dict(call=params, num_partitions=1, shape=list(data.shape)),
         frames=frames, raw_results=raw.astype(np.complex64),
         frame_sums=frames.sum(axis=(1, 2), dtype=np.float64))


def g_complex_input():
    # complex64 frames (reference udf/masks.py:360-368; tests/analysis/test_analysis_masks.py:
    # 151-212, tests/analysis/test_analysis_com.py:132-231): real masks, complex masks, and the
    # COMAnalysis formulation (3 CoM masks + center_shifts / apply_correction of udf/com.py)
    from libertem.udf.com import center_shifts, apply_correction
    shape = (6, 7, 24, 32)
    data = (synth.dataset(shape, np.float32, seed=301)
            + 1j * (synth.dataset(shape, np.float32, seed=302) - 0.5)).astype(np.complex64)
    real_masks = mixed_masks(24, 32, 3, 31)
    cmasks = (mixed_masks(24, 32, 2, 41) + 1j * mixed_masks(24, 32, 2, 51)).astype(np.complex64)
    out = {}
    for name, kw in (('p2', dict(num_partitions=2)),
                     ('tiled', dict(num_partitions=3, tileshape=(5, 8, 32)))):
        dskw = dict(data=data, sig_dims=2, **kw)
        _, b = run(dskw, [ApplyMasksUDF(mask_factories=lambda: real_masks, mask_count=3,
                                        mask_dtype=np.float32, use_sparse=False)])
        out['real_masks_' + name] = b[0]['intensity'].raw_data
        _, b = run(dskw, [ApplyMasksUDF(mask_factories=lambda: cmasks, mask_count=2,
                                        mask_dtype=np.complex64, use_sparse=False)])
        out['complex_masks_' + name] = b[0]['intensity'].raw_data
    raw = com_raw(dict(data=data, num_partitions=2, sig_dims=2), cy=0, cx=0)
    out['com_raw'] = raw
    img = raw.reshape(6, 7, 3)
    y, x = center_shifts(img[..., 0], img[..., 1], img[..., 2], 0, 0)
    y, x = apply_correction(y, x, scan_rotation=0., flip_y=False)
    out['com_x'] = np.asarray(x)
    out['com_y'] = np.asarray(y)
    save('complex_input', dict(shape=shape, seeds=[301, 302], mask_seeds=[31, 41, 51]), **out)


def g_com_params():
    # CoM with disk/ring, rotation, flip, regression on a non-square nav/sig
    shape = (12, 10, 32, 40)
    data = synth.dataset(shape, np.float32, seed=105)
    cases = [
        dict(cy=15.2, cx=21.7, r=11.5, ri=0., scan_rotation=33., flip_y=True, regression=1),
        dict(cy=14., cx=18., r=13., ri=4., scan_rotation=-70., flip_y=False, regression=0),
        dict(cy=None, cx=None, r=float('inf'), ri=0., scan_rotation=0., flip_y=False,
             regression=-1),
    ]
    for i, c in enumerate(cases):
        _, bufs = run(dict(data=data, num_partitions=3, sig_dims=2),
                      [CoMUDF.with_params(**c)])
        arrays = {k: bufs[0][k].raw_data for k in COM_KEYS}
        arrays['raw_mask_result'] = com_raw(dict(data=data, num_partitions=3, sig_dims=2),
                                            cy=c['cy'], cx=c['cx'], r=c['r'], ri=c['ri'])
        c = dict(c)
        if c['r'] == float('inf'):
            c['r'] = 'inf'
        save(f'com_params_{i}', dict(shape=shape, data_seed=105, com=c, num_partitions=3),
             **arrays)


def g_roi():
    shape = (7, 9, 24, 20)
    data = synth.dataset(shape, np.float32, seed=106)
    roi = (synth.hash_u32(0, 63, 306) % 3 != 0).reshape(7, 9)
    stack = mixed_masks(24, 20, 3, seed=206)
    _, bufs = run(dict(data=data, num_partitions=3, sig_dims=2),
                  [ApplyMasksUDF(mask_factories=lambda: stack),
                   CoMUDF.with_params(regression=1), SumUDF(), SumSigUDF()], roi=roi)
    arrays = dict(intensity=bufs[0]['intensity'].raw_data, sum=bufs[2]['intensity'].raw_data,
                  sumsig=bufs[3]['intensity'].raw_data,
                  intensity_full=bufs[0]['intensity'].data)
    for k in COM_KEYS:
        arrays['com_' + k] = bufs[1][k].raw_data
    arrays['com_raw_mask_result'] = com_raw(dict(data=data, num_partitions=3, sig_dims=2),
                                            roi=roi)
    arrays['com_divergence_full'] = bufs[1]['divergence'].data
    save('roi', dict(shape=shape, data_seed=106, mask_seed=206, roi_seed=306, n_masks=3,
                     num_partitions=3), **arrays)


def g_odd():
    # ragged shapes / sub-frame tiles / u16 decode path (tests/udf/test_sum.py tileshape (8,17,23))
    shape = (3, 5, 17, 23)
    stack = mixed_masks(17, 23, 5, seed=207)
    for dt, seed in ((np.float32, 107), (np.uint16, 108)):
        data = synth.dataset(shape, dt, seed=seed)
        ds, bufs = run(dict(data=data, num_partitions=2, sig_dims=2, tileshape=(4, 17, 23)),
                       [ApplyMasksUDF(mask_factories=lambda: stack), CoMUDF(), SumUDF(),
                        SumSigUDF()])
        save('odd_' + np.dtype(dt).name,
             dict(shape=shape, data_seed=seed, mask_seed=207, n_masks=5, num_partitions=2),
             intensity=bufs[0]['intensity'].raw_data,
             com_raw_mask_result=com_raw(dict(data=data, num_partitions=2, sig_dims=2,
                                              tileshape=(4, 17, 23))),
             com_raw_com=bufs[1]['raw_com'].raw_data,
             sum=bufs[2]['intensity'].raw_data, sumsig=bufs[3]['intensity'].raw_data)
    # forced sub-frame tiling of dense data (tests/udf/test_multi_udf.py:12-42 knobs)
    shape = (4, 4, 30, 16)
    data = synth.dataset(shape, np.float32, seed=109)
    stack2 = mixed_masks(30, 16, 3, seed=209)
    _, bufs = run(dict(data=data, num_partitions=2, sig_dims=2, base_shape=(1, 10, 16),
                       force_need_decode=True, tileshape=(4, 10, 16)),
                  [ApplyMasksUDF(mask_factories=lambda: stack2), SumUDF(), SumSigUDF()])
    save('subframe', dict(shape=shape, data_seed=109, mask_seed=209, n_masks=3,
                          num_partitions=2, tileshape=(4, 10, 16)),
         intensity=bufs[0]['intensity'].raw_data, sum=bufs[1]['intensity'].raw_data,
         sumsig=bufs[2]['intensity'].raw_data)


def g_dtypes():
    # dtype rules: int32 data -> float64 compute; float64 masks on float32 data; complex masks
    shape = (2, 3, 8, 8)
    out = {}
    stack32 = mixed_masks(8, 8, 2, seed=210)
    d_i32 = (synth.hash_u32(0, 2 * 3 * 64, 110) % 1000).astype(np.int32).reshape(shape)
    _, bufs = run(dict(data=d_i32, num_partitions=2, sig_dims=2),
                  [ApplyMasksUDF(mask_factories=lambda: stack32)])
    out['i32_f32'] = bufs[0]['intensity'].raw_data
    d_f32 = synth.dataset(shape, np.float32, seed=111)
    stack64 = stack32.astype(np.float64) * 1.000000123
    _, bufs = run(dict(data=d_f32, num_partitions=2, sig_dims=2),
                  [ApplyMasksUDF(mask_factories=lambda: stack64)])
    out['f32_f64'] = bufs[0]['intensity'].raw_data
    _, bufs = run(dict(data=d_f32, num_partitions=2, sig_dims=2),
                  [ApplyMasksUDF(mask_factories=lambda: stack64, mask_dtype=np.float32)])
    out['f32_f64_forced32'] = bufs[0]['intensity'].raw_data
    stackc = (stack32 + 1j * stack32[::-1]).astype(np.complex64)
    _, bufs = run(dict(data=d_f32, num_partitions=2, sig_dims=2),
                  [ApplyMasksUDF(mask_factories=lambda: stackc)])
    out['f32_c64'] = bufs[0]['intensity'].raw_data
    d_u8 = (synth.hash_u32(0, 2 * 3 * 64, 112) % 256).astype(np.uint8).reshape(shape)
    _, bufs = run(dict(data=d_u8, num_partitions=2, sig_dims=2),
                  [ApplyMasksUDF(mask_factories=lambda: stack32), SumUDF(), SumSigUDF()])
    out['u8_f32'] = bufs[0]['intensity'].raw_data
    out['u8_sum'] = bufs[1]['intensity'].raw_data
    out['u8_sumsig'] = bufs[2]['intensity'].raw_data
    save('dtypes', dict(shape=shape, mask_seed=210, seeds=dict(i32=110, f32=111, u8=112)), **out)
    print({k: str(v.dtype) for k, v in out.items()})


def g_shifts():
    # shifted masks (udf/masks.py:85-124): constant and per-frame shifts
    shape = (4, 5, 16, 12)
    data = synth.dataset(shape, np.float32, seed=113)
    stack = mixed_masks(16, 12, 3, seed=213)
    _, bufs = run(dict(data=data, num_partitions=2, sig_dims=2),
                  [ApplyMasksUDF(mask_factories=lambda: stack, shifts=(2, -3))])
    const = bufs[0]['intensity'].raw_data
    sh = (synth.hash_u32(0, 40, 313) % 9).astype(np.int64).reshape(20, 2) - 4
    sh[3] = (20, 1)   # no overlap at all -> 0
    udf = ApplyMasksUDF(
        mask_factories=lambda: stack,
        shifts=ApplyMasksUDF.aux_data(sh.ravel(), kind='nav', extra_shape=(2,), dtype=sh.dtype),
    )
    _, bufs = run(dict(data=data, num_partitions=2, sig_dims=2), [udf])
    save('shifts', dict(shape=shape, data_seed=113, mask_seed=213, n_masks=3, const=(2, -3)),
         const=const, perframe=bufs[0]['intensity'].raw_data, shifts=sh)


def g_masks_gen():
    arrays = dict(
        circular=M.circular(7.3, 5.1, 20, 16, 4.6),
        ring=M.ring(9, 8, 20, 16, 7.5, 3.2),
        gradient_x=M.gradient_x(20, 16), gradient_y=M.gradient_y(20, 16),
        radial_bins=M.radial_bins(9.5, 8.2, 20, 16, radius=9., radius_inner=0, n_bins=4,
                                  use_sparse=False, dtype=np.float32),
        radial_bins_default=M.radial_bins(10, 8, 20, 16, n_bins=5, use_sparse=False,
                                          dtype=np.float64),
        polar_r=M.polar_map(9.5, 8.2, 20, 16)[0], polar_phi=M.polar_map(9.5, 8.2, 20, 16)[1],
        bounding_radius=np.array(M.bounding_radius(9.5, 8.2, 20, 16)),
    )
    save('masks_gen', {}, **arrays)


def g_rmatmul():
    left = synth.uniform_f32(0, 37 * 300, 114).reshape(37, 300)
    dense = synth.uniform_f32(0, 300 * 6, 214).reshape(300, 6)
    dense[synth.hash_u32(0, 1800, 314).reshape(300, 6) % 5 != 0] = 0
    csr = sp.csr_matrix(dense)
    csc = sp.csc_matrix(dense)
    save('rmatmul', dict(left_seed=114, right_seed=214, sel_seed=314),
         csr=rmatmul(left, csr), csc=rmatmul(left, csc))




def g_corrections():
    # detector corrections (io/corrections/corrset.py): dark, gain, excluded pixels
    from libertem.io.corrections import CorrectionSet
    shape = (4, 5, 16, 12)
    stack = mixed_masks(16, 12, 3, seed=215)
    dark = synth.uniform_f32(0, 16 * 12, 415).reshape(16, 12) * 0.3
    gain = 0.5 + synth.uniform_f32(0, 16 * 12, 515).reshape(16, 12)
    excl = np.zeros((16, 12), dtype=bool)
    for (y, x) in [(0, 0), (3, 4), (3, 5), (15, 11), (8, 0), (9, 7)]:
        excl[y, x] = True
    out = {}
    for dt, seed in ((np.float32, 115), (np.uint16, 116)):
        data = synth.dataset(shape, dt, seed=seed)
        for name, corr in (
            ('dg', CorrectionSet(dark=dark, gain=gain)),
            ('dge', CorrectionSet(dark=dark, gain=gain, excluded_pixels=excl)),
            ('e', CorrectionSet(excluded_pixels=excl)),
        ):
            ex = InlineJobExecutor()
            # NOTE a fresh copy per run: for C-contiguous float32 input the reference corrects
            # the MemoryDataSet's array IN PLACE (io/dataset/memory.py:99-106 hands the view to
            # preprocess), so re-using `data` would stack corrections across runs
            ds = MemoryDataSet(data=data.copy(), num_partitions=2, sig_dims=2)
            ds.initialize(ex)
            res = UDFRunner([ApplyMasksUDF(mask_factories=lambda: stack), CoMUDF(), SumUDF(),
                             SumSigUDF()]).run_for_dataset(ds, ex, corrections=corr)
            b = res.buffers
            key = f'{np.dtype(dt).name}_{name}_'
            out[key + 'intensity'] = b[0]['intensity'].raw_data
            out[key + 'raw_com'] = b[1]['raw_com'].raw_data
            out[key + 'sum'] = b[2]['intensity'].raw_data
            out[key + 'sumsig'] = b[3]['intensity'].raw_data
    save('corrections', dict(shape=shape, mask_seed=215, dark_seed=415, gain_seed=515,
                             seeds=dict(float32=115, uint16=116),
                             excluded=[(0, 0), (3, 4), (3, 5), (15, 11), (8, 0), (9, 7)]), **out)


def g_int_detector():
    # integer detectors (the inputs of the int8 tensor-core path of libertem_b200): CoM with a
    # disk, SumUDF, SumSigUDF and two binary masks on uint16 (256x256) and uint8 (64x64) frames
    cases = {
        'u16': dict(shape=(16, 64, 256, 256), dtype=np.uint16, seed=109,
                    com=dict(cy=120, cx=131, r=100),
                    masks=lambda: np.stack([M.circular(128, 128, 256, 256, 40),
                                            M.ring(128, 128, 256, 256, 90, 60)])),
        'u8': dict(shape=(16, 32, 64, 64), dtype=np.uint8, seed=111, com=dict(),
                   masks=lambda: np.stack([M.circular(32, 32, 64, 64, 10),
                                           M.ring(32, 32, 64, 64, 30, 20)])),
    }
    for name, c in cases.items():
        shape = c['shape']
        if c['dtype'] == np.uint16:
            data = synth.dataset(shape, np.uint16, seed=c['seed'])
        else:
            data = (synth.hash_u32(0, int(np.prod(shape)), c['seed']) % 23).astype(
                np.uint8).reshape(shape)
        stack = c['masks']().astype(np.float32)
        kw = dict(data=data, num_partitions=2, sig_dims=2)
        _, bufs = run(kw, [CoMUDF.with_params(**c['com']), SumUDF(), SumSigUDF(),
                           ApplyMasksUDF(mask_factories=lambda: stack, mask_count=2,
                                         mask_dtype=np.float32, use_sparse=False)])
        arrays = {'com_' + k: bufs[0][k].raw_data for k in COM_KEYS}
        arrays['com_raw_mask_result'] = com_raw(kw, cy=c['com'].get('cy'), cx=c['com'].get('cx'),
                                                r=c['com'].get('r', float('inf')))
        arrays['sum'] = bufs[1]['intensity'].data
        arrays['sumsig'] = bufs[2]['intensity'].raw_data
        arrays['intensity'] = bufs[3]['intensity'].raw_data
        save('int_detector_' + name, dict(shape=shape, data_seed=c['seed'], com=c['com'],
                                          num_partitions=2), **arrays)


if __name__ == '__main__':
    which = sys.argv[1:] or None
    for name, fn in list(globals().items()):
        if name.startswith('g_') and (which is None or name[2:] in which):
            fn()
