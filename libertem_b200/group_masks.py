"""Group-sparse mask plans for the K4 kernel (``ltb200_group_masks``).

A *group* is a run of consecutive masks that share one pixel support -- e.g. all orders
``exp(i*o*phi)`` of one ring of the reference's ``radial_mask_factory``
(src/libertem/analysis/radialfourier.py:106-146).  Per group the contraction is a dense GEMM
over the support's pixels only, so the kernel gathers those pixels once and applies all of the
group's (<= 28 complex) columns to them.

Packed table layout (what the kernel's TMA box expects): 28 pair rows x (2 * n_entries) floats;
within every block of 32 entries the 64 floats of a row are ordered
``(c // 2) * 32 + q * 4 + (c % 2) * 2 + which`` for entry ``4*q + c`` and component ``which``
(0 = real, 1 = imag), which makes the two LDS.128 of a pixel lane bank-conflict free.
"""
import os

import numpy as np
import torch

from . import _lib
from . import walk_plan
from ._lib import get_lib, check

KT = 128          # entries per pipeline stage (K4_KT)
MAX_PAIRS = 28    # K4_NPR


def tf32_round(a):
    """float32 array rounded to TF32 precision (10 explicit mantissa bits, nearest, ties away):
    the values ``tcgen05.mma.kind::tf32`` sees when the low 13 bits are already zero"""
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split_table(table, n_cols):
    """(pairs, n_entries, 2) complex weights -> the (N, n_entries) hi/lo table of the tensor-core
    kernel K7 (row order documented in include/ltb200.h, ltb200_group_masks_tc)"""
    pairs, n, _ = table.shape
    nq = n_cols // 4
    real = np.ascontiguousarray(table.transpose(0, 2, 1)).reshape(2 * pairs, n)   # r = 2*pair + w
    hi = tf32_round(real)
    lo = tf32_round(real - hi)
    out = np.zeros((n_cols, n), dtype=np.float32)
    for r in range(min(2 * pairs, 2 * nq)):
        h, j = divmod(r, nq)
        out[h * 2 * nq + j] = hi[r]
        out[h * 2 * nq + nq + j] = lo[r]
    return out


class GroupPlan:
    def __init__(self, entry_px, table_packed, group_off, n_groups, n_pairs, n_masks, device,
                 table_split=None, n_cols=0, banded=None):
        self.entry_px = torch.from_numpy(entry_px).to(device)
        self.table = torch.from_numpy(table_packed).to(device)
        #: (N, n_entries) hi/lo weight table of the tensor-core kernel (None: FFMA2 kernel only)
        self.table_split = None if table_split is None else torch.from_numpy(table_split).to(device)
        self.n_cols = n_cols
        self.group_off_host = np.ascontiguousarray(group_off, dtype=np.int32)
        self.group_off_dev = torch.from_numpy(self.group_off_host).to(device)
        self.n_groups = n_groups
        self.n_pairs = n_pairs
        self.n_masks = n_masks
        self.n_entries = int(group_off[-1])
        self.workspace = torch.zeros(64, dtype=torch.int32, device=device)
        #: banded twin of the entry list for the tensor-core kernel on large signals:
        #: dict(n_bands, entry_px, table_split, group_off_host, group_off_dev, n_groups)
        self.banded = banded
        self._band_ws = None
        #: opt-in mirror-symmetric plan (dict(main=..., rest=...)) or None
        self.sym = None
        #: dense-walk plan of K10 (device tensors of walk_plan.build_walk) or None
        self.walk = None
        self._walk_ws = None

    def walk_workspace(self, n_frames, accumulate):
        need = get_lib().ltb200_group_masks_walk_workspace(
            n_frames, self.n_groups, self.n_pairs, int(bool(accumulate)))
        if self._walk_ws is None or self._walk_ws.numel() < need:
            self._walk_ws = torch.zeros(need, dtype=torch.uint8, device=self.entry_px.device)
        return self._walk_ws

    def band_workspace(self, n_frames, which=None):
        which = self.banded if which is None else which
        need = get_lib().ltb200_group_masks_tc_workspace(
            n_frames, which['n_groups'], self.n_pairs, which['n_bands'])
        if self._band_ws is None or self._band_ws.numel() < need:
            self._band_ws = torch.zeros(need, dtype=torch.uint8, device=self.entry_px.device)
        return self._band_ws


def find_groups(stack2d, max_pairs=MAX_PAIRS):
    """split consecutive masks (rows of the (M, K) complex matrix) into runs with identical
    support; returns the common group size or None if the runs are not uniform"""
    support = stack2d != 0
    M = stack2d.shape[0]
    sizes = []
    i = 0
    while i < M:
        j = i + 1
        while j < M and j - i < max_pairs and np.array_equal(support[j], support[i]):
            j += 1
        sizes.append(j - i)
        i = j
    if len(set(sizes)) != 1:
        return None
    return sizes[0]


def pack_rows(table):
    """(rows, n_entries, 2) -> (rows, 2*n_entries) in the kernel's conflict-free order"""
    rows, n, _ = table.shape
    assert n % 32 == 0
    t = table.reshape(rows, n // 32, 8, 4, 2)            # [row][block][q][c][which]
    t = t.reshape(rows, n // 32, 8, 2, 2, 2)             # c -> (c//2, c%2)
    t = t.transpose(0, 1, 3, 2, 4, 5)                    # [row][block][c//2][q][c%2][which]
    return np.ascontiguousarray(t.reshape(rows, 2 * n))


TC_KT = 64                        # entries per stage of the tensor-core kernel (K7_KT)
#: mirror-symmetric plan (ltb200_group_masks_tc_sym) for stacks with m(sy - y, x) = conj m(y, x):
#: validated on B200 (tests/test_k4_gpu.py::test_group_masks_sym) and timed in round 2 -- 2.94 ms
#: vs 3.02 ms for the banded plan on the cfg4 geometry (8192 frames): the kernel is bound by the
#: gather / L2->SM fabric, not by the tensor work or the table bytes the symmetry saves, so it
#: stays opt-in (build_plan(sym=True) or LTB200_K7_SYM=1); the banded plan is the default.
SYM_PATH = os.environ.get('LTB200_K7_SYM', '0') == '1'
#: dense-walk plan (K10): the default for stacks it accepts; LTB200_K10=0 restores K7 for A/B
WALK_PATH = os.environ.get('LTB200_K10', '1') != '0'
#: frame-stream bytes of one band of one 128-frame block aimed at by the band count (a handful
#: of frame blocks x this must stay L2-resident together with the band's weight-table slices)
BAND_TARGET_BYTES = 32 << 20


def default_bands(sig_size):
    n = int(round(sig_size * 4 * 128 / BAND_TARGET_BYTES))
    return max(1, min(32, n))


def build_banded(flat, group_size, n_bands, n_cols):
    """(band, ring) groups of QUADS for the tensor-core kernel: band b holds the pixels
    [b * K / n_bands, (b + 1) * K / n_bands) of the flattened signal (row bands; K % (4 n_bands)
    == 0).  A ring's pixels inside a band are covered by quads = 4 consecutive pixels starting at
    a multiple of 4 (one 16-byte copy per frame); the pixels of a quad outside the ring carry
    weight 0.  Every group is padded to a multiple of TC_KT entries.  Returns quad_px (first
    pixel of every quad), the split weight table over the entries (4 per quad) and offsets."""
    M, K = flat.shape
    n_rings = M // group_size
    if K % (4 * n_bands):
        return None
    edges = [(b * K) // n_bands for b in range(n_bands + 1)]
    quad_list, tabs, offs = [], [], [0]
    supports = []
    for g in range(n_rings):
        rows = flat[g * group_size:(g + 1) * group_size]
        px = np.nonzero(np.any(rows != 0, axis=0))[0]
        supports.append((rows, np.unique(px >> 2).astype(np.int64)))
    for b in range(n_bands):
        for rows, quads_all in supports:
            q = quads_all[(quads_all * 4 >= edges[b]) & (quads_all * 4 < edges[b + 1])]
            nq = len(q)
            nq_pad = ((4 * nq + TC_KT - 1) // TC_KT) * (TC_KT // 4)
            qpx = np.zeros(nq_pad, dtype=np.int32)
            qpx[:nq] = 4 * q
            if nq:
                qpx[nq:] = 4 * q[-1]          # padding re-reads the last quad (weight 0)
            px = (4 * q[:, None] + np.arange(4)[None, :]).reshape(-1)
            tab = np.zeros((group_size, 4 * nq_pad, 2), dtype=np.float32)
            vals = rows[:, px]
            tab[:, :4 * nq, 0] = vals.real
            tab[:, :4 * nq, 1] = vals.imag
            quad_list.append(qpx)
            tabs.append(tab)
            offs.append(offs[-1] + 4 * nq_pad)
    if offs[-1] == 0:
        return None
    return dict(n_bands=n_bands, n_groups=n_bands * n_rings, entry_px=np.concatenate(quad_list),
                table_split=split_table(np.concatenate(tabs, axis=1), n_cols),
                group_off=np.array(offs, dtype=np.int32))


def split_table_sym(avg):
    """(pairs <= 28, n_orbits) complex orbit-averaged weights -> the (128, n_orbits) table of the
    symmetric kernel: rows [0, 32) hi(Re), [32, 64) lo(Re), [64, 96) hi(Im), [96, 128) lo(Im)"""
    pairs, n = avg.shape
    out = np.zeros((128, n), dtype=np.float32)
    for part, base in ((np.ascontiguousarray(avg.real, dtype=np.float32), 0),
                       (np.ascontiguousarray(avg.imag, dtype=np.float32), 64)):
        hi = tf32_round(part)
        out[base:base + pairs] = hi
        out[base + 32:base + 32 + pairs] = tf32_round(part - hi)
    return out


#: largest |m(sy - y, x) - conj(m(y, x))| for which a stack counts as mirror-symmetric (the
#: reference's complex64 radial masks: 1.1e-5, mean 7e-11)
SYM_TOL = 2e-5


def build_sym(stack3, group_size, n_bands):
    """Mirror-symmetric plan of the tensor-core kernel (opt-in, ltb200_group_masks_tc_sym)
    for a (M, sy, sx) complex stack with m(sy - y, x) = conj(m(y, x)), or None.

    Orbits are pixel pairs {(y, x), (sy - y, x)}, 1 <= y < sy/2.  Groups are (row band of the
    upper half, ring) pairs, band-major.  Per stage of 64 entries the quad list holds 8 upper
    quads followed by their 8 mirrored quads; the weight table has one column per orbit (32 per
    stage) with the orbit-averaged weight (m(p) + conj(m(p'))) / 2 of the upper pixel.  Rows 0
    and sy/2 have no partner: ``residual`` is a boolean (sy, sx) map of the pixels the caller
    must still run through the ordinary quad plan."""
    M, sy, sx = stack3.shape
    if sy % 2 or sx % 4 or sy < 4 or group_size > 28:
        return None
    cy = sy // 2
    n_rings = M // group_size
    up = stack3[:, 1:cy, :]
    dn = stack3[:, sy - 1:cy:-1, :]                # row sy - y for y = 1 .. cy - 1
    if np.abs(dn - np.conj(up)).max() > SYM_TOL:
        return None
    avg = ((up.astype(np.complex128) + np.conj(dn.astype(np.complex128))) / 2)
    rows_up = cy - 1
    n_bands = max(1, min(n_bands, rows_up))
    edges = [1 + (b * rows_up) // n_bands for b in range(n_bands + 1)]     # rows of band b
    quad_list, tabs, offs = [], [], [0]
    supports = []
    for g in range(n_rings):
        rows = avg[g * group_size:(g + 1) * group_size]                    # (size, cy-1, sx)
        yy, xx = np.nonzero(np.any(rows != 0, axis=0))
        q = np.unique((yy + 1) * (sx // 4) + xx // 4)                      # quad id = y * sx/4 + x/4
        supports.append((rows, q))
    for b in range(n_bands):
        for rows, q_all in supports:
            qy = q_all // (sx // 4)
            q = q_all[(qy >= edges[b]) & (qy < edges[b + 1])]
            nq = len(q)
            nq_pad = ((nq + 7) // 8) * 8
            qp = np.zeros(nq_pad, dtype=np.int64)
            qp[:nq] = q
            if nq:
                qp[nq:] = q[-1]
            y, x0 = qp // (sx // 4), (qp % (sx // 4)) * 4
            upper = (y * sx + x0).astype(np.int32)
            lower = ((sy - y) * sx + x0).astype(np.int32)
            # per stage: 8 upper quads, then their 8 mirror images
            quads = np.concatenate([upper.reshape(-1, 8), lower.reshape(-1, 8)], axis=1)
            quad_list.append(quads.reshape(-1))
            tab = np.zeros((rows.shape[0], 4 * nq_pad), dtype=np.complex128)
            if nq:
                px_y = np.repeat(y[:nq], 4) - 1                            # row index into avg
                px_x = (x0[:nq, None] + np.arange(4)[None, :]).reshape(-1)
                tab[:, :4 * nq] = rows[:, px_y, px_x]
            tabs.append(tab)
            offs.append(offs[-1] + 8 * nq_pad)                             # entries: 2 x 4 per quad
    if offs[-1] == 0:
        return None
    residual = np.zeros((sy, sx), dtype=bool)
    residual[[0, cy], :] = True
    return dict(n_bands=n_bands, n_groups=n_bands * n_rings,
                entry_px=np.concatenate(quad_list).astype(np.int32),
                table_sym=split_table_sym(np.concatenate(tabs, axis=1)),
                group_off=np.array(offs, dtype=np.int32), residual=residual)


def build_plan(stack, group_size, device, n_bands=None, sig_shape=None, sym=None,
               walk_max_dup=1.5):
    """stack: complex (M, *sig) dense array with M = n_groups * group_size (``sig_shape`` gives
    the 2D signal shape when the stack comes flattened)"""
    M = stack.shape[0]
    if sig_shape is None:
        sig_shape = tuple(np.asarray(stack).shape[1:])
    sig_shape = tuple(int(v) for v in sig_shape)
    flat = np.asarray(stack).reshape(M, -1).astype(np.complex64)
    if int(np.prod(sig_shape)) != flat.shape[1]:
        sig_shape = (flat.shape[1],)
    n_groups = M // group_size
    assert n_groups * group_size == M and group_size <= MAX_PAIRS
    px_list, offs = [], [0]
    tables = []
    for g in range(n_groups):
        rows = flat[g * group_size:(g + 1) * group_size]
        px = np.nonzero(np.any(rows != 0, axis=0))[0].astype(np.int32)
        n = len(px)
        n_pad = ((n + KT - 1) // KT) * KT
        ent = np.zeros(n_pad, dtype=np.int32)
        ent[:n] = px
        tab = np.zeros((MAX_PAIRS, n_pad, 2), dtype=np.float32)
        vals = rows[:, px]
        tab[:group_size, :n, 0] = vals.real
        tab[:group_size, :n, 1] = vals.imag
        px_list.append(ent)
        tables.append(tab)
        offs.append(offs[-1] + n_pad)
    entry_px = np.concatenate(px_list) if px_list else np.zeros(0, np.int32)
    table = np.concatenate(tables, axis=1) if tables else np.zeros((MAX_PAIRS, 0, 2), np.float32)
    if entry_px.size == 0:
        # keep the TMA descriptor valid: one all-zero stage
        entry_px = np.zeros(KT, dtype=np.int32)
        table = np.zeros((MAX_PAIRS, KT, 2), dtype=np.float32)
    n_cols = get_lib().ltb200_group_masks_tc_columns(group_size)
    split = split_table(table[:group_size], n_cols) if n_cols else None
    if n_bands is None:
        n_bands = int(os.environ.get('LTB200_K7_BANDS', 0)) or default_bands(flat.shape[1])
    banded = None
    while n_bands > 1 and flat.shape[1] % (4 * n_bands):
        n_bands -= 1
    if n_cols and n_groups > 0 and flat.shape[1] % 4 == 0:
        b = build_banded(flat, group_size, n_bands, n_cols)
        if b is not None:
            banded = dict(n_bands=b['n_bands'], n_groups=b['n_groups'],
                          entry_px=torch.from_numpy(b['entry_px']).to(device),
                          table_split=torch.from_numpy(b['table_split']).to(device),
                          group_off_host=b['group_off'],
                          group_off_dev=torch.from_numpy(b['group_off']).to(device))
    plan = GroupPlan(entry_px, pack_rows(table), np.array(offs, dtype=np.int32), n_groups,
                     group_size, M, device, table_split=split, n_cols=n_cols, banded=banded)
    if WALK_PATH and n_cols and n_groups > 0:
        w = walk_plan.build_walk(flat, group_size, max_dup=walk_max_dup)
        if w is not None:
            def dev_u32(a):
                if len(a) == 0:                     # keep the pointer valid
                    a = np.zeros(1, dtype=np.uint32)
                return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(device)
            plan.walk = dict(
                n_segments=w['n_segments'], n_entries=w['n_entries'],
                boxes=dev_u32(w['boxes']),
                ops=[dev_u32(w[f'ops{c}']) for c in range(4)],
                events=[dev_u32(w['events0']), dev_u32(w['events1'])],
                # (an issuer without ops has an empty table: keep the pointer valid)
                table=[torch.from_numpy(w[f'table{c}'] if len(w[f'table{c}']) else
                                        np.zeros((1,) + w[f'table{c}'].shape[1:], np.float32)
                                        ).to(device) for c in range(4)],
                seg_off_host=np.ascontiguousarray(
                    np.stack([w['visit_off']] + [w[f'op_off{c}'] for c in range(4)] +
                             [w[f'tab_off{c}'] for c in range(4)] +
                             [w['ev_off0'], w['ev_off1']]), dtype=np.int32))
    if (SYM_PATH if sym is None else sym) and banded is not None and len(sig_shape) == 2:
        plan.sym = _sym_to_device(flat, sig_shape, group_size, max(1, n_bands // 2), n_cols,
                                  device)
    return plan


def _sym_to_device(flat, sig_shape, group_size, n_bands, n_cols, device):
    """device form of build_sym + the ordinary quad plan of the rows without a mirror partner"""
    sy, sx = sig_shape
    b = build_sym(flat.reshape(-1, sy, sx), group_size, n_bands)
    if b is None:
        return None
    keep = b['residual'].reshape(-1)
    rest = build_banded(np.where(keep[None, :], flat, 0), group_size, 1, n_cols)

    def dev(d, table_key):
        return dict(n_bands=d['n_bands'], n_groups=d['n_groups'],
                    entry_px=torch.from_numpy(d['entry_px']).to(device),
                    table=torch.from_numpy(d[table_key]).to(device),
                    group_off_host=d['group_off'],
                    group_off_dev=torch.from_numpy(d['group_off']).to(device))
    return dict(main=dev(b, 'table_sym'), rest=None if rest is None else dev(rest, 'table_split'))


#: frames from which the tensor-core kernel (128-frame items) is preferred over the FFMA2 one
TC_MIN_FRAMES = 96


def group_masks(tile, plan, out=None, accumulate=False, kernel='auto', chain=0):
    """out (F, n_masks) complex64 (+)= group-sparse contraction of the float32 tile.

    kernel: 'auto' (tensor cores from TC_MIN_FRAMES frames: the dense-walk kernel K10 when the
    stack has such a plan and the frame rows are 16-byte aligned, else K7 with the quad / banded
    plan, else K7's 4-byte ring-major gather), 'walk' (K10), 'tc' (K7, 4-byte gather,
    ring-major), 'banded' (K7, quad gather, banded schedule) or 'ffma' (K4)"""
    lib = get_lib()
    if not tile.is_cuda:
        raise _lib.LTB200Error('tile must be a CUDA tensor (no CPU fallback)')
    if tile.dtype != torch.float32:
        raise TypeError('group-sparse masks need float32 tiles')
    if tile.shape[1] > 0 and tile.stride(1) != 1:
        tile = tile.contiguous()
    F, K = tile.shape
    if out is None:
        out = torch.zeros((F, plan.n_masks), dtype=torch.complex64, device=tile.device)
        accumulate = False
    real = torch.view_as_real(out).reshape(F, 2 * plan.n_masks)
    ld_tile = tile.stride(0) if F > 1 else max(K, 1)
    ld_out = real.stride(0) if F > 1 else max(2 * plan.n_masks, 1)
    use_tc = kernel in ('tc', 'banded', 'sym', 'walk') or (kernel == 'auto' and F >= TC_MIN_FRAMES)
    if use_tc and plan.table_split is None:
        if kernel in ('tc', 'banded', 'sym', 'walk'):
            raise _lib.LTB200Error('no tensor-core table for this plan')
        use_tc = False
    if kernel == 'banded' and plan.banded is None:
        raise _lib.LTB200Error('no quad plan (sig_size not a multiple of 4)')
    aligned = tile.data_ptr() % 16 == 0 and ld_tile % 4 == 0
    if kernel == 'banded' and not aligned:
        raise _lib.LTB200Error('the banded plan needs 16-byte aligned frame rows')
    if kernel == 'walk' and (plan.walk is None or not aligned):
        raise _lib.LTB200Error('no dense-walk plan for this stack / tile (sig_size % 32, 16-byte '
                               'aligned frame rows)')
    if use_tc and plan.walk is not None and aligned and (
            kernel == 'walk' or (kernel == 'auto' and plan.sym is None)):
        w = plan.walk
        ws = plan.walk_workspace(F, accumulate)
        with torch.cuda.device(tile.device):
            check(lib.ltb200_group_masks_walk(
                tile.data_ptr(), F, K, ld_tile, w['boxes'].data_ptr(), w['ops'][0].data_ptr(),
                w['ops'][1].data_ptr(), w['ops'][2].data_ptr(), w['ops'][3].data_ptr(),
                w['events'][0].data_ptr(), w['events'][1].data_ptr(),
                w['table'][0].data_ptr(), w['table'][1].data_ptr(), w['table'][2].data_ptr(),
                w['table'][3].data_ptr(), w['seg_off_host'].ctypes.data,
                w['n_segments'], plan.n_groups, plan.n_pairs, real.data_ptr(), ld_out,
                int(bool(accumulate)), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream(tile.device).cuda_stream))
        return out
    if kernel == 'sym' and plan.sym is None:
        raise _lib.LTB200Error('no mirror-symmetric plan (build_plan(sym=True) and a symmetric stack)')
    if use_tc and plan.sym is not None and aligned and kernel in ('sym', 'auto'):
        stream = torch.cuda.current_stream(tile.device).cuda_stream
        with torch.cuda.device(tile.device):
            m = plan.sym['main']
            ws = plan.band_workspace(F, m)
            check(lib.ltb200_group_masks_tc_sym(
                tile.data_ptr(), F, K, ld_tile, m['entry_px'].data_ptr(), m['table'].data_ptr(),
                m['group_off_host'].ctypes.data, m['group_off_dev'].data_ptr(), m['n_groups'],
                plan.n_pairs, m['n_bands'], real.data_ptr(), ld_out, int(bool(accumulate)),
                int(chain), ws.data_ptr(), ws.numel(), stream))
            r = plan.sym['rest']
            if r is not None:
                ws = plan.band_workspace(F, r)
                check(lib.ltb200_group_masks_tc_banded(
                    tile.data_ptr(), F, K, ld_tile, r['entry_px'].data_ptr(),
                    r['table'].data_ptr(), r['group_off_host'].ctypes.data,
                    r['group_off_dev'].data_ptr(), r['n_groups'], plan.n_pairs, r['n_bands'],
                    real.data_ptr(), ld_out, 1, int(chain), ws.data_ptr(), ws.numel(), stream))
        return out
    if use_tc and plan.banded is not None and aligned and kernel in ('banded', 'auto'):
        b = plan.banded
        ws = plan.band_workspace(F)
        with torch.cuda.device(tile.device):
            check(lib.ltb200_group_masks_tc_banded(
                tile.data_ptr(), F, K, ld_tile, b['entry_px'].data_ptr(),
                b['table_split'].data_ptr(), b['group_off_host'].ctypes.data,
                b['group_off_dev'].data_ptr(), b['n_groups'], plan.n_pairs, b['n_bands'],
                real.data_ptr(), ld_out, int(bool(accumulate)), int(chain), ws.data_ptr(),
                ws.numel(), torch.cuda.current_stream(tile.device).cuda_stream))
        return out
    if use_tc:
        with torch.cuda.device(tile.device):
            check(lib.ltb200_group_masks_tc(
                tile.data_ptr(), F, K, ld_tile, plan.entry_px.data_ptr(),
                plan.table_split.data_ptr(), plan.group_off_host.ctypes.data,
                plan.group_off_dev.data_ptr(), plan.n_groups, plan.n_pairs, real.data_ptr(),
                ld_out, int(bool(accumulate)), int(chain), plan.workspace.data_ptr(),
                plan.workspace.numel() * 4, torch.cuda.current_stream(tile.device).cuda_stream))
        return out
    with torch.cuda.device(tile.device):
        check(lib.ltb200_group_masks(
            tile.data_ptr(), _lib.LTB_F32, F, K, ld_tile, plan.entry_px.data_ptr(),
            plan.table.data_ptr(), plan.group_off_host.ctypes.data, plan.group_off_dev.data_ptr(),
            plan.n_groups, plan.n_pairs, real.data_ptr(), ld_out, int(bool(accumulate)),
            plan.workspace.data_ptr(), plan.workspace.numel() * 4,
            torch.cuda.current_stream(tile.device).cuda_stream))
    return out
