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
import numpy as np
import torch

from . import _lib
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
                 table_split=None, n_cols=0):
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


def build_plan(stack, group_size, device):
    """stack: complex (M, *sig) dense array with M = n_groups * group_size"""
    M = stack.shape[0]
    flat = np.asarray(stack).reshape(M, -1).astype(np.complex64)
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
    return GroupPlan(entry_px, pack_rows(table), np.array(offs, dtype=np.int32), n_groups,
                     group_size, M, device, table_split=split, n_cols=n_cols)


#: frames from which the tensor-core kernel (128-frame items) is preferred over the FFMA2 one
TC_MIN_FRAMES = 96


def group_masks(tile, plan, out=None, accumulate=False, kernel='auto', chain=0):
    """out (F, n_masks) complex64 (+)= group-sparse contraction of the float32 tile.

    kernel: 'auto' (tensor cores, K7, from TC_MIN_FRAMES frames), 'tc' (K7) or 'ffma' (K4)"""
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
    use_tc = kernel == 'tc' or (kernel == 'auto' and F >= TC_MIN_FRAMES)
    if use_tc and plan.table_split is None:
        if kernel == 'tc':
            raise _lib.LTB200Error('no tensor-core table for this plan')
        use_tc = False
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
