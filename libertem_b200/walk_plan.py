"""Dense-walk plan of the group-sparse tensor-core kernel K10 (``ltb200_group_masks_walk``).

The quad / banded plan of K7 gathers every ring's pixels with 16-byte ``cp.async`` copies and
is bound by the rate of those copies.  Here every pixel of a frame crosses the
L2 -> SM fabric ONCE, as part of a dense TMA box ``[128 frames x 32 px]`` (a 128-byte piece of
an image row), and the separation into groups (rings) is done by the weights: the unit of
tensor-core work (an *op*) is a pair (slice of 8 consecutive pixels, group touching it) whose
``[8 x 2G]`` weight block is zero for the pixels of the slice outside the group.

What makes this possible is the ORDER of the boxes: they are visited sorted by the first group
they touch.  The groups of the reference's ``radial_mask_factory``
(src/libertem/analysis/radialfourier.py:106-146) are concentric rings, a box touches <= 4
adjacent ones, so at any time only a window of ``W_LIVE`` = 4 groups has an open accumulator in
tensor memory.  Everything that depends on that order is decided HERE, once per mask stack,
and the kernel only follows lists (the same for every block of 128 frames):

* ``boxes``  -- per visit: first pixel of the box | 4-bit mask of the slices that are used;
* ``ops``    -- per (slice, group): the TMEM accumulator buffer (a pool of ``NBUF``), whether the
  op starts a chain (zero-initialise, wait for the drain of the buffer's previous chain),
  whether it ends one (hand the buffer to the drain warps), its slice of the box, whether it
  is the first / last op of its box (A-operand stage hand-over with the converter warps);
* ``events`` -- per chain, in commit order: buffer, register slot, group (its parity selects
  the drain warps), last-chain-of-the-group (write the result);
* ``table``  -- the split-TF32 weight blocks in op order as the byte image of the kernel's
  shared-memory stages (4 ops = 32 entries per stage, rows [hi | lo], 128-byte swizzle applied
  on the host so that the kernel issues one contiguous bulk copy per stage).

The accumulate of the tensor core truncates, so chains are cut after ``CHAIN``
ops and summed in float32 registers by the drain warps, as in K6 / K7.

Boxes whose groups span more than ``W_LIVE`` are visited once per window (generic stacks);
segments (contiguous runs of the visit order, at least ``W_LIVE`` groups apart) are independent
work items, so a group receives partial sums from at most two of them.
"""
import numpy as np

W_LIVE = 4        # groups with an open accumulator (window of consecutive group ids)
NBUF = 6          # TMEM accumulator buffers of 64 columns
CHAIN = 8         # ops per accumulation chain
BOX = 32          # pixels per TMA box (128 bytes)
SL = 8            # pixels per slice (K of one tf32 MMA)
HR = 56           # weight rows per half ([hi | lo]): 2 * group_size <= HR
STAGE_OPS = 4     # ops per table stage (32 entries)
STAGE_ROWS = 2 * HR
STAGE_FLOATS = STAGE_ROWS * 32
MAX_SEGMENTS = 8

OP_FIRST = 1 << 3
OP_COMMIT = 1 << 4
OP_NEW_BOX = 1 << 5
OP_END_BOX = 1 << 6
OP_SLICE_SHIFT = 8    # bits 8-9: slice of the box (A-operand slot inside the box's stage)
OP_PARITY_SHIFT = 10  # first op of a chain: parity of the buffer's use count in the segment
OP_OWNER_SHIFT = 11   # bits 11-12: group parity = MMA warp / drain group that owns the op
OP_NOP = 3 << OP_OWNER_SHIFT   # padding: owned by no warp

EV_SLOT = 1 << 3
EV_LAST = 1 << 4
EV_PARITY_SHIFT = 5   # parity of the buffer's use count in the segment


def tf32_round(a):
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _visits(bx):
    """bx: (n_groups, n_boxes) bool -> visit arrays (key, box, lo, hi), sorted by (key, box)"""
    n_groups, n_boxes = bx.shape
    has = bx.any(axis=0)
    gmin = np.argmax(bx, axis=0)
    gmax = n_groups - 1 - np.argmax(bx[::-1], axis=0)
    span = np.where(has, gmax - gmin + 1, 0)
    simple = has & (span <= W_LIVE)
    b = np.nonzero(simple)[0]
    key, box, lo, hi = [gmin[b]], [b], [gmin[b]], [gmax[b]]
    extra = ([], [], [], [])
    for bb in np.nonzero(has & ~simple)[0]:
        g = gmin[bb]
        while g <= gmax[bb]:
            touched = np.nonzero(bx[g:g + W_LIVE, bb])[0]
            if len(touched) == 0:
                g += W_LIVE
                continue
            g0 = g + touched[0]
            g1 = min(g0 + W_LIVE - 1, gmax[bb])
            g1 = g0 + np.nonzero(bx[g0:g1 + 1, bb])[0][-1]
            for lst, v in zip(extra, (g0, bb, g0, g1)):
                lst.append(v)
            g = g1 + 1
    if extra[0]:
        for lst, e in zip((key, box, lo, hi), extra):
            lst.append(np.array(e, dtype=np.int64))
    key, box, lo, hi = (np.concatenate(v).astype(np.int64) for v in (key, box, lo, hi))
    order = np.lexsort((box, key))
    return key[order], box[order], lo[order], hi[order]


def _segment_bounds(key, n_ops_visit, n_segments):
    """split the visit order into <= n_segments runs of about equal op count whose first keys
    are at least W_LIVE apart; returns visit offsets (len = segments + 1)"""
    n = len(key)
    if n == 0:
        return [0, 0]
    cum = np.concatenate([[0], np.cumsum(n_ops_visit)])
    total = cum[-1]
    bounds = [0]
    last_key = key[0]
    for s in range(1, n_segments):
        target = total * s / n_segments
        v = int(np.searchsorted(cum, target))
        if v >= n:
            break
        # move to the first visit of that key
        k = key[v]
        v = int(np.searchsorted(key, k, side='left'))
        if k - last_key < W_LIVE or v <= bounds[-1]:
            k = last_key + W_LIVE
            v = int(np.searchsorted(key, k, side='left'))
            if v >= n:
                break
            k = key[v]
        bounds.append(v)
        last_key = k
    bounds.append(n)
    return bounds


def build_walk(flat, group_size, n_segments=None, chain=CHAIN, max_dup=1.5):
    """flat: (M, K) complex stack, M = n_groups * group_size.  Returns the plan as a dict of
    numpy arrays or None when the stack does not fit this form (K % 32, group_size > HR / 2, or
    boxes that would have to be visited more than ``max_dup`` times on average)."""
    M, K = flat.shape
    if group_size < 1 or M % group_size or K % BOX or 2 * group_size > HR or K >= (1 << 31):
        return None
    n_groups = M // group_size
    if n_groups >= (1 << 20):
        return None
    sup = np.zeros((n_groups, K), dtype=bool)
    for g in range(n_groups):
        sup[g] = np.any(flat[g * group_size:(g + 1) * group_size] != 0, axis=0)
    sl = sup.reshape(n_groups, K // SL, SL).any(axis=2)              # (G, K/8)
    bx = sl.reshape(n_groups, K // BOX, BOX // SL).any(axis=2)       # (G, K/32)
    key, box, lo, hi = _visits(bx)
    n_visits = len(key)
    if n_visits == 0:
        return None
    if n_visits > max_dup * max(1, int(bx.any(axis=0).sum())):
        return None
    # ops of every visit: (slice, group) pairs, slices ascending, groups ascending inside
    sl4 = sl.reshape(n_groups, K // BOX, BOX // SL)
    n_ops_visit = np.zeros(n_visits, dtype=np.int64)
    for v in range(n_visits):
        n_ops_visit[v] = sl4[lo[v]:hi[v] + 1, box[v]].sum()
    if n_segments is None:
        n_segments = max(1, min(MAX_SEGMENTS, n_groups // W_LIVE))
    bounds = _segment_bounds(key, n_ops_visit, n_segments)
    n_seg = len(bounds) - 1

    boxes = np.zeros(n_visits, dtype=np.uint32)
    ops, op_slice, op_group = [], [], []
    events = []
    visit_off, op_off, ev_off = [0], [0], [0]
    # last visit index per group (a group's final op closes its chain and writes the result);
    # per segment, so that a segment is self-contained
    for s in range(n_seg):
        v0, v1 = bounds[s], bounds[s + 1]
        last_op_of_group = {}
        seq = []                                   # (visit, slice, group)
        for v in range(v0, v1):
            m = sl4[lo[v]:hi[v] + 1, box[v]]       # (groups in window, 4)
            mask = 0
            for j in range(BOX // SL):
                gs = np.nonzero(m[:, j])[0]
                if len(gs) == 0:
                    continue
                mask |= 1 << j
                for g in gs:
                    seq.append([v, j, int(lo[v] + g), False, False])
            boxes[v] = np.uint32(box[v] * BOX) | np.uint32(mask)
        for i, e in enumerate(seq):
            last_op_of_group[e[2]] = i
        # accumulator buffers: a pool of NBUF / 2 per group parity (the two MMA warps / drain
        # groups are independent pipelines), FIFO = oldest released first
        free = [list(range(NBUF // 2)), list(range(NBUF // 2, NBUF))]
        uses = [0] * NBUF
        open_buf, open_len = {}, {}
        words = []
        for i, (v, j, g, _, _) in enumerate(seq):
            word = j << OP_SLICE_SHIFT
            par = g & 1
            if g not in open_buf:
                if not free[par]:
                    raise AssertionError('walk plan: accumulator pool exhausted')
                b = open_buf[g] = free[par].pop(0)
                open_len[g] = 0
                word |= OP_FIRST | ((uses[b] & 1) << OP_PARITY_SHIFT)
                if len(open_buf) > W_LIVE:
                    raise AssertionError('walk plan: more than W_LIVE live groups')
            b = open_buf[g]
            word |= b | (par << OP_OWNER_SHIFT)
            open_len[g] += 1
            final = last_op_of_group[g] == i
            if open_len[g] == chain or final:
                word |= OP_COMMIT
                ev = (b | (EV_SLOT if (g >> 1) & 1 else 0) | (EV_LAST if final else 0) |
                      ((uses[b] & 1) << EV_PARITY_SHIFT) | (g << 8))
                events.append(ev)
                uses[b] += 1
                free[par].append(b)
                del open_buf[g], open_len[g]
            words.append(word)
            op_slice.append(int(box[v]) * (BOX // SL) + j)
            op_group.append(g)
        # every buffer is used an even number of times per segment, so that the mbarrier
        # parities of a chain are static: an odd count gets one empty chain (zero weights) on
        # the last slice of the segment
        if seq:
            v, j = seq[-1][0], seq[-1][1]
            for b in range(NBUF):
                if uses[b] & 1:
                    par = 1 if b >= NBUF // 2 else 0
                    words.append((j << OP_SLICE_SHIFT) | b | (par << OP_OWNER_SHIFT) | OP_FIRST |
                                 OP_COMMIT | ((uses[b] & 1) << OP_PARITY_SHIFT))
                    events.append(b | ((uses[b] & 1) << EV_PARITY_SHIFT) | (par << 8))
                    uses[b] += 1
                    seq.append([v, j, -2, False, False])
                    op_slice.append(int(box[v]) * (BOX // SL) + j)
                    op_group.append(-2)
        for i, e in enumerate(seq):                # first / last op of every visit
            if i == 0 or seq[i - 1][0] != e[0]:
                words[i] |= OP_NEW_BOX
            if i == len(seq) - 1 or seq[i + 1][0] != e[0]:
                words[i] |= OP_END_BOX
        ops.extend(words)
        assert not open_buf
        while len(ops) % STAGE_OPS:
            ops.append(OP_NOP)
            op_slice.append(-1)
            op_group.append(-1)
        visit_off.append(v1)
        op_off.append(len(ops))
        ev_off.append(len(events))

    ops = np.array(ops, dtype=np.uint32)
    op_slice = np.array(op_slice, dtype=np.int64)
    op_group = np.array(op_group, dtype=np.int64)
    table = _table_image(flat, group_size, op_slice, op_group)
    return dict(
        n_groups=n_groups, group_size=group_size, sig_size=K, n_segments=n_seg, chain=chain,
        boxes=boxes, ops=ops,
        events=np.array(events, dtype=np.uint32),
        visit_off=np.array(visit_off, dtype=np.int32), op_off=np.array(op_off, dtype=np.int32),
        ev_off=np.array(ev_off, dtype=np.int32),
        table=table, n_entries=int(len(ops)) * SL, op_slice=op_slice, op_group=op_group,
        n_real_ops=int((op_group >= 0).sum()),
    )


def _table_image(flat, group_size, op_slice, op_group):
    """(n_stages, STAGE_ROWS, 32) float32: rows [hi(c) for c < HR | lo(c)], real column
    c = 2 * pair + (0 real, 1 imag); entries of op j of a stage at floats [8 j, 8 j + 8) of a
    row; 16-byte chunks XOR-swizzled with (row & 7) like the TMA 128-byte swizzle"""
    n_ops = len(op_slice)
    n_stages = n_ops // STAGE_OPS
    img = np.zeros((n_stages, STAGE_ROWS, STAGE_OPS, SL), dtype=np.float32)
    real = op_group >= 0
    idx = np.nonzero(real)[0]
    CH = 8192
    for c0 in range(0, len(idx), CH):
        ii = idx[c0:c0 + CH]
        rows = op_group[ii][:, None] * group_size + np.arange(group_size)[None, :]     # (n, G)
        px = op_slice[ii][:, None] * SL + np.arange(SL)[None, :]                       # (n, 8)
        vals = flat[rows[:, :, None], px[:, None, :]].astype(np.complex64)             # (n, G, 8)
        r = np.empty((len(ii), 2 * group_size, SL), dtype=np.float32)
        r[:, 0::2] = vals.real
        r[:, 1::2] = vals.imag
        hi = tf32_round(r)
        lo = tf32_round(r - hi)
        st, sub = ii // STAGE_OPS, ii % STAGE_OPS
        img[st[:, None], np.arange(2 * group_size)[None, :], sub[:, None]] = hi
        img[st[:, None], HR + np.arange(2 * group_size)[None, :], sub[:, None]] = lo
    img = img.reshape(n_stages, STAGE_ROWS, 8, 4)                   # 16-byte chunks
    chunk = np.arange(8)[None, :] ^ (np.arange(STAGE_ROWS)[:, None] & 7)               # (rows, 8)
    out = np.empty_like(img)
    out[:, np.arange(STAGE_ROWS)[:, None], chunk] = img
    return np.ascontiguousarray(out.reshape(n_stages, STAGE_ROWS, 32))


def unswizzle_table(table):
    """inverse of the chunk swizzle: (n_stages, rows, STAGE_OPS, SL) weights"""
    n_stages = table.shape[0]
    t = table.reshape(n_stages, STAGE_ROWS, 8, 4)
    chunk = np.arange(8)[None, :] ^ (np.arange(STAGE_ROWS)[:, None] & 7)
    return t[:, np.arange(STAGE_ROWS)[:, None], chunk].reshape(n_stages, STAGE_ROWS, STAGE_OPS, SL)


def emulate(plan, tile):
    """numpy model of what the kernel does with the lists (float64 arithmetic; checks the
    invariants the kernel relies on).  tile: (F, K) -> (F, n_groups * group_size) complex128"""
    F = tile.shape[0]
    G, gs = plan['n_groups'], plan['group_size']
    out = np.zeros((F, G, 2 * gs), dtype=np.float64)
    w = unswizzle_table(plan['table']).astype(np.float64)
    w = w[:, :HR] + w[:, HR:]                                        # (stages, HR, 4, 8)
    written = np.zeros(G, dtype=np.int64)
    for s in range(plan['n_segments']):
        bufs = np.zeros((NBUF, F, HR))
        busy = [False] * NBUF
        uses = [0] * NBUF
        acc = np.zeros((2, 2, F, HR))
        slot_owner = [[None, None], [None, None]]
        evs = list(plan['events'][plan['ev_off'][s]:plan['ev_off'][s + 1]])
        v = plan['visit_off'][s] - 1
        px0 = mask = 0
        open_groups = {}
        for i in range(plan['op_off'][s], plan['op_off'][s + 1]):
            word = int(plan['ops'][i])
            if (word & OP_NOP) == OP_NOP:
                assert word == OP_NOP
                continue
            g = int(plan['op_group'][i])
            dummy = g == -2
            if word & OP_NEW_BOX:
                v += 1
                px0, mask = int(plan['boxes'][v]) & ~31, int(plan['boxes'][v]) & 15
            j = (word >> OP_SLICE_SHIFT) & 3
            assert mask >> j & 1                   # the converters fill only the masked slices
            assert px0 + j * SL == plan['op_slice'][i] * SL
            b = word & 7
            assert ((word >> OP_OWNER_SHIFT) & 1) == (1 if b >= NBUF // 2 else 0)
            if dummy:
                ev = int(evs.pop(0))
                assert word & OP_FIRST and word & OP_COMMIT and not busy[b]
                assert (ev & 7) == b and not ev & EV_LAST
                assert ((word >> OP_PARITY_SHIFT) & 1) == (uses[b] & 1) == ((ev >> EV_PARITY_SHIFT) & 1)
                assert not np.any(w[i // STAGE_OPS, :, i % STAGE_OPS])
                uses[b] += 1
                continue
            assert ((word >> OP_OWNER_SHIFT) & 1) == (g & 1)
            x = tile[:, px0 + j * SL:px0 + (j + 1) * SL].astype(np.float64)    # (F, 8)
            prod = x @ w[i // STAGE_OPS, :, i % STAGE_OPS].T                   # (F, HR)
            if word & OP_FIRST:
                assert not busy[b] and g not in open_groups
                assert ((word >> OP_PARITY_SHIFT) & 1) == (uses[b] & 1)
                busy[b] = True
                open_groups[g] = b
                bufs[b] = prod
            else:
                assert busy[b] and open_groups[g] == b
                bufs[b] += prod
            assert len(open_groups) <= W_LIVE
            assert max(open_groups) - min(open_groups) < W_LIVE
            if word & OP_COMMIT:
                p = g & 1
                ev = int(evs.pop(0))
                assert (ev & 7) == b and (ev >> 8) == g
                assert ((ev >> EV_PARITY_SHIFT) & 1) == (uses[b] & 1)
                uses[b] += 1
                slot = 1 if ev & EV_SLOT else 0
                assert slot == ((g >> 1) & 1)
                assert slot_owner[p][slot] in (None, g)
                slot_owner[p][slot] = g
                acc[p, slot] += bufs[b]
                busy[b] = False
                del open_groups[g]
                if ev & EV_LAST:
                    out[:, g] += acc[p, slot][:, :2 * gs]
                    written[g] += 1
                    acc[p, slot] = 0
                    slot_owner[p][slot] = None
        assert v == plan['visit_off'][s + 1] - 1
        assert not any(busy) and not evs and not open_groups
        assert not any(u & 1 for u in uses)
        assert slot_owner == [[None, None], [None, None]]
    assert written.max() <= 2
    return (out[..., 0::2] + 1j * out[..., 1::2]).reshape(F, G * gs)
