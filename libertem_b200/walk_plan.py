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
tensor memory.  Everything that depends on that order is decided HERE, once per mask stack:
the kernel is an interpreter of static lists ("microcode", the same for every block of 128
frames) and computes nothing but shifts and masks of their words.

The groups of even and odd id form two independent pipelines (an MMA issuer warp, a pool of
``NBUF / 2`` accumulator buffers, four drain warps and a weight-table stream each) that share
only the boxes and their A-operand stages:

* ``boxes``   -- per visit: first pixel of the box | 4-bit mask of the slices that are used;
  every segment holds a multiple of 4 visits (padded with empty ones), so that the A stage and
  the mbarrier parity of a box are static;
* ``ops[c]``  -- per issuer warp c = g % 4 (two per pipeline) its own list; per op, in walk order: accumulator buffer, first / last op of an
  accumulation chain (+ the static mbarrier parity), slice and A stage of the box, first / last
  op of the pipeline in its box (A-stage hand-over with the converter warps; a box without ops
  of the pipeline gets an empty marker word);
* ``events[p]`` -- per chain: buffer, register slot, group, last-chain-of-the-group (write the
  result), mbarrier parity;
* ``table[c]`` -- the split-TF32 weight blocks of ``ops[c]`` as the byte image of the kernel's
  shared-memory stages (4 ops = 32 entries per stage, rows [hi | lo], 128-byte swizzle applied
  on the host so that the kernel issues one contiguous bulk copy per stage).

The accumulate of the tensor core truncates, so chains are cut after ``CHAIN``
ops and summed in float32 registers by the drain warps, as in K6 / K7.  Every buffer is used an
even number of times per segment (an empty chain is added otherwise): all parities are static.

Boxes whose groups span more than ``W_LIVE`` are visited once per window (generic stacks);
segments (contiguous runs of the visit order, at least ``W_LIVE`` groups apart) are independent
work items, so a group receives partial sums from at most two of them.
"""
import numpy as np

W_LIVE = 4        # groups with an open accumulator (window of consecutive group ids)
NBUF = 5          # TMEM accumulator buffers of 64 columns: ONE pool (4 live groups + 1 in drain)
CHAIN = 8         # ops per accumulation chain
A_STAGES = 3      # A-operand stages (boxes in flight between converters and MMA issuers)
BOX = 32          # pixels per TMA box (128 bytes)
SL = 8            # pixels per slice (K of one tf32 MMA)
HR = 56           # weight rows per half ([hi | lo]): 2 * group_size <= HR
STAGE_OPS = 4     # ops per table stage (32 entries)
STAGE_ROWS = 2 * HR
STAGE_FLOATS = STAGE_ROWS * 32
MAX_SEGMENTS = 8

OP_FIRST = 1 << 3
OP_COMMIT = 1 << 4
OP_NEW_BOX = 1 << 5     # first word of the pipeline in its box: wait for the converters
OP_END_BOX = 1 << 6     # last word of the pipeline in its box: release the A stage
OP_NOMMA = 1 << 7       # marker / padding: no tensor work
OP_SLICE_SHIFT = 8      # bits 8-9: slice of the box
OP_PARITY_SHIFT = 10    # first op of a chain: mbarrier parity of the issuer's wait for the drain
#                         of the buffer's previous chain
OP_ASTAGE_SHIFT = 11    # bits 11-12: A stage of the box (box index % A_STAGES)
OP_APARITY_SHIFT = 13   # mbarrier parity of the A stage ((box index // A_STAGES) & 1)
OP_NOP = OP_NOMMA       # padding word

EV_SLOT = 1 << 3
EV_LAST = 1 << 4
EV_PARITY_SHIFT = 5     # parity of the buffer's use count in the segment
EV_NEXT_SHIFT = 6       # bits 6-7: which issuer (g % 4) uses the buffer next
BOX_PAD = 2 * A_STAGES  # visits per segment are padded to a multiple of this

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


def build_walk(flat, group_size, n_segments=None, chain=None, max_dup=1.5, max_per_slice=None):
    """flat: (M, K) complex stack, M = n_groups * group_size.  Returns the plan as a dict of
    numpy arrays or None when the stack does not fit this form: K % 32, group_size > HR / 2, or
    boxes that would have to be visited more than ``max_dup`` times on average (a box whose
    groups span more than W_LIVE is visited once per window: groups much narrower than a quarter
    of a box are better served by the gathered plan of K7).  ``max_per_slice`` optionally bounds
    the groups on one 8-pixel slice."""
    if chain is None:
        import os
        chain = int(os.environ.get('LTB200_K10_CHAIN', 0)) or CHAIN
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
    if max_per_slice is not None and int(sl.sum(axis=0).max()) > max_per_slice:
        return None
    sl4 = sl.reshape(n_groups, K // BOX, BOX // SL)
    n_ops_visit = np.zeros(n_visits, dtype=np.int64)
    for v in range(n_visits):
        n_ops_visit[v] = sl4[lo[v]:hi[v] + 1, box[v]].sum()
    if n_segments is None:
        import os
        n_segments = int(os.environ.get('LTB200_K10_SEGS', 0)) or \
            max(1, min(MAX_SEGMENTS, n_groups // W_LIVE))
    bounds = _segment_bounds(key, n_ops_visit, n_segments)
    n_seg = len(bounds) - 1

    boxes = []
    ops = ([], [], [], [])         # per issuer (g % 4): its words
    op_group = ([], [], [], [])    # group of every word (-1 marker / padding, <= -2 empty chain)
    op_seq = ([], [], [], [])      # position of every op in the walk order (emulator only)
    slot_slice = ([], [], [], [])  # per table slot of the issuer: global slice index (-1: pad)
    slot_group = ([], [], [], [])  # per table slot: group (< 0: zero block)
    events = ([], [])              # per pipeline (= drain group)
    visit_off, ev_off = [0], ([0], [0])
    op_off, tab_off = ([0], [0], [0], [0]), ([0], [0], [0], [0])
    for s in range(n_seg):
        v0, v1 = bounds[s], bounds[s + 1]
        seq = []                                   # [box index in the segment, slice, group]
        seg_boxes = []
        for v in range(v0, v1):
            m = sl4[lo[v]:hi[v] + 1, box[v]]       # (groups in window, 4)
            mask = 0
            for j in range(BOX // SL):
                gs = np.nonzero(m[:, j])[0]
                if len(gs) == 0:
                    continue
                mask |= 1 << j
                for g in gs:
                    seq.append([v - v0, j, int(lo[v] + g)])
            seg_boxes.append((int(box[v]) * BOX) | mask)
        while len(seg_boxes) % BOX_PAD:            # empty visits: the box is loaded, not used
            seg_boxes.append(seg_boxes[-1] & ~15)
        last_op_of_group = {}
        for i, e in enumerate(seq):
            last_op_of_group[e[2]] = i
        # accumulator buffers: ONE pool for both pipelines, FIFO = oldest released first (TMEM
        # holds 5 buffers next to 3 A stages; 4 groups are live, so one buffer is in drain).
        # Any of the four issuers may use a buffer next and they may be a (short) chain apart,
        # so the "buffer drained" barriers are per buffer AND waiting issuer: the drain of a
        # chain arrives on the barrier of the issuer that uses the buffer NEXT (a static fact),
        # and each barrier is waited on by one thread in a fixed order -- no parity can alias.
        # For the same reason "chain complete" has one barrier per buffer and drain group.
        free = list(range(NBUF))
        uses_p = [[0, 0] for _ in range(NBUF)]     # chains per buffer and pipeline (drain group)
        chains = [[] for _ in range(NBUF)]         # per buffer: [issuer, pipeline, event, word]
        open_buf, open_len = {}, {}
        words = ([], [])                           # per pipeline: [word, box, slice, group, seq]
        n_seq = [0]

        def empty_chain(b, cls, b_i, j):
            """a chain without weights on buffer b, issued by issuer cls"""
            par = cls & 1
            w = [(j << OP_SLICE_SHIFT) | b | OP_FIRST | OP_COMMIT, b_i, j, -2 - cls, n_seq[0]]
            n_seq[0] += 1
            words[par].append(w)
            events[par].append(b | ((uses_p[b][par] & 1) << EV_PARITY_SHIFT) | (par << 8))
            chains[b].append([cls, par, len(events[par]) - 1, w])
            uses_p[b][par] += 1

        if seq:
            # every segment starts with an empty chain of issuer 0 on every buffer: the last
            # drain of an item can then always be addressed to issuer 0
            for b in range(NBUF):
                empty_chain(b, 0, seq[0][0], seq[0][1])
        for i, (b_i, j, g) in enumerate(seq):
            par = g & 1
            w = [j << OP_SLICE_SHIFT, b_i, j, g, n_seq[0]]
            n_seq[0] += 1
            if g not in open_buf:
                if not free:
                    raise AssertionError('walk plan: accumulator pool exhausted')
                b = open_buf[g] = free.pop(0)
                open_len[g] = 0
                w[0] |= OP_FIRST
                chains[b].append([g & 3, par, None, w])
                if len(open_buf) > W_LIVE:
                    raise AssertionError('walk plan: more than W_LIVE live groups')
            b = open_buf[g]
            w[0] |= b
            open_len[g] += 1
            final = last_op_of_group[g] == i
            if open_len[g] == chain or final:
                w[0] |= OP_COMMIT
                events[par].append(b | (EV_SLOT if (g >> 1) & 1 else 0) |
                                   (EV_LAST if final else 0) |
                                   ((uses_p[b][par] & 1) << EV_PARITY_SHIFT) | (g << 8))
                chains[b][-1][2] = len(events[par]) - 1
                uses_p[b][par] += 1
                free.append(b)
                del open_buf[g], open_len[g]
            words[par].append(w)
        assert not open_buf
        if seq:
            # every buffer is used an even number of times per segment by EVERY issuer (static
            # mbarrier parities): odd counts get one more empty chain
            b_i, j = seq[-1][0], seq[-1][1]
            for b in range(NBUF):
                for cls in range(4):
                    if sum(1 for c in chains[b] if c[0] == cls) & 1:
                        empty_chain(b, cls, b_i, j)
            for b in range(NBUF):
                n_before = [0, 0, 0, 0]
                for n, (cls, par, ev, w) in enumerate(chains[b]):
                    # the issuer's wait for "drained": the parity of the phase of ITS barrier
                    # that the drain before this chain completes.  Issuer 0 consumes one
                    # arrival that precedes the segment (the last drain of the previous item;
                    # none in the first item, where the fresh barrier lets parity 1 pass)
                    w[0] |= ((n_before[cls] & 1) ^ (1 if cls == 0 else 0)) << OP_PARITY_SHIFT
                    n_before[cls] += 1
                    nxt = chains[b][n + 1][0] if n + 1 < len(chains[b]) else 0
                    events[par][ev] |= nxt << EV_NEXT_SHIFT
                assert chains[b][0][0] == 0 and not any(v & 1 for v in n_before)
                assert not uses_p[b][0] & 1 and not uses_p[b][1] & 1
        # per ISSUER (groups g % 4; two per pipeline): its own word list -- its ops in walk
        # order, one marker word in a box without own ops -- with its box hand-over flags (wait
        # for the converters before its first word of a box, release the A stage after its
        # last one), and its own weight-table stream (the blocks of its ops, 4 per stage;
        # markers take no table slot; a segment starts on a new stage)
        for cls in range(4):
            by_box = {}
            for w in words[cls & 1]:
                if w[3] >= 0 and (w[3] & 3) == cls or w[3] == -2 - cls:
                    by_box.setdefault(w[1], []).append(w)
            for b_i in range(len(seg_boxes)):
                ws = by_box.get(b_i) or [[OP_NOMMA, b_i, 0, -1, -1]]
                stage_bits = (((b_i % A_STAGES) << OP_ASTAGE_SHIFT) |
                              (((b_i // A_STAGES) & 1) << OP_APARITY_SHIFT))
                for n, w in enumerate(ws):
                    word = w[0] | stage_bits
                    if n == 0:
                        word |= OP_NEW_BOX
                    if n == len(ws) - 1:
                        word |= OP_END_BOX
                    ops[cls].append(word)
                    op_seq[cls].append(w[4])
                    op_group[cls].append(w[3])
                    if w[3] != -1:
                        slot_slice[cls].append((seg_boxes[b_i] & ~31) // SL + w[2])
                        slot_group[cls].append(w[3])
            while len(ops[cls]) % STAGE_OPS:           # whole rounds of 4 words
                ops[cls].append(OP_NOP)
                op_seq[cls].append(-1)
                op_group[cls].append(-1)
            while len(slot_slice[cls]) % STAGE_OPS:    # the last table stage of the segment
                slot_slice[cls].append(-1)
                slot_group[cls].append(-1)
            op_off[cls].append(len(ops[cls]))
            tab_off[cls].append(len(slot_slice[cls]) // STAGE_OPS)
        for par in range(2):
            ev_off[par].append(len(events[par]))
        boxes.extend(seg_boxes)
        visit_off.append(len(boxes))

    plan = dict(
        n_groups=n_groups, group_size=group_size, sig_size=K, n_segments=n_seg, chain=chain,
        boxes=np.array(boxes, dtype=np.uint32),
        visit_off=np.array(visit_off, dtype=np.int32),
        n_real_ops=sum(int((np.array(slot_group[c]) >= 0).sum()) for c in range(4)),
        n_entries=sum(len(slot_slice[c]) for c in range(4)) * SL,
    )
    for p in range(2):
        plan[f'events{p}'] = np.array(events[p], dtype=np.uint32)
        plan[f'ev_off{p}'] = np.array(ev_off[p], dtype=np.int32)
    for c in range(4):
        sl_c = np.array(slot_slice[c], dtype=np.int64)
        gr_c = np.array(slot_group[c], dtype=np.int64)
        plan[f'ops{c}'] = np.array(ops[c], dtype=np.uint32)
        plan[f'op_off{c}'] = np.array(op_off[c], dtype=np.int32)
        plan[f'tab_off{c}'] = np.array(tab_off[c], dtype=np.int32)
        plan[f'op_group{c}'] = np.array(op_group[c], dtype=np.int64)
        plan[f'op_seq{c}'] = np.array(op_seq[c], dtype=np.int64)
        plan[f'slot_slice{c}'] = sl_c
        plan[f'slot_group{c}'] = gr_c
        plan[f'table{c}'] = _table_image(flat, group_size, sl_c, gr_c)
    return plan


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
    invariants the kernel relies on).  The four issuers are replayed in the walk order of the
    plan (any interleaving that respects the mbarrier waits gives the same result).
    tile: (F, K) -> (F, n_groups * group_size) complex128"""
    F = tile.shape[0]
    G, gs = plan['n_groups'], plan['group_size']
    out = np.zeros((F, G, 2 * gs), dtype=np.float64)
    written = np.zeros(G, dtype=np.int64)
    w = []
    for c in range(4):
        t = unswizzle_table(plan[f'table{c}']).astype(np.float64)
        w.append(t[:, :HR] + t[:, HR:])                              # (stages, HR, 4, 8)
    stage_mask = (3 << OP_ASTAGE_SHIFT) | (1 << OP_APARITY_SHIFT)
    for s in range(plan['n_segments']):
        v0, v1 = plan['visit_off'][s], plan['visit_off'][s + 1]
        assert (v1 - v0) % BOX_PAD == 0
        bufs = np.zeros((NBUF, F, HR))
        busy = [False] * NBUF
        uses_p = [[0, 0] for _ in range(NBUF)]     # chains per buffer and drain group
        uses_c = [[0] * 4 for _ in range(NBUF)]    # chains per buffer and issuer
        next_user = [0] * NBUF                     # an item starts with issuer 0 everywhere
        acc = np.zeros((2, 2, F, HR))
        slot_owner = [[None, None], [None, None]]
        evs = [list(plan[f'events{p}'][plan[f'ev_off{p}'][s]:plan[f'ev_off{p}'][s + 1]])
               for p in range(2)]
        open_groups = {}
        order = []
        for c in range(4):
            o0, o1 = plan[f'op_off{c}'][s], plan[f'op_off{c}'][s + 1]
            assert o0 % STAGE_OPS == 0 and o1 % STAGE_OPS == 0
            b_i, in_box = -1, False
            slot = plan[f'tab_off{c}'][s] * STAGE_OPS      # table slot of the next op
            for i in range(o0, o1):
                word = int(plan[f'ops{c}'][i])
                if word == OP_NOP:
                    assert not in_box and plan[f'op_group{c}'][i] == -1
                    continue                           # padding of the last round
                if word & OP_NEW_BOX:
                    assert not in_box
                    b_i += 1
                    in_box = True
                assert in_box
                assert ((word >> OP_ASTAGE_SHIFT) & 3) == b_i % A_STAGES
                assert ((word >> OP_APARITY_SHIFT) & 1) == ((b_i // A_STAGES) & 1)
                if not word & OP_NOMMA:
                    order.append((int(plan[f'op_seq{c}'][i]), c, i, b_i, slot))
                    slot += 1
                else:
                    assert (word & ~stage_mask) == OP_NOMMA | OP_NEW_BOX | OP_END_BOX
                if word & OP_END_BOX:
                    in_box = False
            assert b_i == v1 - v0 - 1 and not in_box       # every issuer sees every box
            assert -(-slot // STAGE_OPS) == plan[f'tab_off{c}'][s + 1]
        order.sort()
        last_seq = [-1] * 4
        for seq_no, c, i, b_i, slot in order:
            assert seq_no > last_seq[c]                # an issuer keeps the walk order
            last_seq[c] = seq_no
            p = c & 1
            word = int(plan[f'ops{c}'][i])
            box_word = int(plan['boxes'][v0 + b_i])
            px0, mask = box_word & ~31, box_word & 15
            g = int(plan[f'op_group{c}'][i])
            assert ((g & 3) if g >= 0 else -2 - g) == c
            j = (word >> OP_SLICE_SHIFT) & 3
            assert mask >> j & 1                       # the converters fill only masked slices
            assert plan[f'slot_slice{c}'][slot] * SL == px0 + j * SL
            assert plan[f'slot_group{c}'][slot] == g
            b = word & 7
            assert b < NBUF
            x = tile[:, px0 + j * SL:px0 + (j + 1) * SL].astype(np.float64)
            prod = x @ w[c][slot // STAGE_OPS, :, slot % STAGE_OPS].T              # (F, HR)
            if g <= -2:                                # empty chain
                assert word & OP_FIRST and word & OP_COMMIT and not busy[b]
                assert not np.any(prod)
            if word & OP_FIRST:
                assert not busy[b] and g not in open_groups
                assert ((word >> OP_PARITY_SHIFT) & 1) == \
                    (uses_c[b][c] & 1) ^ (1 if c == 0 else 0)
                assert next_user[b] == c               # the last drain was addressed to it
                uses_c[b][c] += 1
                busy[b] = True
                open_groups[g] = b
                bufs[b] = prod
            else:
                assert busy[b] and open_groups[g] == b
                bufs[b] += prod
            live = [q for q in open_groups if q >= 0]
            assert len(live) <= W_LIVE and (not live or max(live) - min(live) < W_LIVE)
            if word & OP_COMMIT:
                ev = int(evs[p].pop(0))
                assert (ev & 7) == b
                assert ((ev >> EV_PARITY_SHIFT) & 1) == (uses_p[b][p] & 1)
                next_user[b] = (ev >> EV_NEXT_SHIFT) & 3
                uses_p[b][p] += 1
                busy[b] = False
                del open_groups[g]
                if g <= -2:
                    assert not ev & EV_LAST
                else:
                    assert (ev >> 8) == g
                    slot_r = 1 if ev & EV_SLOT else 0
                    assert slot_r == ((g >> 1) & 1)
                    assert slot_owner[p][slot_r] in (None, g)
                    slot_owner[p][slot_r] = g
                    acc[p, slot_r] += bufs[b]
                    if ev & EV_LAST:
                        out[:, g] += acc[p, slot_r][:, :2 * gs]
                        written[g] += 1
                        acc[p, slot_r] = 0
                        slot_owner[p][slot_r] = None
        assert not any(busy) and not evs[0] and not evs[1] and not open_groups
        assert slot_owner == [[None, None], [None, None]]
        assert not any(v & 1 for u in uses_p for v in u) and not any(v & 1 for u in uses_c for v in u)
        assert all(n == 0 for n in next_user)
    assert written.max() <= 2
    return (out[..., 0::2] + 1j * out[..., 1::2]).reshape(F, G * gs)
