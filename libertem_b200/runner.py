"""UDFRunner: partitions -> tiles -> process_tile -> merge -> get_results, on one GPU per rank.

Mirrors the ordering of the reference's ``UDFRunner`` / ``UDFPartRunner``
(src/libertem/udf/base.py:2100-2145 run_for_partition, :2147-2206 tile loop, :2256-2309
_run_tile, :2311-2335 wrap-up, :2340-2358 merge, :2360-2386 results) with two B200-first
differences:

* **fused pass** -- the reference loops UDFs per tile, so ApplyMasksUDF, CoMUDF, SumSigUDF and
  SumUDF each re-read the tile (base.py:2187-2194).  Here every UDF that publishes a
  ``_fused_spec()`` contributes mask rows to ONE launch of the dense kernel per tile
  (<= 24 columns per pass over HBM), whose columns are then scattered into the UDFs' own
  result-buffer views; merge / get_results run unchanged.  UDFs without a spec fall back to
  their ``process_tile``.
* **ranks instead of workers** -- with ``torch.distributed`` initialised, rank r processes the
  contiguous block of partitions r of world_size and the dataset-sized buffers are assembled
  with one all-gather (nav) / all-reduce (sig) at the end; partial results never
  leave the device before that.
"""
import os

import numpy as np
import torch

from .common.buffers import BufferWrapper, torch_dtype, to_numpy, _NP2TORCH
from .common.shape import Shape
from .udf.base import UDFMeta, UDFData, MergeAttrMapping, UDFException
from .udf.base import UDF as _BaseUDF
from .udf.sumsigudf import ones_row
from . import engine

MAX_FUSED_COLUMNS = 24
MAX_FUSED_COLUMNS_F32 = 32
K8_MAX_ROWS = 32             # int8 rows per pass of the integer tensor-core kernel (K8)
# uint16 tiles x integer-valued masks -> exact int8 tensor-core kernel (LTB200_INT8=0: off)
INT8_PATH = os.environ.get('LTB200_INT8', '1') != '0'


def _get_dtype(udfs, dtype, corrections=None):
    """input dtype of the run: result_type over every UDF's preferred dtype
    (reference udf/base.py:106-123)"""
    if corrections is not None and corrections.have_corrections():
        tmp = np.result_type(np.float32, dtype)
    else:
        tmp = np.dtype(dtype)
    for udf in udfs:
        pref = udf.get_preferred_input_dtype()
        if pref is bool or pref is udf.USE_NATIVE_DTYPE:
            continue
        tmp = np.result_type(pref, tmp)
    return tmp


#: fixed-point digits (base 128) of a non-integer mask weight: 4 digits = 28 bits relative to the
#: power of two above the row's largest weight, i.e. |m - q s| <= 2^-27 max|m| -- below float32
#: resolution for every weight within a factor 8 of the largest one
FLOAT_MASK_DIGITS = 4
FLOAT_MASKS_INT8 = os.environ.get('LTB200_FLOAT_MASKS_INT8', '1') != '0'


def int8_digit_plan(rows, max_rows=16, float_digits=FLOAT_MASK_DIGITS):
    """int8 form of a float32 mask stack (M, K) for the integer tensor-core kernel K8, or None:
    ``(int8 rows (R, K), combine)`` with ``masks = combine @ int8 rows`` -- ``combine`` is None
    when the int8 rows ARE the masks, else a float64 (M, R) matrix.

    * integer weights with ``|m| <= 127``: the row itself;
    * integer weights up to ``127 * 129`` (CoM coordinate masks of detectors up to 16k wide): two
      base-128 digits ``m = d0 + 128 d1`` (both in [-127, 127], ``d1 = trunc(m / 128)``) -- exact;
    * any other float32 row (``float_digits > 0``): fixed point ``m ~ s * q`` with the power of two
      ``s = 2^(e - 27)``, ``2^e >= max|m|``, ``q = round(m / s)`` split into four balanced
      base-128 digits.  The integer sums of K8 are exact, so the only error of the pass is the
      quantisation ``|m - s q| <= s / 2 <= 2^-27 max|m|`` per weight (bound on a result:
      ``2^-27 max|m| sum|x|``; the reference's float32 GEMM carries ~1e-7 sum|x||m|).
    None when the digits do not fit in ``max_rows`` rows."""
    M = rows.shape[0]
    amax = rows.abs().amax(dim=1)
    is_int = (rows == rows.round()).all(dim=1) & (amax <= 127 * 129)
    if bool(is_int.all()) and float(amax.max()) <= 127:
        return (rows.to(torch.int8).contiguous(), None)
    if not bool(is_int.all()) and (float_digits <= 0 or not FLOAT_MASKS_INT8):
        return None
    if not bool(torch.isfinite(rows).all()):
        return None
    pieces, combine = [], []
    for c in range(M):
        r = rows[c]
        if bool(is_int[c]):
            if float(amax[c]) <= 127:
                pieces.append(r[None])
                combine.append((c, 1.0))
            else:
                d1 = torch.trunc(r / 128.0)
                pieces.append(torch.stack([r - 128.0 * d1, d1]))
                combine += [(c, 1.0), (c, 128.0)]
            continue
        e = int(np.ceil(np.log2(float(amax[c]))))
        if float(amax[c]) > 2.0 ** e:          # guard against log2 rounding
            e += 1
        shift = 7 * float_digits - 1 - e        # q = m * 2^shift, |q| <= 2^(7 D - 1)
        q = torch.round(r.double() * (2.0 ** shift)).to(torch.int64)
        digits = []
        for _ in range(float_digits - 1):
            d = torch.remainder(q + 64, 128) - 64
            digits.append(d)
            q = (q - d) // 128
        digits.append(q)                        # top digit: |q| <= 64
        pieces.append(torch.stack(digits).to(torch.float32))
        combine += [(c, 2.0 ** -shift * 128.0 ** j) for j in range(float_digits)]
    if len(combine) > max_rows:
        return None
    i8 = torch.cat(pieces).to(torch.int8).contiguous()
    C = torch.zeros((M, len(combine)), dtype=torch.float64, device=rows.device)
    for j, (c, w) in enumerate(combine):
        C[c, j] = w
    return (i8, C)


class UDFResults:
    def __init__(self, buffers, damage, pending=None):
        self.buffers = buffers
        self.damage = damage
        self._pending = pending or []     # (collective work handle, tensors kept alive)

    def wait(self):
        """``async_merge`` runs: make the current stream wait for this run's collectives"""
        for work, _keep in self._pending:
            work.wait()
        self._pending = []
        return self


class ResultBuffer(BufferWrapper):
    """what run_udf hands to the user: numpy-backed, with ``.data`` / ``.raw_data``"""

    @classmethod
    def wrap(cls, decl, arr, ds_shape, roi):
        buf = cls(decl.kind, decl.extra_shape, decl.dtype, None, decl.use)
        buf.set_shape_ds(ds_shape, roi)
        buf.replace_array(arr)
        return buf

    @property
    def raw_data(self):
        return to_numpy(self._data)


class UDFRunner:
    def __init__(self, udfs, debug=False, fuse=True, rank_weights=None):
        self._udfs = list(udfs)
        self._debug = debug
        self._fuse = fuse
        #: multi-rank runs: relative share of the partitions every rank takes (default: equal).
        #: Host-resident data is bound by each GPU's host link, and the links of one box are not
        #: always equal (bench.py measures 23 vs 35 GB/s per GPU with 8 concurrent copies): a
        #: share proportional to the link rate lets every rank finish at the same time.
        self._rank_weights = None if rank_weights is None else [float(w) for w in rank_weights]
        self.stats = {'tiles': 0, 'fused_launch_groups': 0, 'unfused_calls': 0}
        self._cat_cache = {}
        self._int8_cache = {}
        self._slab = None
        self._slab_map = {}
        self._corr = None
        self._corr_fold = False
        self._fold_cache = {}
        self._slice_corr = {}
        self._sig_sum_folded = []

    # -- distributed helpers ---------------------------------------------------------------------
    @staticmethod
    def _dist():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
        return None

    @staticmethod
    def my_partitions(partitions, rank, world, weights=None):
        """contiguous block of partitions per rank; with ``weights`` the block sizes
        are proportional to them (every rank computes the same boundaries)"""
        n = len(partitions)
        if weights is None:
            b = np.linspace(0, n, world + 1, dtype=int)
        else:
            w = np.asarray(weights, dtype=np.float64)
            if len(w) != world or not np.all(w > 0):
                raise UDFException('rank_weights: one positive weight per rank')
            b = np.concatenate([[0], np.rint(np.cumsum(w) / w.sum() * n)]).astype(int)
        return partitions[b[rank]:b[rank + 1]]

    # -- main entry -------------------------------------------------------------------------------
    def run_for_dataset(self, dataset, executor=None, roi=None, progress=False,
                        corrections=None, backends=None, dry=False, device=None,
                        finalize=True, use_merge_all=False, async_merge=False):
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        device = torch.device(device)
        udfs = self._udfs
        ds_shape = dataset.shape
        n_frames = ds_shape.nav.size
        if roi is not None:
            roi = np.asarray(roi)
            if roi.dtype != bool or roi.size != n_frames:
                raise UDFException('roi must be a boolean array of the navigation shape')
            roi = roi.reshape(tuple(ds_shape.nav))
        roi_flat = None if roi is None else roi.reshape(-1)
        if corrections is not None and not corrections.have_corrections():
            corrections = None
        input_dtype = _get_dtype(udfs, dataset.dtype, corrections)
        self._corr = corrections
        # caches derived from the CorrectionSet of a previous run must not leak into this one
        self._fold_cache = {}
        self._slice_corr = {}
        # corrections fold into the masks (libertem_b200/corrections.py) when every tile is a
        # full frame; sub-frame tilings get explicitly corrected tiles like in the reference
        self._corr_fold = corrections is not None and getattr(dataset, 'tileshape', None) is None
        self._sig_sum_folded = []

        # dataset-level instances + buffers (base.py:2472-2557)
        slab_cols = []      # (udf index, buffer name, n real columns)
        for ui, udf in enumerate(udfs):
            udf.set_meta(UDFMeta(partition_slice=None, dataset_shape=ds_shape, roi=roi,
                                 dataset_dtype=dataset.dtype, input_dtype=input_dtype,
                                 device=device))
            decl = udf.get_result_buffers()
            slab_name = self._slab_buffer_name(udf, decl) if self._fuse else None
            for name, buf in decl.items():
                buf.set_shape_ds(ds_shape, roi)
                if name == slab_name:
                    slab_cols.append((ui, name, int(np.prod(buf.extra_shape, dtype=np.int64))))
                elif buf.use != 'result_only':
                    buf.allocate(device)
            udf.results = UDFData(decl)
        # fused result slab: every float32 nav buffer the dense kernel can write lives in ONE
        # (rows, total columns) device array; the buffers are column-slice views of it.  The
        # kernel stores straight into it (ld_out = total columns), partition buffers are row
        # views (merge is a no-op) and the multi-rank merge is a single all-gather.
        self._slab = None
        self._slab_map = {}
        if slab_cols:
            n_rows = n_frames if roi is None else int(np.count_nonzero(roi))
            total = sum(c for _, _, c in slab_cols)
            self._slab = torch.zeros((n_rows, total), dtype=torch.float32, device=device)
            c0 = 0
            for ui, name, c in slab_cols:
                buf = udfs[ui].results.get_buffer(name)
                view = self._slab[:, c0:c0 + c]
                buf.replace_array(view if buf.extra_shape else view[:, 0])
                self._slab_map[(ui, name)] = (c0, c)
                c0 += c

        partitions = list(dataset.get_partitions())
        dist = self._dist()
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
        mine = self.my_partitions(partitions, rank, world, self._rank_weights)
        damage = np.zeros(n_frames if roi_flat is None else int(roi_flat.sum()), dtype=bool)

        self._input_tdtype = torch_dtype(input_dtype) if np.dtype(input_dtype) in _NP2TORCH \
            else None
        if not dry:
            collected = [dict() for _ in udfs]       # per UDF: {partition Slice: results}
            for part in mine:
                part_udfs = self._run_partition(part, dataset, udfs, roi_flat, input_dtype,
                                                device)
                self._merge_partition(part, udfs, part_udfs, roi_flat, damage,
                                      collect=collected if use_merge_all else None)
            if use_merge_all:
                # the reference's merge_all contract (udf/base.py:944-1002, used by
                # executor/delayed.py:81-88): all partition results of a UDF at once, in order
                for ui, (udf, parts_) in enumerate(zip(udfs, collected)):
                    if parts_:
                        udf._do_merge_all(parts_)
        pending = []
        if dist:
            # async_merge (only with finalize=False, i.e. results stay on the device): the
            # collectives are issued asynchronously -- they wait for this run's kernels but the
            # launching stream does not wait for them, so the next run's kernels overlap them;
            # the caller completes them with UDFResults.wait()
            self._merge_ranks(dist, udfs, partitions, roi_flat, damage, device,
                              pending=pending if (async_merge and not finalize) else None)
        if self._corr is not None and self._sig_sum_folded:
            n_done = int(damage.sum())
            for udf, name in self._sig_sum_folded:
                buf = udf.results.get_buffer(name)
                fixed = self._corr.correct_frame_sum(to_numpy(buf.tensor), n_done)
                buf.tensor.copy_(torch.from_numpy(fixed.astype(np.float32)).reshape(
                    buf.tensor.shape))
        if not finalize:
            # hot path only (tiles -> kernels -> merge [-> collectives]); results stay in
            # the UDFs' device buffers (udf.results)
            return UDFResults(buffers=None, damage=damage, pending=pending)
        return UDFResults(buffers=self._make_results(udfs, ds_shape, roi, damage), damage=damage)

    # -- per partition ------------------------------------------------------------------------------
    def _roi_range(self, part, roi_flat):
        if roi_flat is None:
            return part.start, part.stop
        a = int(roi_flat[:part.start].sum())
        return a, a + int(roi_flat[part.start:part.stop].sum())

    def _run_partition(self, part, dataset, udfs, roi_flat, input_dtype, device):
        ds_shape = dataset.shape
        r0, r1 = self._roi_range(part, roi_flat)
        n_part = r1 - r0
        part_udfs = []
        for ui, udf in enumerate(udfs):
            pu = udf.copy_for_partition()
            pu.set_meta(UDFMeta(partition_slice=part.slice, dataset_shape=ds_shape,
                                roi=udf.meta.roi, dataset_dtype=dataset.dtype,
                                input_dtype=input_dtype, device=device))
            decl = pu.get_result_buffers()
            for name, buf in decl.items():
                buf.set_shape_partition(ds_shape, n_part)
                if (ui, name) in self._slab_map:
                    # zero-copy: the partition buffer IS its row range of the dataset slab
                    buf.replace_array(udf.results.get_buffer(name).tensor[r0:r1])
                    pu._slab_cols = self._slab_map[(ui, name)]
                elif buf.use != 'result_only':
                    buf.allocate(device)
            pu.results = UDFData(decl)
            pu.task_data = pu.get_task_data()
            part_udfs.append(pu)
        if n_part == 0:
            return part_udfs
        for pu in part_udfs:
            pu.preprocess()

        specs = [self._spec(pu) for pu in part_udfs] if self._fuse else [None] * len(part_udfs)
        for tile, f0, f1, tslice in part.get_tiles(device, roi=roi_flat):
            self.stats['tiles'] += 1
            t0 = tslice.origin[0] - r0           # first row of this tile in partition buffers
            t1 = t0 + tslice.shape[0]
            for pu in part_udfs:
                pu.meta.slice = tslice
                pu.results.clear_views()
                for name, buf in pu.results.items():
                    if buf.has_data():
                        pu.results.set_view(name, buf.rows(t0, t1))
            self._run_tile(part_udfs, specs, tile, tslice, ds_shape, device,
                           t0_rows=(tslice.origin[0], tslice.origin[0] + tslice.shape[0]))
        for pu in part_udfs:
            pu.results.clear_views()
            pu.meta.slice = None
            pu.postprocess()
        return part_udfs

    @staticmethod
    def _spec(pu):
        fn = getattr(pu, '_fused_spec', None)
        if fn is None:
            return None
        if getattr(pu, 'get_method', lambda: 'tile')() == 'frame':
            return None
        return fn()

    def _folded(self, rows, fold):
        """(rows', constant) with the corrections folded into the mask rows, cached"""
        if not fold:
            return (rows, None)
        key = (rows.data_ptr(), rows.shape[0], id(self._corr))
        hit = self._fold_cache.get(key)
        if hit is None:
            r64, const = self._corr.fold_masks(rows.detach().cpu().numpy())
            hit = (torch.from_numpy(r64.astype(np.float32)).to(rows.device),
                   torch.from_numpy(const.astype(np.float32)).to(rows.device), rows)
            self._fold_cache[key] = hit
        return (hit[0], hit[1])

    def _corr_for_slice(self, sig_slice):
        """CorrectionSet restricted to a sub-frame tile: repair environments are clipped to the
        tile, as the reference does per tile (corrset.py:171-180)"""
        from .corrections import CorrectionSet
        key = sig_slice
        hit = self._slice_corr.get(key)
        if hit is None:
            sl = sig_slice.get()
            c = self._corr
            hit = CorrectionSet(
                dark=None if c.get_dark_frame() is None else c.get_dark_frame()[sl],
                gain=None if c.get_gain_map() is None else c.get_gain_map()[sl],
                excluded_pixels=None if c.get_excluded_pixels() is None
                else c.get_excluded_pixels()[sl], allow_empty=True)
            self._slice_corr[key] = hit
        return hit

    def _slab_target(self, grp, t0_rows):
        """(rows, columns) window of the slab that a fused group writes, when every member's
        buffer is a slab member and their columns are consecutive in group order"""
        if self._slab is None or not grp:
            return None
        c_first = c_next = None
        for pu, spec, rows, _const in grp:
            cols = getattr(pu, '_slab_cols', None)
            if cols is None or cols[1] != rows.shape[0]:
                return None
            if c_first is None:
                c_first = c_next = cols[0]
            if cols[0] != c_next:
                return None
            c_next = cols[0] + cols[1]
        r0, r1 = t0_rows
        return self._slab[r0:r1, c_first:c_next]

    @staticmethod
    def _slab_buffer_name(udf, decl):
        """name of the float32 nav buffer a hot-path UDF lets the dense kernel write, if any"""
        if getattr(udf, '_fused_spec', None) is None:
            return None
        if type(udf).merge is not _BaseUDF.merge:
            return None
        name = getattr(udf, '_slab_buffer', None)
        if name is None or name not in decl:
            return None
        buf = decl[name]
        if buf.kind != 'nav' or buf.dtype != np.float32 or buf.where != 'device':
            return None
        if getattr(udf, 'get_method', lambda: 'tile')() == 'frame':
            return None
        return name

    def _run_tile(self, part_udfs, specs, tile, tslice, ds_shape, device, t0_rows=(0, 0)):
        flat = tile.reshape(tile.shape[0], -1)
        full_frame = tslice.shape.sig.size == ds_shape.sig.size
        sig_slice = tslice.discard_nav()
        fusable_dtype = flat.dtype in (torch.float32, torch.uint16, torch.uint8, torch.int16,
                                       torch.int8)
        dense = []       # (pu, spec, rows tensor[, constant per column])
        sig_sum_view = None
        corr = self._corr
        fold = corr is not None and self._corr_fold and full_frame
        corrected_tile = None

        def explicit_tile():
            # reference behaviour: the UDF sees a corrected float tile (corrset.py:141-169)
            nonlocal corrected_tile
            if corr is None:
                return tile
            if corrected_tile is None:
                cs = corr if full_frame else self._corr_for_slice(sig_slice)
                corrected_tile = cs.apply(tile)
            return corrected_tile

        for pu, spec in zip(part_udfs, specs):
            if spec is None or not fusable_dtype:
                self._run_unfused(pu, explicit_tile())
                continue
            kind = spec['kind']
            if corr is not None and not fold:
                self._run_unfused(pu, explicit_tile())
                continue
            if corr is not None and kind not in ('dense', 'ones', 'sig_sum', 'csc'):
                self._run_unfused(pu, explicit_tile())
                continue
            if kind == 'dense':
                dense.append((pu, spec) + self._folded(spec['engine'].dense_rows(sig_slice),
                                                      fold))
            elif kind == 'ones':
                dense.append((pu, spec) + self._folded(ones_row(flat.shape[1], device), fold))
            elif kind == 'sig_sum':
                if full_frame and sig_sum_view is None:
                    sig_sum_view = getattr(pu.results, spec['buffer']).reshape(-1)
                    if fold:
                        key = (self._udfs[part_udfs.index(pu)], spec['buffer'])
                        if key not in self._sig_sum_folded:
                            self._sig_sum_folded.append(key)
                else:
                    self._run_unfused(pu, explicit_tile())
            elif kind == 'own_pass':
                # group-sparse complex masks (K4): a pass of their own through the engine
                if flat.dtype == torch.float32:
                    view = getattr(pu.results, spec['buffer'])
                    out = self._real_view(view, 2 * view.shape[1])
                    spec['engine'].process_flat(flat, out=out, accumulate=True,
                                                sig_slice=sig_slice)
                    self.stats['fused_launch_groups'] += 1
                else:
                    self._run_unfused(pu, tile)
            elif kind == 'csc':
                eng = spec['engine']
                tma_able = (flat.dtype in (torch.float32, torch.uint16)
                            and flat.shape[1] % 8 == 0 and flat.shape[1] >= 128
                            and flat.shape[0] >= 8)
                if tma_able and len(eng.masks) <= MAX_FUSED_COLUMNS:
                    # the frames are streamed once anyway: a few sparse masks ride along as
                    # dense rows of the fused pass (exact same sums; zeros contribute nothing)
                    dense.append((pu, spec) + self._folded(eng.dense_rows(sig_slice), fold))
                elif corr is not None:
                    self._run_unfused(pu, explicit_tile())
                else:
                    view = getattr(pu.results, spec['buffer'])
                    eng.process_flat(flat, out=view, accumulate=True, sig_slice=sig_slice)
                    self.stats['fused_launch_groups'] += 1
            else:
                self._run_unfused(pu, explicit_tile())
        if not dense and sig_sum_view is None:
            return
        # one pass over the tile per group of <= 24 columns (FFMA2 kernel) or <= 32 columns
        # (float32 tiles: the tensor-core kernel K6 takes 32 per pass)
        max_cols = MAX_FUSED_COLUMNS_F32 if flat.dtype == torch.float32 else MAX_FUSED_COLUMNS
        groups, cur, ncols = [], [], 0
        for item in dense:
            c = item[2].shape[0]
            if cur and ncols + c > max_cols:
                groups.append(cur)
                cur, ncols = [], 0
            cur.append(item)
            ncols += c
        if cur:
            groups.append(cur)
        if not groups:
            groups = [[]]
        for gi, grp in enumerate(groups):
            ss = sig_sum_view if gi == 0 else None
            direct = self._slab_target(grp, t0_rows)
            consts = [g[3] for g in grp]
            if direct is not None:
                rows = self._cat_rows([g[2] for g in grp], flat.shape[1], device)
                self._dense(flat, rows, out=direct, accumulate=True, sig_sum=ss)
                if any(c is not None for c in consts):
                    direct += torch.cat([c if c is not None else
                                         torch.zeros(g[2].shape[0], device=device)
                                         for c, g in zip(consts, grp)])
            elif len(grp) == 1 and ss is None:
                pu, spec, rows, const = grp[0]
                view = getattr(pu.results, spec['buffer'])
                out = self._real_view(view, rows.shape[0])
                self._dense(flat, rows, out=out, accumulate=True)
                if const is not None:
                    out += const
            else:
                rows = self._cat_rows([g[2] for g in grp], flat.shape[1], device)
                res = self._dense(flat, rows, sig_sum=ss)
                c0 = 0
                for pu, spec, r, const in grp:
                    c = r.shape[0]
                    view = getattr(pu.results, spec['buffer'])
                    out = self._real_view(view, c)
                    out += res[:, c0:c0 + c]
                    if const is not None:
                        out += const
                    c0 += c
            self.stats['fused_launch_groups'] += 1

    def _dense(self, flat, rows, out=None, accumulate=False, sig_sum=None):
        """One fused pass.  uint16 / uint8 tiles take the integer tensor-core kernel (K8) whenever
        the mask rows fit its int8 digit rows: integer-valued rows (binary virtual detectors, the
        all-ones row of SumSigUDF, the CoM coordinate masks) exactly, other float32 rows as
        28-bit fixed point (error <= 2^-27 max|m| sum|x|); everything else the float kernels behind ``masks_dense``."""
        plan = self._int8_rows(flat, rows)
        if plan is None:
            return engine.masks_dense(flat, rows, out=out, accumulate=accumulate,
                                      sig_sum=sig_sum)
        i8, combine = plan
        self.stats['int8_passes'] = self.stats.get('int8_passes', 0) + 1
        if combine is None:
            return engine.masks_dense_i8(flat, i8, out=out, accumulate=accumulate,
                                         sig_sum=sig_sum)
        # digit rows (wide integer weights, fixed-point float weights): recombine the exact
        # per-digit sums in float64, round once
        res = engine.masks_dense_i8(flat, i8, sig_sum=sig_sum)
        val = (res.double() @ combine.T).float()
        if out is None:
            return val
        if accumulate:
            out += val
        else:
            out.copy_(val)
        return out

    def _int8_rows(self, flat, rows):
        """int8 form of the mask rows when K8 applies to this tile (``int8_digit_plan``), cached
        per row stack, else None"""
        F, K = flat.shape
        M = rows.shape[0]
        if flat.dtype not in (torch.uint16, torch.uint8):
            return None
        align = 16 // flat.element_size()            # pixels per 16 bytes
        # 1..32 int8 rows per pass (digit rows included); signals beyond 65536 pixels are K-split
        # inside the library so that every int32 accumulator stays exact (<= 4 Mi pixels)
        if (not INT8_PATH or not 1 <= M <= K8_MAX_ROWS or F < 256 or K % align
                or not 64 * align // 2 <= K <= (4 << 20) or flat.stride(1) != 1
                or (F > 1 and flat.stride(0) % align) or flat.data_ptr() % 16
                or rows.dtype != torch.float32):
            return None
        key = (rows.data_ptr(), tuple(rows.shape))
        hit = self._int8_cache.get(key)
        if hit is None:
            hit = (int8_digit_plan(rows, max_rows=K8_MAX_ROWS), rows)   # source kept alive (key)
            self._int8_cache[key] = hit
        return hit[0]

    def _cat_rows(self, row_tensors, k, device):
        """stacked mask rows of one fused group, cached across tiles / partitions / runs"""
        if not row_tensors:
            return torch.empty((0, k), dtype=torch.float32, device=device)
        key = tuple((t.data_ptr(), t.shape[0]) for t in row_tensors)
        hit = self._cat_cache.get(key)
        if hit is None:
            hit = (torch.cat(row_tensors, dim=0), row_tensors)   # keep sources alive
            self._cat_cache[key] = hit
        return hit[0]

    @staticmethod
    def _real_view(view, ncols):
        """(F, ...) result view as real (F, ncols) matrix (complex64 -> interleaved floats)"""
        if view.is_complex():
            return torch.view_as_real(view).reshape(view.shape[0], ncols)
        return view.reshape(view.shape[0], ncols)

    def _run_unfused(self, pu, tile):
        self.stats['unfused_calls'] += 1
        method = pu.get_method()
        # the reference guarantees tile.dtype == meta.input_dtype (udf/base.py:106-123,
        # io/dataset/memory.py:99-108); the built-in UDFs' kernels convert on the fly and keep
        # the native dtype, everything else gets the converted tile
        want = getattr(self, '_input_tdtype', None)
        if (want is not None and tile.dtype != want
                and getattr(pu, '_fused_spec', None) is None):
            tile = tile.to(want)
        if method == 'partition':
            part = pu.meta.partition_slice
            if tile.shape[0] != part.shape[0] and pu.meta.roi is None:
                raise UDFException('process_partition needs the whole partition in one tile '
                                   '(device-resident data or tile_depth >= partition size)')
            pu.process_partition(tile)
        elif method == 'frame':
            shifts = pu.params.get('shifts')
            tslice = pu.meta.slice
            if shifts is not None and hasattr(pu, 'process_tile_shifted'):
                if hasattr(shifts, 'for_frames'):
                    arr = shifts.for_frames(pu.meta.dataset_shape, pu.meta.roi)
                    arr = arr[tslice.origin[0]:tslice.origin[0] + tile.shape[0]]
                else:
                    arr = np.asarray(shifts).reshape(1, 2)
                if pu.process_tile_shifted(tile, arr.astype(np.int64)):
                    return
            views = dict(pu.results._views)
            for i in range(tile.shape[0]):
                pu.results.clear_views()
                for name, v in views.items():
                    buf = pu.results.get_buffer(name)
                    if buf.kind != 'nav':
                        pu.results.set_view(name, v)
                    elif buf.extra_shape:
                        pu.results.set_view(name, v[i])
                    else:
                        # (1,) view like get_view_for_frame (common/buffers.py:792-821), so
                        # ``self.results.x[:] = value`` works
                        pu.results.set_view(name, v[i:i + 1])
                if hasattr(shifts, 'for_frames'):
                    arr = shifts.for_frames(pu.meta.dataset_shape, pu.meta.roi)
                    pu._current_shift = arr[tslice.origin[0] + i].astype(int)
                elif shifts is not None:
                    pu._current_shift = np.asarray(shifts).astype(int)
                pu.process_frame(tile[i])
            for name, v in views.items():
                pu.results.set_view(name, v)
        else:
            pu.process_tile(tile)

    # -- merging ---------------------------------------------------------------------------------
    def _merge_partition(self, part, udfs, part_udfs, roi_flat, damage, collect=None):
        r0, r1 = self._roi_range(part, roi_flat)
        for ui, (udf, pu) in enumerate(zip(udfs, part_udfs)):
            dest, src = {}, {}
            for name, buf in udf.results.items():
                if buf.use == 'result_only' or not buf.has_data():
                    continue
                if (ui, name) in self._slab_map:
                    continue        # partition buffer aliases the dataset slab rows already
                dest[name] = buf.rows(r0, r1)
                src[name] = pu.results.get_buffer(name).tensor
            if not dest:
                continue
            can_merge_all = (getattr(udf, 'merge_all', None) is not None
                             or (type(udf).merge is _BaseUDF.merge
                                 and not udf.requires_custom_merge_all))
            if collect is not None and can_merge_all:
                collect[ui][part.slice] = MergeAttrMapping(src)
            else:
                udf.merge(dest=MergeAttrMapping(dest), src=MergeAttrMapping(src))
        damage[r0:r1] = True

    @staticmethod
    def _gather_rows(dist, comm, bounds, equal, pending=None):
        """assemble a nav buffer whose rows [a_r, b_r) are valid on rank r: all-gather of the
        contiguous row blocks; ragged blocks are padded to the largest one.  bool / complex
        buffers travel as bytes / real pairs (NCCL has no bool, gloo no complex)."""
        rank = dist.get_rank()
        a, b = bounds[rank]
        orig = comm
        if comm.dtype == torch.bool:
            comm = comm.view(torch.uint8)
        elif comm.is_complex():
            comm = torch.view_as_real(comm)
        if not comm.is_contiguous():
            raise UDFException('nav buffers must be contiguous for the multi-rank merge')
        if equal:
            if pending is not None:
                work = dist.all_gather_into_tensor(comm.view(-1), comm[a:b].contiguous().view(-1),
                                                   async_op=True)
                pending.append((work, comm))
                return orig
            dist.all_gather_into_tensor(comm.view(-1), comm[a:b].contiguous().view(-1))
            return orig
        width = max(bb - aa for aa, bb in bounds)
        if width == 0:
            return orig
        block = torch.zeros((width,) + tuple(comm.shape[1:]), dtype=comm.dtype,
                            device=comm.device)
        block[:b - a] = comm[a:b]
        gathered = torch.empty((len(bounds),) + tuple(block.shape), dtype=comm.dtype,
                               device=comm.device)
        dist.all_gather_into_tensor(gathered.view(-1), block.view(-1))
        for r, (aa, bb) in enumerate(bounds):
            if r != rank and bb > aa:
                comm[aa:bb] = gathered[r, :bb - aa]
        return orig

    def _merge_ranks(self, dist, udfs, partitions, roi_flat, damage, device, pending=None):
        """assemble the dataset-sized buffers across ranks.

        * the fused slab and every nav buffer of a UDF with the *default* merge: all-gather of
          each rank's contiguous row block (the default merge is a row copy,
          udf/base.py:1420-1453);
        * UDFs that declare ``_additive_merge`` (SumUDF: ``dest += src``, udf/sum.py:51-53):
          all-reduce(sum);
        * any other custom ``merge``: every rank's locally merged buffers are gathered and
          ``udf.merge(dest, src)`` is replayed in rank (= partition) order onto freshly zeroed
          dataset buffers on every rank -- a rank acts as one large partition, which is what an
          associative merge (max / min, variance-style, 'single' buffers) needs."""
        world = dist.get_world_size()
        rank = dist.get_rank()
        backend = dist.get_backend()
        bounds = []
        for r in range(world):
            mine = self.my_partitions(partitions, r, world, self._rank_weights)
            if mine:
                a, _ = self._roi_range(mine[0], roi_flat)
                _, b = self._roi_range(mine[-1], roi_flat)
            else:
                a = b = 0
            bounds.append((a, b))
        sizes = [b - a for a, b in bounds]
        equal = len(set(sizes)) == 1 and sizes[0] > 0

        def on_wire(t):
            return t if backend == 'nccl' or not t.is_cuda else t.cpu()

        if self._slab is not None:
            t = self._slab
            comm = on_wire(t)
            self._gather_rows(dist, comm, bounds, equal,
                              pending=pending if comm is t else None)
            if comm is not t:
                t.copy_(comm)
        for ui, udf in enumerate(udfs):
            names = [name for name, buf in udf.results.items()
                     if buf.use != 'result_only' and buf.has_data()
                     and (ui, name) not in self._slab_map]
            if not names:
                continue
            default_merge = type(udf).merge is _BaseUDF.merge
            if default_merge or getattr(udf, '_additive_merge', False):
                for name in names:
                    buf = udf.results.get_buffer(name)
                    t = buf.tensor
                    comm = on_wire(t)
                    if default_merge and buf.kind == 'nav':
                        self._gather_rows(dist, comm, bounds, equal)
                    elif default_merge:
                        raise UDFException(
                            "buffer '%s' (kind=%s) needs a custom merge" % (name, buf.kind))
                    elif comm.is_complex():
                        dist.all_reduce(torch.view_as_real(comm), op=dist.ReduceOp.SUM)
                    else:
                        dist.all_reduce(comm, op=dist.ReduceOp.SUM)
                    if comm is not t:
                        t.copy_(comm)
                continue
            # custom merge: gather every rank's buffers, replay merge in rank order
            per_rank = {}
            for name in names:
                t = udf.results.get_buffer(name).tensor
                comm = on_wire(t).contiguous()
                wire = comm.view(torch.uint8) if comm.dtype == torch.bool else (
                    torch.view_as_real(comm) if comm.is_complex() else comm)
                gathered = torch.empty((world,) + tuple(wire.shape), dtype=wire.dtype,
                                       device=wire.device)
                dist.all_gather_into_tensor(gathered.view(-1), wire.reshape(-1))
                if comm.dtype == torch.bool:
                    gathered = gathered.view(torch.bool)
                elif comm.is_complex():
                    gathered = torch.view_as_complex(gathered)
                per_rank[name] = gathered.to(t.device)
            for name in names:
                udf.results.get_buffer(name).tensor.zero_()
            for r in range(world):
                a, b = bounds[r]
                if b <= a:
                    continue
                dest, src = {}, {}
                for name in names:
                    buf = udf.results.get_buffer(name)
                    dest[name] = buf.rows(a, b)
                    src[name] = (per_rank[name][r][a:b] if buf.kind == 'nav'
                                 else per_rank[name][r])
                udf.merge(dest=MergeAttrMapping(dest), src=MergeAttrMapping(src))
        damage[:] = True

    # -- results ---------------------------------------------------------------------------------
    def _make_results(self, udfs, ds_shape, roi, damage):
        out = []
        for udf in udfs:
            udf.meta._valid_nav_mask = damage
            decl = udf.get_result_buffers()
            res = udf.get_results()
            for k, v in decl.items():
                if k not in res and v.use is None:
                    res[k] = udf.results.get_buffer(k).raw_data
            wrapped = {}
            for name, arr in res.items():
                d = decl[name]
                arr_np = to_numpy(arr) if not isinstance(arr, np.ndarray) else arr
                if np.dtype(arr_np.dtype).kind != np.dtype(d.dtype).kind:
                    raise UDFException(
                        "the returned ndarray '%s' has a different dtype kind (%s) than "
                        "declared (%s)" % (name, arr_np.dtype, d.dtype))
                wrapped[name] = ResultBuffer.wrap(d, arr_np, ds_shape, roi)
            out.append(wrapped)
        return out


def run_udf(dataset, udf, roi=None, device=None, fuse=True, corrections=None):
    """``Context.run_udf`` for this runtime (reference api.py:914-1051): one UDF -> dict of
    result buffers, a list of UDFs -> list of dicts."""
    many = isinstance(udf, (list, tuple))
    udfs = list(udf) if many else [udf]
    res = UDFRunner(udfs, fuse=fuse).run_for_dataset(dataset, roi=roi, device=device,
                                                     corrections=corrections)
    return res.buffers if many else res.buffers[0]
