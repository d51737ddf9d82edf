#!/usr/bin/env python
"""bench.py -- headline benchmark of the ApplyMasksUDF / CoM hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload at N=1: BASELINE configs[1] -- 256x256 nav x 256x256 sig float32, ApplyMasksUDF with 8
dense masks + CoMUDF (11 fused mask columns), one B200.  For N>1 (torchrun, one rank per GPU)
every rank owns one such 256x256-nav shard of a (256*N)x256 nav dataset (weak scaling) and the
nav-shaped result buffers are assembled with one NCCL all-gather per step.

A *step* = one pass of the hot path over the whole (per-rank) dataset through the plugin API:
partition -> tile -> fused kernel -> merge (-> all-gather).  JSON keys (one line, rank 0):
  value        frames/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  roofline     the dominant kernel (k6_tensor_kernel, tcgen05 dense masked reduction) timed live
               with CUDA events on its stream: algorithmic bytes (frames x sig_size x 4,
               SURVEY 8d) / mean launch duration vs MEASURED_PEAKS hbm_gbs (a COPY benchmark) and
               vs `peak_read_only`, a read-only streaming probe measured in this run
  parity       after the timed loop: the first and last 256 frames of every rank's shard are
               recomputed with the CPU oracle on oracle.synth data; max relative error, the run
               FAILS (exit 1) above 1e-5
  configs      (N=1) device-timed passes of the other BASELINE configs: cfg3 (uint16, SumUDF +
               SumSigUDF + 4 ring masks), cfg4 (RadialFourierAnalysis, 32 bins), cfg5 shard
               (131072 frames, 16 masks + CoM): ms, frames/s, roofline.frac, kernel, parity
  multi        (N>1) cfg5 (1024x1024 nav over 8 GPUs: a 128x1024-nav shard per rank) and cfg2
               strong scaling (ONE 256x256-nav dataset split over the N ranks), both with the
               all-gather inside the step
  e2e          the same metric through run_udf() on a HOST (pinned) numpy dataset: H2D of every
               frame, the all-gather (N>1) and D2H / get_results of every result inside the
               timed region; `h2d_peak_gbs` = bare concurrent cudaMemcpyAsync from the same
               pinned buffers
  cpu_baseline the oracle port of the reference's CPU formulation on the host cores, on a bounded
               frame sample of the same workload, timed BEFORE the GPU legs with the same
               procedure as `--impl reference` (rank 0, N=1 only)
--impl reference times the reference's own CPU formulation (oracle port; the reference is pure
Python and cannot travel to the GPU box) on the host cores only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NAV = (256, 256)           # per-GPU shard of the navigation axis
SIG = (256, 256)
N_MASKS = 8
DATA_SEED = 1001
MASK_SEED = 2001
METRIC = 'frames/s on 256^2 nav x 256^2 sig float32 ApplyMasksUDF (8 dense masks) + CoM'
WORKLOAD = ('cfg2: 256x256 nav x 256x256 sig float32, ApplyMasksUDF 8 dense masks + CoMUDF '
            '(11 fused mask columns), per GPU')
PARITY_TOL = 1e-5
PARITY_FRAMES = 256
CPU_SAMPLE_FRAMES = 2048   # 512 MiB of the cfg2 frame stream per CPU step
KERNEL_NAMES = {
    1: 'k1_dense_tma_kernel (FFMA2, even/odd tile)', 2: 'generic kernel',
    3: 'k1_pair_kernel (FFMA2, mask-pair tile)', 4: 'k4_group_kernel (FFMA2 group-sparse)',
    6: 'k6_tensor_kernel (tcgen05 kind::tf32, split-TF32, TMEM accumulators)',
    7: 'k7_group_tensor_kernel (tcgen05, ring-major plan)',
    70: 'k7_group_tensor_kernel (tcgen05, quad / banded plan)',
    71: 'k7_group_tensor_kernel (tcgen05, mirror-symmetric plan)',
    10: 'k10_walk_kernel (tcgen05, dense-walk plan: TMA boxes in ring order)',
    8: 'k8_int_kernel (tcgen05 kind::i8, exact integer path)', 20: 'k2_csc_kernel',
}


def bench_masks(sy, sx, count, seed, uniform_fn):
    """cfg2/cfg5 mask mix: uniform random / disk / ring / gradient; built from the
    product mask generators (the oracle twin lives in tests/golden_inputs.py)."""
    from libertem_b200 import masks as M
    out = []
    cy, cx = sy // 2, sx // 2
    for i in range(count):
        kind = i % 4
        if kind == 0:
            m = uniform_fn(sy * sx, seed + i).reshape(sy, sx)
        elif kind == 1:
            m = M.circular(cx, cy, sx, sy, radius=min(sy, sx) / 4 + i).astype(np.float32)
        elif kind == 2:
            m = M.ring(cx, cy, sx, sy, radius=min(sy, sx) / 3 + i,
                       radius_inner=min(sy, sx) / 6).astype(np.float32)
        else:
            # first-moment (gradient) mask as the reference builds it: the pixel index ramps
            # masks.gradient_x / gradient_y (src/libertem/masks.py:415-422), like CoMUDF's own
            # y*D / x*D rows.  (Round 1 used the detector-centred form; a zero-mean mask makes
            # "error relative to the result" measure the cancellation, not the kernel -- the
            # `abs_scale` figures of the parity block cover that case.)
            m = (M.gradient_x(sx, sy) * 0.5 + M.gradient_y(sx, sy) * 0.25).astype(np.float32)
        out.append(m)
    return np.stack(out)


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json, copy)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def physical_cores():
    try:
        import psutil
        return psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        return os.cpu_count()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '10', '-i', str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()

        def collect(lo, hi):
            sm, smax, reasons, power = [], None, set(), []
            self.masks = getattr(self, 'masks', set())
            for ts, line in self.lines:
                parts = [p.strip() for p in line.split(',')]
                if len(parts) < 9:
                    continue
                try:
                    smax = float(parts[2])
                    if lo <= ts <= hi:
                        sm.append(float(parts[1]))
                        power.append(float(parts[3]))
                        self.masks.add(parts[4])
                        for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown',
                                              'sw_thermal_slowdown', 'sw_power_cap'),
                                             parts[5:9]):
                            if val.lower().startswith('active'):
                                reasons.add(name)
                except ValueError:
                    continue
            return sm, smax, reasons, power

        # nvidia-smi prints a sample some ms after taking it: accept a small lag
        sm, smax, reasons, power = collect(t0, t1 + 0.03)
        window = 'timed region'
        if not sm:
            sm, smax, reasons, power = collect(t0 - 0.25, t1 + 0.25)
            window = 'timed region +-0.25 s (region shorter than the sampling lag)'
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window,
                'sm_mhz_min': min(sm) if sm else None, 'sm_mhz_max_seen': max(sm) if sm else None,
                'active_bitmasks': sorted(self.masks),
                'power_w_max': max(power) if power else None}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU formulation on the host cores.  ONE procedure
# for both (`--impl reference` and the GPU arm's `cpu_baseline`): the same sample, the same
# warm-up and step counts, run before any GPU leg.
# ----------------------------------------------------------------------------------------------

def cpu_pass(flat, masks_t, com_t, use_torch=True):
    """what the reference does per partition tile on CPU: ApplyMasksUDF and CoMUDF are separate
    UDFs, each one GEMM over the tile -- torch.mm by default (udf/masks.py:59-66,
    udf/com.py:568-582) or numpy ``@`` with use_torch=False (udf/masks.py:76-77)"""
    from oracle import udf_oracle as O
    a = O.process_flat(flat, masks_t, use_torch=use_torch)
    b = O.process_flat(flat, com_t, use_torch=use_torch)
    return a, b


def oracle_masks(sig, n_masks, seed):
    from oracle import synth, udf_oracle as O
    stack = bench_masks(sig[0], sig[1], n_masks, seed, lambda n, s: synth.uniform_f32(0, n, s))
    full = (slice(0, sig[0]), slice(0, sig[1]))
    masks_t = O.masks_for_sig_slice(stack, full, np.float32)
    com_t = O.masks_for_sig_slice(O.com_mask_stack(sig, sig[0] // 2, sig[1] // 2), full,
                                  np.float32)
    return masks_t, com_t


def oracle_frames(f0, f1, k, seed, dtype=np.float32):
    """frames [f0, f1) of the synthetic stream as the oracle generates them (host)"""
    from oracle import synth
    if np.dtype(dtype) == np.uint16:
        return synth.poisson3_u16(f0 * k, (f1 - f0) * k, seed).reshape(f1 - f0, k)
    out = np.empty((f1 - f0, k), dtype=np.float32)
    for a in range(f0, f1, 256):
        b = min(f1, a + 256)
        out[a - f0:b - f0] = synth.uniform_f32(a * k, (b - a) * k, seed).reshape(b - a, k)
    return out


def cpu_leg(steps, warmup):
    """the bounded CPU sample: CPU_SAMPLE_FRAMES frames of the cfg2 stream, both GEMM formulations
    the reference has (torch.mm default, numpy @); returns the FASTER one -- the most favourable
    reading of the reference's CPU path"""
    import torch
    cores = physical_cores()
    torch.set_num_threads(cores)
    k = SIG[0] * SIG[1]
    flat = oracle_frames(0, CPU_SAMPLE_FRAMES, k, DATA_SEED)
    masks_t, com_t = oracle_masks(SIG, N_MASKS, MASK_SEED)
    steps = max(1, min(int(steps), 20))
    warmup = max(1, min(int(warmup), 3))
    best = None
    for use_torch in (True, False):
        for _ in range(warmup):
            cpu_pass(flat, masks_t, com_t, use_torch)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_pass(flat, masks_t, com_t, use_torch)
        dt = time.perf_counter() - t0
        fps = flat.shape[0] * steps / dt
        if best is None or fps > best[0]:
            best = (fps, dt / steps, 'torch.mm' if use_torch else 'numpy @')
    sample = (f'{CPU_SAMPLE_FRAMES} frames ({flat.nbytes / 2**20:.0f} MiB) of the cfg2 stream per '
              f'step, {steps} steps after {warmup} warm-up, oracle port of the reference '
              'formulation: one GEMM per UDF (8 masks, then 3 CoM masks), faster of torch.mm / '
              f'numpy @ = {best[2]}, data resident in host RAM, run before any GPU leg')
    return {'value': best[0], 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'sample': sample, 'ms_per_step': best[1] * 1e3, 'steps': steps, 'warmup': warmup}


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cpu = cpu_leg(args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': cpu['value'], 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': cpu['steps'], 'warmup': cpu['warmup'],
        'ms_per_step': cpu['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'nav': list(NAV), 'sig': list(SIG),
                   'n_masks': N_MASKS, 'com': True, 'sample_frames': CPU_SAMPLE_FRAMES,
                   'note': 'frames/s on a bounded sample of the same frame stream (the GPU arm '
                           'runs all 65536 frames per step); steps capped at 20'},
        'cpu_baseline': {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': cpu['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def _gpu_numa_node(torch, local_rank):
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f'/sys/bus/pci/devices/{bus}/numa_node') as f:
            return int(f.read().strip())
    except Exception:
        return None


def _bind_near_gpu(torch, local_rank):
    """Multi-GPU runs: pin this rank to the CPUs NVML names for its GPU, so that the pinned host
    buffers of the e2e leg (first touch) and the copy threads sit on the GPU's NUMA node
    instead of wherever torchrun started the process.  Returns the number of CPUs or None."""
    if os.environ.get('LTB200_NO_AFFINITY') == '1' or not hasattr(os, 'sched_setaffinity'):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(local_rank)
            bus = '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {i for i in range(n_cpu) if (int(words[i // 64]) >> (i % 64)) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


class Ctx:
    """per-process state of the GPU arm"""

    def __init__(self, args):
        import logging
        logging.getLogger('libertem_b200').setLevel(logging.ERROR)
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise RuntimeError('bench.py needs a GPU (there is no CPU fallback; use --impl '
                               'reference for the CPU arm)')
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device('cuda', self.local_rank)
        self.cpu_affinity = None
        self.numa_node = _gpu_numa_node(torch, self.local_rank)
        if self.world > 1:
            self.cpu_affinity = _bind_near_gpu(torch, self.local_rank)
            dist.init_process_group('nccl', device_id=self.device)
            if self.rank == 0:
                try:       # for the record (stderr): which GPUs share a PCIe switch / NUMA node
                    topo = subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True,
                                          text=True, timeout=20).stdout
                    print(topo, file=sys.stderr, flush=True)
                except Exception:
                    pass

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device=self.device, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def dev_uniform(self, n, seed):
        from libertem_b200 import engine
        return engine.synth_fill((n,), np.float32, seed, self.device).cpu().numpy()


def timed_steps(ctx, step, steps, finish=None):
    """EXACTLY `steps` calls of step() between barrier + synchronize on both sides, CUDA events
    on the launching stream, max over ranks.  Returns (ms per step, wall t0, wall t1)."""
    torch = ctx.torch
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    t0 = time.time()
    e0.record()
    for _ in range(steps):
        step()
    if finish is not None:
        finish()
    e1.record()
    ctx.barrier()
    t1 = time.time()
    return ctx.max_over_ranks(e0.elapsed_time(e1)) / steps, t0, t1


def rel_err_cols(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).reshape(-1, ref.shape[-1]).max(axis=0) + 1e-30
    return float((np.abs(got - ref).reshape(-1, ref.shape[-1]) / scale).max())


def parity_masks_com(ctx, udfs, shard_rows, first_frame, sig, n_masks, mask_seed, data_seed):
    """first and last PARITY_FRAMES frames of this rank's shard: ApplyMasksUDF 'intensity' and
    CoMUDF 'raw_mask_result' of the timed run vs the CPU oracle (oracle.udf_oracle.process_flat
    on oracle.synth data).  shard_rows = (row0, row1) of the shard in the result buffers,
    first_frame = dataset frame index of row0.  Returns (max error relative to the column's
    largest result -- the pass/fail figure --, this path's and the float32 oracle's max error
    against the float64 sums in units of sum|x||m|)."""
    k = sig[0] * sig[1]
    masks_t, com_t = oracle_masks(sig, n_masks, mask_seed)
    inten = udfs[0].results.get_buffer('intensity').tensor
    raw = udfs[1].results.get_buffer('raw_mask_result').tensor
    r0, r1 = shard_rows
    n = min(PARITY_FRAMES, r1 - r0)
    worst = ours64 = ref64 = 0.0
    for a in sorted({r0, r1 - n}):
        f0 = first_frame + (a - r0)
        flat = oracle_frames(f0, f0 + n, k, data_seed)
        ref_m, ref_c = cpu_pass(flat, masks_t, com_t)
        got_m, got_c = inten[a:a + n].cpu().numpy(), raw[a:a + n].cpu().numpy()
        worst = max(worst, rel_err_cols(got_m, ref_m), rel_err_cols(got_c, ref_c))
        f64 = flat.astype(np.float64)
        for got, ref, mt in ((got_m, ref_m, masks_t), (got_c, ref_c, com_t)):
            exact = f64 @ mt.astype(np.float64)
            scale = (np.abs(f64) @ np.abs(mt).astype(np.float64)).max(axis=0) + 1e-30
            ours64 = max(ours64, float((np.abs(got - exact) / scale).max()))
            ref64 = max(ref64, float((np.abs(ref - exact) / scale).max()))
    return (ctx.max_over_ranks(worst), ctx.max_over_ranks(ours64), ctx.max_over_ranks(ref64))


def read_only_probe(ctx, buf):
    """GB/s of the read-only streaming probes over `buf` (device-resident, >> L2)"""
    from libertem_b200 import engine
    torch = ctx.torch
    nbytes = buf.numel() * buf.element_size()
    out = {}
    for mode, name in ((0, 'bulk_tma'), (1, 'ldg128')):
        for _ in range(2):
            engine.probe_read(buf, mode)
        torch.cuda.synchronize()
        best = None
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.probe_read(buf, mode)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        out[name] = nbytes / best / 1e6
    return out


def dense_leg(ctx, nav_rank, n_masks, mask_seed, data_seed, steps, strong=False,
              async_merge=True):
    """ApplyMasksUDF(n_masks dense) + CoMUDF on a float32 256x256-sig dataset: every rank owns a
    contiguous `nav_rank` shard of a (nav_rank[0] * world, nav_rank[1]) scan (weak; strong: the
    caller passes the per-rank slice of a fixed scan).  Returns a dict incl. parity."""
    torch = ctx.torch
    from libertem_b200 import engine
    from libertem_b200.io import SyntheticDataSet
    from libertem_b200.runner import UDFRunner
    from libertem_b200.udf import ApplyMasksUDF, CoMUDF
    world, rank, device = ctx.world, ctx.rank, ctx.device
    nav = (nav_rank[0] * world, nav_rank[1])
    frames_per_rank = nav_rank[0] * nav_rank[1]
    total = frames_per_rank * world
    k = SIG[0] * SIG[1]
    stack = bench_masks(SIG[0], SIG[1], n_masks, mask_seed, ctx.dev_uniform)
    ds = SyntheticDataSet(nav + SIG, np.float32, seed=data_seed, num_partitions=world)
    parts = list(ds.get_partitions())
    mine = UDFRunner.my_partitions(parts, rank, world)
    ds.materialize(device, mine)
    torch.cuda.synchronize()
    udfs = [ApplyMasksUDF(mask_factories=lambda: stack, mask_count=n_masks,
                          mask_dtype=np.float32, use_sparse=False), CoMUDF()]
    runner = UDFRunner(udfs)
    pending = []

    def step():
        res = runner.run_for_dataset(ds, device=device, finalize=False,
                                     async_merge=async_merge and world > 1)
        pending.append(res)
        if len(pending) > 2:          # bound the number of result slabs in flight
            pending.pop(0).wait()

    def finish():
        while pending:
            pending.pop(0).wait()

    for _ in range(3):
        step()
    finish()
    engine.launch_count(reset=True)
    engine.EVENT_LOG = []
    ms, t0, t1 = timed_steps(ctx, step, steps, finish)
    launches = engine.launch_count()
    ev_log, engine.EVENT_LOG = engine.EVENT_LOG, None
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b, *_ in ev_log]))
    bytes_per_launch = float(np.mean([f * kk * isz for _, _, f, kk, isz in ev_log]))
    peak, _src = measured_peak()
    achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9
    err, err_abs, ref_abs = parity_masks_com(ctx, udfs, (mine[0].start, mine[-1].stop),
                                             mine[0].start, SIG, n_masks, mask_seed, data_seed)
    out = {'nav': list(nav), 'sig': list(SIG), 'n_masks': n_masks, 'fused_columns': n_masks + 3,
           'frames': total, 'frames_per_gpu': frames_per_rank, 'steps': steps,
           'ms_per_step': ms, 'frames_per_s': total / (ms * 1e-3),
           'kernel': KERNEL_NAMES.get(engine.last_kernel(), str(engine.last_kernel())),
           'kernel_ms': kern_ms, 'roofline_frac': achieved / peak, 'achieved_gbs': achieved,
           'kernel_share_of_step': kern_ms * len(ev_log) / steps / ms,
           'gpu_launches_per_step': launches / steps,
           'merge': ('nccl all_gather of the nav slab per step, issued asynchronously (overlaps '
                     'the next step\'s kernel), all completed inside the timed region')
           if world > 1 else 'none (single rank)',
           'parity_max_rel_err': err, 'parity_ok': bool(err <= PARITY_TOL),
           'parity_err_vs_f64_abs_scale': err_abs, 'oracle_err_vs_f64_abs_scale': ref_abs}
    return out, ds, mine, stack, (t0, t1), launches, udfs


def config_legs(ctx):
    """cfg3, cfg4 and the cfg5 shard on one GPU: one pass each through UDFRunner, device-timed,
    with a first/last-frames oracle check"""
    torch = ctx.torch
    from libertem_b200 import engine, masks as M
    from libertem_b200.io import SyntheticDataSet
    from libertem_b200.runner import UDFRunner
    from libertem_b200.udf import ApplyMasksUDF, SumUDF, SumSigUDF
    from libertem_b200.api import Context
    from oracle import udf_oracle as O
    device = ctx.device
    peak, _ = measured_peak()
    out = {}

    def timed(runner, ds, steps):
        ds.materialize(device)
        for _ in range(2):
            runner.run_for_dataset(ds, device=device, finalize=False)
        torch.cuda.synchronize()
        engine.launch_count(reset=True)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            runner.run_for_dataset(ds, device=device, finalize=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        frames = ds.shape.nav.size
        nbytes = ds.shape.size * ds.dtype.itemsize
        return {'frames': frames, 'ms_per_pass': ms, 'frames_per_s': frames / ms * 1e3,
                'achieved_gbs': nbytes / ms / 1e6, 'roofline_frac': nbytes / ms / 1e6 / peak,
                'algorithmic_bytes': nbytes, 'launches_per_pass': engine.launch_count() / steps,
                'kernel': KERNEL_NAMES.get(engine.last_kernel(), str(engine.last_kernel())),
                'unfused_calls': runner.stats['unfused_calls']}

    # ---- cfg3: 512x512 nav x 128x128 sig uint16, SumUDF + SumSigUDF + 4 sparse ring masks
    try:
        shape = (512, 512, 128, 128)
        ds = SyntheticDataSet(shape, np.uint16, seed=103, num_partitions=1)
        rings = [(8, 16), (20, 28), (32, 40), (44, 52)]
        facs = [lambda ri=ri, ro=ro: M.ring(64, 64, 128, 128, ro, ri) for ri, ro in rings]
        udfs = [SumUDF(), SumSigUDF(), ApplyMasksUDF(mask_factories=facs, use_sparse=True,
                                                     mask_dtype=np.float32)]
        runner = UDFRunner(udfs)
        line = timed(runner, ds, 10)
        # parity (bit-exact): first/last frames' SumSig + ring sums, and the frame sum of a slice
        from oracle import masks_gen
        k = 128 * 128
        ring_stack = np.stack([masks_gen.ring(64, 64, 128, 128, ro, ri).astype(np.float32)
                               for ri, ro in rings]).reshape(4, k)
        n = shape[0] * shape[1]
        sumsig = udfs[1].results.get_buffer('intensity').tensor
        inten = udfs[2].results.get_buffer('intensity').tensor
        exact = True
        for f0 in (0, n - PARITY_FRAMES):
            flat = oracle_frames(f0, f0 + PARITY_FRAMES, k, 103, np.uint16).astype(np.float32)
            exact &= bool(np.array_equal(sumsig[f0:f0 + PARITY_FRAMES].cpu().numpy(),
                                         flat.sum(axis=1, dtype=np.float64).astype(np.float32)))
            exact &= bool(np.array_equal(inten[f0:f0 + PARITY_FRAMES].cpu().numpy(),
                                         (flat.astype(np.float64) @ ring_stack.T.astype(np.float64)
                                          ).astype(np.float32)))
        line.update(parity='bit-exact vs oracle on the first / last %d frames (SumSigUDF, ring '
                           'masks)' % PARITY_FRAMES if exact else 'MISMATCH', parity_ok=exact,
                    workload='cfg3: 512x512 nav x 128x128 sig uint16, SumUDF + SumSigUDF + 4 '
                             'sparse ring masks, one fused pass')
        out['cfg3'] = line
        del ds, runner, udfs
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out['cfg3'] = {'error': repr(e)}

    # ---- cfg5 shard: 128x1024 nav x 256x256 sig f32, 16 masks + CoM (what one of 8 GPUs holds)
    try:
        saved = ctx.world, ctx.rank
        ctx.world, ctx.rank = 1, 0
        try:
            line, ds5, _, _, _, _, _ = dense_leg(ctx, (128, 1024), 16, 2005, 105, 5)
        finally:
            ctx.world, ctx.rank = saved
        line['workload'] = ('cfg5 shard: 128x1024 nav x 256x256 sig float32, 16 dense masks + '
                            'CoM (19 fused columns) = one of the 8 ranks of cfg5')
        out['cfg5_shard'] = line
        del ds5
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out['cfg5_shard'] = {'error': repr(e)}

    # ---- cfg4: 256x256 nav x 512x512 sig f32 (64 GiB resident), RadialFourierAnalysis 32 bins
    try:
        free, _total = torch.cuda.mem_get_info(device)
        nav0 = 256 if free > (80 << 30) else 32
        shape = (nav0, 256, 512, 512)
        ds = SyntheticDataSet(shape, np.float32, seed=104, num_partitions=1)
        a = Context().create_radial_fourier_analysis(ds, n_bins=32)
        udf = a.get_udf()
        runner = UDFRunner([udf])
        line = timed(runner, ds, 2)
        # parity: 16 first + 16 last frames vs the oracle's own mask stack (float64 direct sums)
        p = O.radial_fourier_params((512, 512), n_bins=32)
        stack = O.radial_mask_stack((512, 512), **p).reshape(800, -1)
        k = 512 * 512
        res = udf.results.get_buffer('intensity').tensor
        n = shape[0] * shape[1]
        worst = 0.0
        for f0 in (0, n - 16):
            flat = oracle_frames(f0, f0 + 16, k, 104).astype(np.float64)
            ref = np.zeros((16, 800), dtype=np.complex128)
            for b in range(32):
                rows = stack[b * 25:(b + 1) * 25]
                sup = np.nonzero(np.any(rows != 0, axis=0))[0]
                ref[:, b * 25:(b + 1) * 25] = flat[:, sup] @ rows[:, sup].T.astype(np.complex128)
            got = res[f0:f0 + 16].cpu().numpy().astype(np.complex128)
            scale = np.abs(ref.reshape(16, 32, 25)[:, :, :1]).max(axis=0, keepdims=True)
            worst = max(worst, float((np.abs(got - ref).reshape(16, 32, 25) / scale).max()))
        line.update(parity_max_rel_err=worst, parity_ok=bool(worst <= PARITY_TOL),
                    workload='cfg4: %dx256 nav x 512x512 sig float32%s, RadialFourierAnalysis 32 '
                             'bins x 25 orders (800 complex masks), one partition' %
                             (nav0, '' if nav0 == 256 else ' (nav sub-sample: not enough HBM)'),
                    bound='tensor pipe / A-operand hand-over (group-sparse GEMM on dense TMA '
                          'boxes), reported against the HBM roofline of the frame bytes')
        out['cfg4'] = line
        del ds, runner, udf, a
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out['cfg4'] = {'error': repr(e)}
    return out


def gpu_arm(args):
    ctx = Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    world, rank, device = ctx.world, ctx.rank, ctx.device
    from libertem_b200 import engine
    k = SIG[0] * SIG[1]
    steps = args.steps
    line_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        line_cpu = cpu_leg(args.steps, args.warmup)        # before any GPU leg

    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.lines and time.time() - t_wait < 5.0:   # wait for the first sample
        time.sleep(0.01)
    head, ds, mine, stack, (t0, t1), launches, _udfs = dense_leg(
        ctx, NAV, N_MASKS, MASK_SEED, DATA_SEED, steps)
    clocks = sampler.stop(t0, t1)
    peak, peak_src = measured_peak()
    probe = read_only_probe(ctx, ds.partition_tensor(mine[0], device))
    peak_ro = max(probe.values())
    roofline = {'bound': 'hbm', 'achieved': head['achieved_gbs'], 'peak': peak, 'unit': 'GB/s',
                'frac': head['roofline_frac'], 'traffic': None,
                'traffic_source': 'not measured in this run; ncu capture of this command: '
                                  'profiles/k6_traffic.json (17.196 GB per launch vs 17.180 GB '
                                  'algorithmic)',
                'peak_source': peak_src,
                'peak_read_only': peak_ro, 'frac_read_only': head['achieved_gbs'] / peak_ro,
                'peak_read_only_source': 'ltb200_probe_read over the same 17.18 GB shard in this '
                                         'run (best of 5): %s' % json.dumps(
                                             {n: round(v, 1) for n, v in probe.items()}),
                'kernel': head['kernel'], 'kernel_ms': head['kernel_ms'],
                'algorithmic_bytes_per_launch': float(NAV[0] * NAV[1] * k * 4),
                'kernel_share_of_step': head['kernel_share_of_step']}
    nav = (NAV[0] * world, NAV[1])
    total_frames = NAV[0] * NAV[1] * world
    line = {
        'metric': METRIC, 'value': head['frames_per_s'], 'unit': 'frames/s', 'n_gpus': world,
        'steps': steps, 'warmup': 3, 'ms_per_step': head['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic (counter-based hash, uniform [0,1), seed %d)' % DATA_SEED,
        'config': {'workload': WORKLOAD, 'nav': list(nav), 'sig': list(SIG), 'n_masks': N_MASKS,
                   'com': True, 'fused_columns': N_MASKS + 3, 'partitions_per_gpu': 1,
                   'sig_bytes_per_frame': k * 4, 'frames_per_gpu': NAV[0] * NAV[1],
                   'l2': 'inputs %.1f GB per GPU >> 126 MB L2, streamed once per step' %
                         (NAV[0] * NAV[1] * k * 4 / 1e9),
                   'merge': head['merge'], 'cpu_affinity': ctx.cpu_affinity,
                   'gpu_numa_node': ctx.numa_node},
        'hbm_gbs': total_frames * k * 4 / (head['ms_per_step'] * 1e-3) / 1e9 / world,
        'roofline': roofline, 'clocks': clocks, 'gpu_launches': int(launches),
        'parity': {'max_rel_err': head['parity_max_rel_err'], 'tol': PARITY_TOL,
                   'ok': head['parity_ok'],
                   'err_vs_f64_abs_scale': head['parity_err_vs_f64_abs_scale'],
                   'oracle_err_vs_f64_abs_scale': head['oracle_err_vs_f64_abs_scale'],
                   'abs_scale': 'max |result - float64 sum| / max_f sum_k |x||m| per column: this '
                                'path and the float32 oracle (torch.mm) against the exact sums',
                   'what': 'ApplyMasksUDF intensity + CoMUDF raw_mask_result of the timed run, '
                           'first and last %d frames of every rank\'s shard vs '
                           'oracle.udf_oracle.process_flat on oracle.synth data' % PARITY_FRAMES},
    }
    ok = head['parity_ok']

    if not args.no_e2e:
        del ds
        torch.cuda.empty_cache()
        line['e2e'] = e2e_leg(ctx, stack, total_frames)
        ds = None
    del ds
    torch.cuda.empty_cache()
    if world == 1 and not args.no_configs:
        line['configs'] = config_legs(ctx)
        ok = ok and all(c.get('parity_ok', False) for c in line['configs'].values()
                        if 'error' not in c)
    if world > 1 and not args.no_configs:
        multi = {}
        # cfg5: every rank holds a 128x1024-nav shard (at N=8 this IS BASELINE configs[4])
        m5, ds5, *_ = dense_leg(ctx, (128, 1024), 16, 2005, 105, max(3, min(steps, 10)))
        m5['workload'] = ('cfg5: %dx1024 nav x 256x256 sig float32, 16 dense masks + CoM, '
                          '128x1024-nav shard per GPU%s' %
                          (128 * world, ' (= BASELINE configs[4])' if world == 8 else
                           ' (cfg5-class weak shard; the full 1024x1024 nav needs 8 GPUs)'))
        multi['cfg5'] = m5
        del ds5
        torch.cuda.empty_cache()
        # cfg2 strong scaling: ONE 256x256-nav dataset split over the ranks
        ms2, ds2, *_ = dense_leg(ctx, (NAV[0] // world, NAV[1]), N_MASKS, MASK_SEED, DATA_SEED,
                                 max(3, min(steps, 20)))
        ms2['workload'] = 'cfg2 strong scaling: one 256x256 nav x 256x256 sig dataset over %d GPUs' % world
        ms2['scaling'] = 'strong'
        multi['cfg2_strong'] = ms2
        del ds2
        torch.cuda.empty_cache()
        line['multi'] = multi
        ok = ok and m5['parity_ok'] and ms2['parity_ok']
    if line_cpu is not None:
        line['cpu_baseline'] = {kk: line_cpu[kk] for kk in ('value', 'unit', 'cores', 'kind',
                                                            'sample')}
    line['parity']['all_ok'] = bool(ok)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        sys.exit(1)


def _host_frames(ctx, f0, f1):
    """frames [f0, f1) of the synthetic scan as a pinned host array (generated on the device in
    chunks by the counter-based generator, then copied down)"""
    torch = ctx.torch
    from libertem_b200 import engine
    k = SIG[0] * SIG[1]
    host = torch.empty((f1 - f0,) + SIG, dtype=torch.float32, pin_memory=True)
    step = 4096
    for a in range(f0, f1, step):
        b = min(f1, a + step)
        t = engine.synth_fill((b - a,) + SIG, np.float32, DATA_SEED, ctx.device, start=a * k)
        host[a - f0:b - f0].copy_(t)
    torch.cuda.synchronize()
    return host


def h2d_peak(ctx, probe_frames=4096, passes=8, reps=2):
    """bare H2D: cudaMemcpyAsync from a pinned 1 GiB host buffer (`passes` times per measurement)
    into one device buffer.  Returns (GB/s per GPU with all ranks copying at the same time, list
    over ranks; GB/s of every rank copying ALONE, list over ranks): placement problems (a rank
    whose pinned memory sits on a far socket) show up in the second list, shared-resource limits
    (host memory / IO-die bandwidth, PCIe switch uplinks) as the difference between the two."""
    torch = ctx.torch
    k = SIG[0] * SIG[1]
    flat = torch.empty((probe_frames, k), dtype=torch.float32, pin_memory=True)
    flat.zero_()
    dst = torch.empty_like(flat, device=ctx.device)
    nbytes = flat.numel() * flat.element_size() * passes

    def once():
        t0 = time.perf_counter()
        for _ in range(passes):
            dst.copy_(flat, non_blocking=True)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def gather(x):
        if ctx.world == 1:
            return [float(x)]
        t = torch.zeros(ctx.world, device=ctx.device, dtype=torch.float64)
        t[ctx.rank] = x
        ctx.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    once()
    best = None
    for _ in range(reps):
        ctx.barrier()
        dt = once()
        best = dt if best is None else min(best, dt)
    together = gather(nbytes / best / 1e9)
    alone = together
    if ctx.world > 1:
        mine = 0.0
        for r in range(ctx.world):
            ctx.barrier()
            if r == ctx.rank:
                mine = nbytes / once() / 1e9
        ctx.barrier()
        alone = gather(mine)
    return together, alone


def e2e_leg(ctx, stack, total_frames):
    """end to end through the public API on HOST pinned input: H2D of every frame, the kernel,
    the NCCL all-gather (N>1) and D2H / get_results of every result buffer inside the timed
    region, every step.  N>1: the scan is cut into 8 x N partitions and every rank takes a
    contiguous share proportional to its measured concurrent host-link rate
    (UDFRunner(rank_weights=...)), so all ranks finish together even when the host links of the
    box are not equal."""
    torch = ctx.torch
    from libertem_b200.io import MemoryDataSet
    from libertem_b200.io.memory import partition_boundaries
    from libertem_b200.runner import UDFRunner
    from libertem_b200.udf import ApplyMasksUDF, CoMUDF
    world, rank = ctx.world, ctx.rank
    together, alone = h2d_peak(ctx)
    weights = None
    if world == 1:
        host = _host_frames(ctx, 0, total_frames)
        hds = MemoryDataSet(data=host.reshape(NAV + SIG), num_partitions=1, sig_dims=2, pin=False)
        my_frames = total_frames
    else:
        n_parts = 8 * world
        weights = [round(w, 1) for w in together]
        bounds = partition_boundaries(total_frames, n_parts)
        mine = UDFRunner.my_partitions(bounds, rank, world, weights)
        f0, f1 = mine[0][0], mine[-1][1]
        host = _host_frames(ctx, f0, f1)
        hds = ShardedHostDataSet(host, (NAV[0] * world, NAV[1]) + SIG, f0, n_parts)
        my_frames = f1 - f0
    steps = min(ctx.args.steps, 3 if world == 1 else 2)

    def one():
        r = UDFRunner([ApplyMasksUDF(mask_factories=lambda: stack, mask_count=N_MASKS,
                                     mask_dtype=np.float32, use_sparse=False), CoMUDF()],
                      rank_weights=weights)
        res = r.run_for_dataset(hds, device=ctx.device).buffers
        return res[0]['intensity'].raw_data, res[1]['field'].raw_data

    one()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        a, b = one()
    torch.cuda.synchronize()
    if world > 1:
        ctx.dist.barrier()
    dt = ctx.max_over_ranks((time.perf_counter() - t0) / steps)
    assert a.shape[0] == total_frames
    k = SIG[0] * SIG[1]
    h2d = int(total_frames * k * 4)
    # every rank moves its own share; the aggregate is what the box's host links deliver
    agg = h2d / dt / 1e9
    peak_agg = float(sum(together))
    return {'value': total_frames / dt, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d,
            'd2h_bytes_per_step': int(total_frames * (N_MASKS + 3) * 4) * world, 'steps': steps,
            'ms_per_step': dt * 1e3, 'h2d_gbs_aggregate': agg,
            'h2d_peak_gbs': peak_agg, 'h2d_frac_of_peak': agg / peak_agg,
            'h2d_peak_gbs_per_rank_concurrent': [round(v, 2) for v in together],
            'h2d_peak_gbs_per_rank_alone': [round(v, 2) for v in alone],
            'frames_this_rank0': my_frames, 'rank_weights': weights,
            'note': 'run_for_dataset on pinned host data: double-buffered H2D tiles overlapped '
                    'with the kernel' + (', NCCL all-gather of the result slab, partition '
                                         'shares proportional to the measured per-rank H2D rate'
                                         if world > 1 else '') +
                    ', D2H of all result buffers and CoM get_results; h2d_peak_gbs = sum over '
                    'ranks of a bare cudaMemcpyAsync loop from pinned memory with all ranks '
                    'copying concurrently (per rank: ..._concurrent; one rank at a time: '
                    '..._alone)'}


class ShardedHostDataSet:
    """MemoryDataSet-compatible view of a multi-rank scan whose partitions live in per-rank
    pinned host memory: `host` holds frames [first, first + len(host)) of the scan, which is
    cut into `num_partitions` partitions; a rank only ever touches its own partitions"""

    def __new__(cls, host, shape, first, num_partitions):
        from libertem_b200.io.memory import MemoryDataSet
        from libertem_b200.common.shape import Shape

        class _Offset:
            """frames [first, first + n) of the scan"""

            def __init__(self, t):
                self.t = t
                self.is_cuda = False
                self.dtype = t.dtype

            def __getitem__(self, sl):
                assert sl.start >= first and sl.stop <= first + self.t.shape[0]
                return self.t[sl.start - first:sl.stop - first]

        class _DS(MemoryDataSet):
            def __init__(self):
                self._is_torch = True
                self.data = host
                self._shape = Shape(tuple(shape), sig_dims=2)
                self._dtype = np.dtype('float32')
                self.tileshape = None
                self.num_partitions = num_partitions
                self.tile_depth = None
                self._pin = False
                self._registered = False
                self._stage = {}

            def _flat(self):
                return _Offset(self.data.reshape((host.shape[0],) + tuple(shape[2:])))

        return _DS()


def set_workload(name):
    """cfg2 (default, the headline) or cfg5: 1024x1024 nav x 256x256 sig, 16 masks + CoM,
    nav-sharded over 8 GPUs (each rank holds a 128x1024 nav shard = 32 GiB)"""
    global NAV, N_MASKS, METRIC, WORKLOAD
    if name == 'cfg5':
        NAV = (128, 1024)
        N_MASKS = 16
        METRIC = 'frames/s on 1024^2 nav x 256^2 sig float32 ApplyMasksUDF (16 dense masks) + CoM'
        WORKLOAD = ('cfg5: 1024x1024 nav x 256x256 sig float32 over 8 GPUs, ApplyMasksUDF 16 dense '
                    'masks + CoMUDF (19 fused mask columns), 128x1024 nav shard per GPU')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-configs', action='store_true')
    ap.add_argument('--workload', default='cfg2', choices=['cfg2', 'cfg5'])
    args = ap.parse_args()
    set_workload(args.workload)
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
