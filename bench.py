#!/usr/bin/env python
"""bench.py -- headline benchmark of the ApplyMasksUDF / CoM hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload at N=1: BASELINE configs[1] -- 256x256 nav x 256x256 sig float32, ApplyMasksUDF with 8
dense masks + CoMUDF (11 fused mask columns), one B200.  For N>1 (torchrun, one rank per GPU)
every rank owns one such 256x256-nav shard of a (256*N)x256 nav dataset (weak scaling) and the
nav-shaped result buffers are assembled with one NCCL all-gather inside the timed step.

A *step* = one pass of the hot path over the whole (per-rank) dataset through the plugin API:
partition -> tile -> fused kernel -> merge (-> all-gather).  JSON keys:
  value      frames/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  roofline   the dominant kernel (k6_tensor_kernel, the tcgen05 dense masked reduction) timed
             live with CUDA events: algorithmic bytes
             (frames x sig_size x 4) / mean launch duration vs MEASURED_PEAKS hbm_gbs
  e2e        same metric through run_udf() on a HOST (pinned) numpy dataset: H2D of every frame
             and D2H/get_results of every result inside the timed region
  cpu_baseline  the oracle port (the reference's torch.mm formulation) on the host cores, on a
             bounded frame sample of the same workload (rank 0, N=1 only)
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
            m = ((M.gradient_x(sx, sy) - cx) * 0.5 + (M.gradient_y(sx, sy) - cy) * 0.25)
            m = m.astype(np.float32)
        out.append(m)
    return np.stack(out)


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
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
# reference arm / cpu baseline: the reference's CPU formulation on the host cores
# ----------------------------------------------------------------------------------------------

def cpu_pass(flat, masks_t, com_t, use_torch=True):
    """what the reference does per partition tile on CPU: ApplyMasksUDF and CoMUDF are separate
    UDFs, each one GEMM over the tile -- torch.mm by default (udf/masks.py:59-66,
    udf/com.py:568-582) or numpy ``@`` with use_torch=False (udf/masks.py:76-77)"""
    from oracle import udf_oracle as O
    a = O.process_flat(flat, masks_t, use_torch=use_torch)
    b = O.process_flat(flat, com_t, use_torch=use_torch)
    return a, b


def cpu_sample(n_frames):
    from oracle import synth, udf_oracle as O
    k = SIG[0] * SIG[1]
    flat = np.empty((n_frames, k), dtype=np.float32)
    step = 256
    for f0 in range(0, n_frames, step):
        f1 = min(n_frames, f0 + step)
        flat[f0:f1] = synth.uniform_f32(f0 * k, (f1 - f0) * k, DATA_SEED).reshape(f1 - f0, k)
    stack = bench_masks(SIG[0], SIG[1], N_MASKS, MASK_SEED,
                        lambda n, s: synth.uniform_f32(0, n, s))
    masks_t = O.masks_for_sig_slice(stack, (slice(0, SIG[0]), slice(0, SIG[1])), np.float32)
    com_t = O.masks_for_sig_slice(O.com_mask_stack(SIG, SIG[0] // 2, SIG[1] // 2),
                                  (slice(0, SIG[0]), slice(0, SIG[1])), np.float32)
    return flat, masks_t, com_t


def run_cpu(flat, masks_t, com_t, steps, warmup):
    """times both GEMM formulations the reference has (torch.mm default, numpy @) and returns
    the FASTER one -- the most favourable reading of the reference's CPU path"""
    import torch
    cores = physical_cores()
    torch.set_num_threads(cores)
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
    return best[0], best[1], cores, best[2]


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n_frames = 2048        # 512 MiB sample of the cfg2 frame stream
    flat, masks_t, com_t = cpu_sample(n_frames)
    fps, per_step, cores, which = run_cpu(flat, masks_t, com_t, args.steps, max(args.warmup, 1))
    sample = (f'{n_frames} frames ({flat.nbytes / 2**20:.0f} MiB) of the cfg2 stream per step, '
              'oracle port of the reference formulation: one GEMM per UDF (8 masks, then 3 CoM '
              f'masks), faster of torch.mm / numpy @ = {which}, data resident in host RAM')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': per_step * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'nav': list(NAV), 'sig': list(SIG),
                   'n_masks': N_MASKS, 'com': True, 'sample_frames': n_frames},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

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


def gpu_arm(args):
    import logging
    logging.getLogger('libertem_b200').setLevel(logging.ERROR)
    import torch
    import torch.distributed as dist
    from libertem_b200 import engine
    from libertem_b200.io import SyntheticDataSet, MemoryDataSet
    from libertem_b200.runner import UDFRunner, run_udf
    from libertem_b200.udf import ApplyMasksUDF, CoMUDF

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a GPU (there is no CPU fallback; use --impl reference '
                           'for the CPU arm)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    cpu_affinity = None
    if world > 1:
        cpu_affinity = _bind_near_gpu(torch, local_rank)
        dist.init_process_group('nccl', device_id=device)
    n_gpus = world

    nav = (NAV[0] * n_gpus, NAV[1])
    frames_per_rank = NAV[0] * NAV[1]
    total_frames = frames_per_rank * n_gpus
    k = SIG[0] * SIG[1]

    def dev_uniform(n, seed):
        return engine.synth_fill((n,), np.float32, seed, device).cpu().numpy()

    stack = bench_masks(SIG[0], SIG[1], N_MASKS, MASK_SEED, dev_uniform)
    ds = SyntheticDataSet(nav + SIG, np.float32, seed=DATA_SEED, num_partitions=n_gpus)
    parts = list(ds.get_partitions())
    my_parts = UDFRunner.my_partitions(parts, rank, n_gpus)
    ds.materialize(device, my_parts)          # 16 GiB per rank, resident in HBM
    torch.cuda.synchronize()

    def make_udfs():
        return [ApplyMasksUDF(mask_factories=lambda: stack, mask_count=N_MASKS,
                              mask_dtype=np.float32, use_sparse=False), CoMUDF()]

    runner = UDFRunner(make_udfs())

    def step():
        runner.run_for_dataset(ds, device=device, finalize=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if os.environ.get('LTB_PROFILE'):
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats('cumulative').print_stats(35)

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.lines and time.time() - t_wait < 5.0:   # wait for the first sample
        time.sleep(0.01)
    for _ in range(3):          # keep the GPU busy right up to the timed region
        step()
    engine.launch_count(reset=True)
    engine.EVENT_LOG = []
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = engine.launch_count()
    ev_log, engine.EVENT_LOG = engine.EVENT_LOG, None
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1)
    if world > 1:
        t = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = total_frames / (ms_per_step * 1e-3)

    # roofline of the dominant kernel, timed live on the launching stream
    kernel_ms = [a.elapsed_time(b) for a, b, *_ in ev_log]
    bytes_per_launch = float(np.mean([f * kk * isz for _, _, f, kk, isz in ev_log]))
    kern_ms = float(np.mean(kernel_ms))
    achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                'kernel': {1: 'k1_dense_tma_kernel (even/odd tile)', 3: 'k1_pair_kernel<float> '
                           '(mask-pair tile)', 6: 'k6_tensor_kernel (tcgen05 kind::tf32, '
                           'split-TF32, TMEM accumulators)'}.get(engine.last_kernel(), 'generic'),
                'kernel_ms': kern_ms,
                'algorithmic_bytes_per_launch': bytes_per_launch,
                'kernel_share_of_step': kern_ms * len(ev_log) / args.steps / ms_per_step}
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of this
    # command (cfg2 shape); None for other workloads
    prof = os.path.join(ROOT, 'profiles',
                        'k6_traffic.json' if engine.last_kernel() == 6 else 'k1_traffic.json')
    if os.path.exists(prof) and args.workload == 'cfg2':
        try:
            roofline['traffic'] = json.load(open(prof)).get('dram_bytes_per_launch')
        except Exception:
            pass

    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': n_gpus,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic (counter-based hash, uniform [0,1), seed %d)' % DATA_SEED,
        'config': {'workload': WORKLOAD, 'nav': list(nav), 'sig': list(SIG), 'n_masks': N_MASKS,
                   'com': True, 'fused_columns': N_MASKS + 3, 'partitions_per_gpu': 1,
                   'sig_bytes_per_frame': k * 4,
                   'frames_per_gpu': frames_per_rank,
                   'l2': 'inputs %.1f GB per GPU >> 126 MB L2, streamed once per step' %
                         (frames_per_rank * k * 4 / 1e9),
                   'merge': 'nccl all_gather of nav buffers inside the step' if world > 1
                   else 'device-side copy into the nav-shaped buffers',
                   'cpu_affinity': cpu_affinity},
        'hbm_gbs': total_frames * k * 4 / (ms_per_step * 1e-3) / 1e9 / n_gpus,
        'roofline': roofline, 'clocks': clocks, 'gpu_launches': int(launches),
    }

    if rank == 0 and world == 1 and not args.no_e2e:
        line['e2e'] = e2e_leg(args, ds, my_parts[0], stack, device)
    elif world > 1:
        line['e2e'] = e2e_multi(args, ds, my_parts[0], stack, device, dist, total_frames)
    if rank == 0 and world == 1 and not args.no_cpu:
        n_frames = 2048
        flat, masks_t, com_t = cpu_sample(n_frames)
        fps, per_step, cores, which = run_cpu(flat, masks_t, com_t, 3, 1)
        line['cpu_baseline'] = {
            'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'sample': f'{n_frames} frames (512 MiB) of the same stream, 3 passes after 1 warm-up, '
                      f'one GEMM per UDF (8 masks + 3 CoM masks) as the reference does, faster '
                      f'of torch.mm / numpy @ = {which}'}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _host_copy(ds, part, device):
    """the rank's shard as a pinned host array (filled by D2H from the resident device data)"""
    import torch
    t = ds.partition_tensor(part, device)
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t)
    torch.cuda.synchronize()
    return host


def e2e_leg(args, ds, part, stack, device):
    """end to end through run_udf(): HOST pinned input, H2D of every frame + D2H / get_results
    of every result buffer inside the timed region"""
    import torch
    from libertem_b200.io import MemoryDataSet
    from libertem_b200.runner import run_udf
    from libertem_b200.udf import ApplyMasksUDF, CoMUDF
    host = _host_copy(ds, part, device)
    hds = MemoryDataSet(data=host.reshape(NAV + SIG), num_partitions=1, sig_dims=2, pin=False)
    steps = min(args.steps, 3)

    def one():
        res = run_udf(hds, [ApplyMasksUDF(mask_factories=lambda: stack, mask_count=N_MASKS,
                                          mask_dtype=np.float32, use_sparse=False), CoMUDF()],
                      device=device)
        return res[0]['intensity'].raw_data, res[1]['field'].raw_data

    one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        a, b = one()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    d2h = NAV[0] * NAV[1] * (N_MASKS + 3) * 4
    return {'value': NAV[0] * NAV[1] / dt, 'unit': 'frames/s',
            'h2d_bytes_per_step': int(host.numel() * host.element_size()),
            'd2h_bytes_per_step': int(d2h), 'steps': steps, 'ms_per_step': dt * 1e3,
            'note': 'run_udf on a pinned host dataset: double-buffered H2D tiles overlapped with '
                    'the kernel; includes CoM get_results on the host'}


def e2e_multi(args, ds, part, stack, device, dist, total_frames):
    import torch
    from libertem_b200.io import MemoryDataSet
    from libertem_b200.runner import run_udf
    from libertem_b200.udf import ApplyMasksUDF, CoMUDF
    host = _host_copy(ds, part, device)
    hds = MemoryDataSet(data=host.reshape(NAV + SIG), num_partitions=1, sig_dims=2, pin=False)
    steps = min(args.steps, 2)

    def one():
        # every rank runs its shard from host memory; results come back per rank
        from libertem_b200.runner import UDFRunner
        r = UDFRunner([ApplyMasksUDF(mask_factories=lambda: stack, mask_count=N_MASKS,
                                     mask_dtype=np.float32, use_sparse=False), CoMUDF()])
        # shard-local run (no collective on the e2e leg; the device-timed leg has it)
        saved = UDFRunner._dist
        UDFRunner._dist = staticmethod(lambda: None)
        try:
            res = r.run_for_dataset(hds, device=device).buffers
        finally:
            UDFRunner._dist = saved
        return res[0]['intensity'].raw_data

    one()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / steps
    t = torch.tensor([dt], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    return {'value': total_frames / dt, 'unit': 'frames/s',
            'h2d_bytes_per_step': int(host.numel() * host.element_size()) * dist.get_world_size(),
            'd2h_bytes_per_step': int(total_frames * (N_MASKS + 3) * 4), 'steps': steps,
            'ms_per_step': dt * 1e3,
            'note': 'each rank streams its shard from pinned host memory (PCIe-bound)'}


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
    ap.add_argument('--workload', default='cfg2', choices=['cfg2', 'cfg5'])
    args = ap.parse_args()
    set_workload(args.workload)
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
