"""Summarise ncu artefacts into the tracked profiles/ directory.

    python scripts/summarize_ncu.py full <rep.ncu-rep> <out.txt> [traffic.json]
    python scripts/summarize_ncu.py launches <launches.csv> <out.txt> [skip_launches]
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
    'l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'lts__t_bytes.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
]


def full(rep, out, traffic=None):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        vals = dict(zip(hdr, r))
        lines.append('kernel: %s  grid %s block %s' % (vals.get('Kernel Name'),
                                                      vals.get('Grid Size'),
                                                      vals.get('Block Size')))
        for i, h in enumerate(hdr):
            if h in WANT:
                lines.append('  %-80s %-10s %s' % (h, units[i], r[i]))
        try:
            rd = float(vals['dram__bytes_read.sum'])
            wr = float(vals['dram__bytes_write.sum'])
            scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}
            u = dict(zip(hdr, units))
            tot = rd * scale[u['dram__bytes_read.sum']] + wr * scale[u['dram__bytes_write.sum']]
            lines.append('  dram traffic (read+write) per launch: %.0f bytes' % tot)
            if traffic:
                json.dump({'kernel': vals.get('Kernel Name'), 'dram_bytes_per_launch': tot,
                           'source': rep}, open(traffic, 'w'), indent=1)
        except Exception as e:  # pragma: no cover
            lines.append('  (traffic unavailable: %s)' % e)
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


def launches(path, out, skip=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    rows = rows[int(skip):]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split('(')[0]
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += float(r[-1])
    total = sum(v[1] for v in agg.values())
    lines = ['ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: '
             'compare SHARES)', 'launches after skipping %s: %d, total %.3f ms' %
             (skip, len(rows), total / 1e6), '%8s %12s %7s  kernel' % ('count', 'sum_us', 'share')]
    for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append('%8d %12.1f %6.2f%%  %s' % (c, ns / 1e3, 100 * ns / total, name))
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    if sys.argv[1] == 'full':
        full(*sys.argv[2:])
    else:
        launches(*sys.argv[2:])
