#!/usr/bin/env python
"""Digest of one `ncu --set full` report: headline metrics + instruction mix per detection + hot code regions.
usage: ncu_digest.py report.ncu-rep N_detections"""
import collections, csv, subprocess, sys, io
rep, N = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active']
d = dict(zip(hdr, zip(units, vals)))
for k in want:
    if k in d:
        print('%-70s %-10s %s' % (k, d[k][0], d[k][1]))
for k in hdr:
    if 'issue_stalled' in k and 'per_issue_active' in k and float(d[k][1] or 0) > 0.3:
        print('%-90s %s' % (k.replace('smsp__average_warps_issue_stalled_', 'stall '), d[k][1]))
try:
    rd = float(d['dram__bytes_read.sum'][1]) * {'Gbyte': 1e9, 'Mbyte': 1e6}[d['dram__bytes_read.sum'][0]]
    wr = float(d['dram__bytes_write.sum'][1]) * {'Gbyte': 1e9, 'Mbyte': 1e6}[d['dram__bytes_write.sum'][0]]
    print('DRAM bytes per detection: %.1f (read %.1f + write %.1f)' % ((rd + wr) / N, rd / N, wr / N))
    print('warp instructions per detection: %.1f' % (float(d['smsp__inst_executed.sum'][1]) / N))
except Exception as e:
    print('n/a', e)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ie, isrc, isamp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
data = [(r[isrc], int(r[ie]), int(r[isamp])) for r in rows[2:] if len(r) > ie]
op = collections.Counter()
for s, e, sm in data:
    t = s.split()
    o = t[1] if t[0].startswith('@') else t[0]
    op[o.split('.')[0]] += e
print('instruction mix per detection:', ', '.join('%s %.2f' % (k, v / N) for k, v in op.most_common(16)))
tot_s = sum(x[2] for x in data) or 1
# regions of equal execution count
seg, start, prev = [], 0, None
for i, (s, e, sm) in enumerate(data):
    if prev is not None and abs(e - prev) > 0.3 * max(prev, 1):
        seg.append((start, i - 1, prev)); start = i
    prev = e
seg.append((start, len(data) - 1, prev))
print('code regions (SASS index range, executions, instr/det, %% of stall samples):')
for a, b, c in seg:
    w = (b - a + 1) * c / N
    sm = sum(x[2] for x in data[a:b + 1])
    if w > 0.4 or sm / tot_s > 0.02:
        print('  %5d-%5d x%-10d %6.2f  %5.1f%%   %s' % (a, b, c, w, 100 * sm / tot_s, data[a][0][:60]))
