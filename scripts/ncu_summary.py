#!/usr/bin/env python
"""Per-kernel summary of an ncu report (the metrics DESIGN.md quotes) + top source lines per kernel.
    python scripts/ncu_summary.py X.ncu-rep [top_lines]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('kernel:', r[hdr.index('Kernel Name')])
    for w in want:
        if w in hdr:
            print('  %-70s %s %s' % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
    st = {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): float(r[i])
          for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')}
    print('  stalls per issue: ' + ', '.join('%s=%.2f' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:7]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True,
                     text=True).stdout
fname = func = hdr = None
agg = {}
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        func = r[1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or not r[0].isdigit():
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        s, i = int(d['# Samples']), int(d['Instructions Executed'])
    except (KeyError, ValueError):
        continue
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and '(' not in k and v.isdigit() and int(v)}
    agg.setdefault(func, []).append((s, i, fname, r[0], r[1].strip()[:95], st))
for func, out in agg.items():
    tot = sum(o[0] for o in out) or 1
    toti = sum(o[1] for o in out) or 1
    print('== %s: %d samples, %d warp instructions' % (func[:90], tot, toti))
    for s, i, f, ln, text, st in sorted(out, key=lambda o: -o[0])[:top]:
        st3 = ' '.join('%s=%d' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print('%5.1f%% smp %5.1f%% inst  %s:%s  %s   [%s]' % (100.0 * s / tot, 100.0 * i / toti, f, ln, text, st3))
