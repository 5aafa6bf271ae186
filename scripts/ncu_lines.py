#!/usr/bin/env python
"""Aggregate an ncu `--page source --csv --print-source cuda,sass` export by CUDA source line:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python scripts/ncu_lines.py src.csv [N]
Prints the top-N lines by stall samples with executed warp instructions and the dominant stall reasons."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
fname, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or r[0] in ('Function Name', ) or not r[0].isdigit():
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        samples = int(d['# Samples'])
        inst = int(d['Instructions Executed'])
    except (KeyError, ValueError):
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and '(' not in k and v.isdigit() and int(v)}
    out.append((samples, inst, fname, r[0], r[1].strip()[:100], stalls))
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print('total samples %d, warp instructions %d' % (tot, toti))
for s, i, f, ln, src, st in sorted(out, key=lambda o: -o[0])[:top]:
    st3 = ' '.join('%s=%d' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print('%5.1f%% smp %5.1f%% inst  %s:%s  %s   [%s]' % (100.0 * s / tot, 100.0 * i / toti, f, ln, src, st3))
