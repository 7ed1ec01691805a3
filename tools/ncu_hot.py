#!/usr/bin/env python3
"""Top stall-sample instructions of a kernel in an .ncu-rep (test/dev tooling).  usage: ncu_hot.py rep [kernel-index] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kid = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': []}; blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(row)
b = blocks[kid]
hdr = b['rows'][0]
si, ai, ei = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
rows = []
for i, r in enumerate(b['rows'][1:]):
    try: rows.append((int(r[ai]), i, r[si].strip(), int(r[ei]), r))
    except ValueError: pass
tot = sum(r[0] for r in rows)
print(b['name'][:90], 'total samples', tot, 'instr rows', len(rows))
for s, i, t, e, r in sorted(rows, reverse=True)[:top]:
    why = sorted(((int(r[c]) if r[c].isdigit() else 0, hdr[c]) for c in stall_cols), reverse=True)[:2]
    print('%6d %5.1f%% #%4d exec=%9d  %-58s %s' % (s, 100.0 * s / max(tot, 1), i, e, t[:58], ' '.join('%s=%d' % (n[6:], v) for v, n in why if v)))
