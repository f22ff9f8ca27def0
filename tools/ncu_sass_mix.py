"""Opcode mix and top stall sites from an `ncu --page source --csv --print-source sass` export (one or more kernels)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
kern = None; hdr = None
agg = {}
for r in rows:
    if r and r[0] == 'Kernel Name':
        kern = r[1][:70]; agg[kern] = {'ops': collections.Counter(), 'samples': collections.Counter(), 'lines': []}; hdr = None; continue
    if r and r[0] == 'Address':
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or kern is None or len(r) < len(hdr): continue
    src = r[hdr['Source']].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0] if not op.startswith('UTC') else op
    n = int(r[hdr['Instructions Executed']] or 0); s = int(r[hdr['# Samples']] or 0)
    agg[kern]['ops'][op] += n; agg[kern]['samples'][op] += s
    agg[kern]['lines'].append((s, n, src))
for k, a in agg.items():
    tot = sum(a['ops'].values()); ts = sum(a['samples'].values())
    print('==', k, 'instr', tot, 'samples', ts)
    for op, n in a['ops'].most_common(28):
        print(f'   {op:14s} {n:10d} {100.0*n/tot:5.1f}%   samples {100.0*a["samples"][op]/max(ts,1):5.1f}%')
    print('   -- top sampled instructions')
    for s, n, src in sorted(a['lines'], reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
        print(f'   {100.0*s/max(ts,1):5.1f}%  x{n:8d}  {src[:90]}')
