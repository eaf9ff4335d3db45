"""Hot spots of an ncu source-page CSV (SASS view): instruction mix by opcode weighted by executed count,
and the top stall-sample instructions."""
import csv, io, subprocess, sys, collections

# usage: ncu_hot.py <source-page.csv | report.ncu-rep> [top N] [kernel-name regex, .ncu-rep only]
if sys.argv[1].endswith('.ncu-rep'):
    cmd = ['ncu', '-i', sys.argv[1], '--page', 'source', '--csv']
    if len(sys.argv) > 3:
        cmd += ['--kernel-name', 'regex:' + sys.argv[3]]
    rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
else:
    rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samples = collections.Counter(); total = 0
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix['Source']].strip()
    n = int(r[ix['Instructions Executed']] or 0)
    s = int(r[ix['# Samples']] or 0)
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '')
    op = op.split('.')[0]
    ops[op] += n; total += n
    recs.append((s, n, r[ix['Address']], src))
print('total warp-instructions', total)
for op, n in ops.most_common(28):
    print(f'{op:12s} {n:12d} {100*n/total:5.1f}%')
print('--- top stall samples')
tot_s = sum(r[0] for r in recs)
for s, n, a, src in sorted(recs, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f'{s:6d} ({100*s/tot_s:4.1f}%) exec {n:9d}  {a[-5:]}  {src[:90]}')
