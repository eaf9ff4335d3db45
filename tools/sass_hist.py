"""Opcode histogram of the library's kernels from `cuobjdump -sass` (what proves tcgen05 / TMA / TMEM use: UTCHMMA =
tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor, SYNCS = mbarrier).
usage: python tools/sass_hist.py [kernel-name regex] > profiles/rN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

LIB = 'egt_b200/lib/libegt_b200.so'
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r'fused_(fwd|bwd)_kernel|wide_(fwd|bwd)_kernel|node_(qkv|out|bwd1|bwd2)_kernel|ffn_tc_(fwd|bwd)_kernel|peer_allreduce')
KEY = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMACCTL', 'SYNCS', 'ELECT', 'R2UR', 'MUFU', 'FFMA', 'FFMA2', 'HMMA',
       'LDS', 'STS', 'LDG', 'STG', 'RED', 'ATOMG', 'BAR']
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
name, hist = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and name:
        hist[name][m.group(1).split('.')[0]] += 1
for fn in sorted(hist):
    dem = subprocess.run(['cu++filt', fn], capture_output=True, text=True).stdout.strip() or fn
    if not pat.search(dem):
        continue
    h = hist[fn]
    print(f'{dem[:150]}')
    print(f'    {sum(h.values())} SASS instructions; ' + ', '.join(f'{k} {h[k]}' for k in KEY if h[k]))
