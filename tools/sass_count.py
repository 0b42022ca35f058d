#!/usr/bin/env python3
"""Static SASS opcode histogram of one kernel in liborbx_b200.so (cuobjdump -sass), optionally only between two byte offsets.
usage: sass_count.py <kernel-substring> [--range lo hi] [--list]"""
import subprocess, sys, re, collections
lib = 'orb_slam2_refactored_b200/lib/liborbx_b200.so'
name = sys.argv[1]
lo = hi = None
if '--range' in sys.argv:
    i = sys.argv.index('--range'); lo = int(sys.argv[i + 1], 16); hi = int(sys.argv[i + 2], 16)
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cur = None; hist = collections.Counter(); lines = []
for l in out.split('\n'):
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1); continue
    if cur and name in cur:
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m:
            off = int(m.group(1), 16); ins = m.group(2).strip()
            if lo is not None and not (lo <= off < hi): continue
            lines.append((off, ins))
            t = ins.split()
            op = t[1] if t[0].startswith('@') else t[0]
            hist[op.split('.')[0]] += 1
tot = sum(hist.values())
print(name, 'instructions:', tot)
for op, n in hist.most_common(): print(f'  {op:10s} {n}')
if '--list' in sys.argv:
    for off, ins in lines: print(f'{off:05x}  {ins}')
