#!/usr/bin/env python3
"""SASS evidence for the built library (run here: cuobjdump cross-reads sm_100a cubins): library-wide mnemonic counts and, per
extractor kernel, TMA tile loads, packed min/max, dot-product and async-copy instructions. usage: sass_summary.py > profiles/<round>_sass_summary.txt"""
import re, subprocess, collections
LIB = 'orb_slam2_refactored_b200/lib/liborbx_b200.so'
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
pats = ['UTMALDG', r'SYNCS\.', r'LDGSTS\.E\.BYPASS\.128', r'IDP\.4A', r'IDP\.2A', r'VIMNMX3\.U16x2', r'VIMNMX\.U16x2', 'PRMT', 'POPC', 'UCGABAR', r'REDG|RED\.', r'LD\.E\.STRONG\.GPU|LDG\.E\.STRONG\.GPU', 'UTC.*MMA', 'LDTM', 'HMMA']
print(f'# SASS evidence for {LIB} (cuobjdump -sass, sm_100a cubins), round 2 final')
print('# command: cuobjdump -sass <lib> | grep -c <mnemonic> (tools/sass_summary.py)')
m = re.search(r'arch = (\S+)', out)
print('arch =', m.group(1) if m else '?')
for p in pats:
    print(f'{p:32s} {len(re.findall(p, out)):5d}')
print('\n# per kernel (extractor): TMA tile loads, packed min/max, dot-product instructions')
cur = None; per = collections.OrderedDict()
for l in out.split('\n'):
    f = re.search(r'Function : (\S+)', l)
    if f:
        cur = f.group(1); per[cur] = collections.Counter(); continue
    if cur:
        for k, p in (('UTMALDG', 'UTMALDG'), ('VIMNMX', 'VIMNMX'), ('IDP', r'IDP\.'), ('LDGSTS', 'LDGSTS')):
            if re.search(p, l): per[cur][k] += 1
for k in sorted(per):
    if any(t in k for t in ('k_fast', 'k_level_strip', 'k_pyramid', 'k_orient')):
        name = re.sub(r'^_ZN\d+_GLOBAL__N__[0-9a-f_]+orbx_extract_cu_[0-9a-f]+\d*', '', k)
        c = per[k]
        print(f'{name[:96]:96s} UTMALDG {c["UTMALDG"]} VIMNMX {c["VIMNMX"]} IDP {c["IDP"]} LDGSTS {c["LDGSTS"]}')
print('\n# No UTC*MMA / LDTM / HMMA: the path is byte / integer work, tensor cores are not used (north_star).')
