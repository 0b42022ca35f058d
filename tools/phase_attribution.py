#!/usr/bin/env python3
"""Per-phase instruction attribution of one kernel: joins the executed-instruction counts of an ncu capture (--import-source on) with the
line table of the SAME build (nvdisasm -gi on the cubin inside liborbx_b200.so) and sums warp-instructions per source-line range.
The capture and the library must come from the same build (run right after tools/make_profiles.sh, before rebuilding).

usage: phase_attribution.py <report.ncu-rep> <kernel substring> <unit count> <unit name> markers|<phases.json> [--launch N] [--mangled <substring of the symbol in the cubin, for templates>]
phases: `//@phase <name>` ... `//@end` comment markers in the kernel sources, or a json list of [phase name, file suffix, first line, last line];
an instruction belongs to the INNERMOST location of its inline chain that lies in a phase (so a helper's body counts as the helper's phase)."""
import csv, json, os, re, subprocess, sys, tempfile, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'orb_slam2_refactored_b200', 'lib', 'liborbx_b200.so')


def line_table(kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', LIB], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith('orbx_extract.') and f.endswith('.cubin')][0]
    out = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    rows, cur, inside, fresh = [], [], False, True
    for ln in out.split('\n'):
        if ln.startswith('\t.section\t.text.'):
            inside = kernel in ln
            cur, fresh = [], True
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            # one //## line per inline level, innermost first; a new group starts after an instruction
            if fresh:
                cur, fresh = [], False
            cur.append((m.group(1), int(m.group(2))))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            rows.append((ln.split('*/', 1)[1].strip().rstrip(';').strip(), list(cur)))
            fresh = True
    return rows


def counts(rep, kernel, launch):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    ks, cur = [], None
    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == 'Kernel Name':
            cur = {'name': row[1], 'hdr': None, 'rows': []}
            ks.append(cur)
        elif cur is not None and cur['hdr'] is None:
            cur['hdr'] = row
        elif cur is not None:
            cur['rows'].append(row)
    plain = lambda t: re.sub(r'[^A-Za-z0-9_]', '', t.replace('(int)', '').replace('(bool)', ''))
    ks = [k for k in ks if plain(kernel) in plain(k['name'])]
    k = ks[launch]
    i_src, i_ex = k['hdr'].index('Source'), k['hdr'].index('Instructions Executed')
    return [(r[i_src].strip(), int(r[i_ex])) for r in k['rows']]


def marker_phases():
    """//@phase <name> ... (next marker or //@end) in the kernel sources -> [name, file suffix, first line, last line]"""
    out = []
    for fn in ('orbx_extract.cu', 'orbx_strip.cuh', 'orbx_describe.cuh'):
        cur = None
        for n, ln in enumerate(open(os.path.join(ROOT, 'orb_slam2_refactored_b200', 'csrc', fn)), 1):
            t = ln.strip()
            if t.startswith('//@phase ') or t.startswith('//@end'):
                if cur:
                    out.append([cur[0], fn, cur[1], n - 1])
                cur = (t[len('//@phase '):], n + 1) if t.startswith('//@phase ') else None
        if cur:
            out.append([cur[0], fn, cur[1], 10 ** 9])
    return out


def main():
    rep, kernel, units, unit_name, spec = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4], sys.argv[5]
    launch = int(sys.argv[sys.argv.index('--launch') + 1]) if '--launch' in sys.argv else 0
    phases = marker_phases() if spec == 'markers' else json.load(open(spec))
    mangled = sys.argv[sys.argv.index('--mangled') + 1] if '--mangled' in sys.argv else kernel
    lt = line_table(mangled)
    ct = counts(rep, kernel, launch)
    if len(lt) != len(ct):
        sys.exit(f'instruction count differs: {len(lt)} in the library vs {len(ct)} in the capture — not the same build')
    agg = collections.OrderedDict()
    for p_ in phases:
        agg.setdefault(p_[0], 0)
    agg['(other)'] = 0
    for (sass, chain), (_, n) in zip(lt, ct):
        hit = '(other)'
        for f, l in chain:                    # innermost location first
            for name, suffix, lo, hi in phases:
                if f.endswith(suffix) and lo <= l <= hi:
                    hit = name
                    break
            if hit != '(other)':
                break
        agg[hit] += n
    tot = sum(agg.values())
    print(f'# {kernel}: {tot} warp-instructions executed in the captured launch = {tot / units:.1f} per {unit_name} ({units:.0f} {unit_name}s)')
    for name, n in agg.items():
        print(f'{name:62s} {n:12d} {n / units:9.1f} per {unit_name} {100.0 * n / tot:6.1f} %')


if __name__ == '__main__':
    main()
