"""Turns gpurun_out/<round>_{launches.csv,step.ncu-rep,k_knn2_partial.ncu-rep} (written by tools/make_profiles.sh on the GPU box) into the
committed text summaries under profiles/: one file per kernel, the launch list, and <round>_traffic.json (DRAM bytes per stage and step,
read by bench.py for roofline.traffic). Run here (needs the ncu CLI, no GPU). usage: summarize_profiles.py r02 [frames_per_step]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else 'r02'
FRAMES = int(sys.argv[2]) if len(sys.argv) > 2 else 128      # frames in one captured launch (half of the bench command's --batch)
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
os.makedirs(P, exist_ok=True)

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
TIME = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
CMD = 'python bench.py --steps 3 --warmup 3 --batch 256 --skip-cpu --skip-stereo --skip-configs --skip-guided --knn-steps 1 --knn-queries 131072 --knn-train-per-gpu 262144'


def short(name):
    return name.split('(')[0].replace('<unnamed>::', '').replace('void ', '').strip()


def launches():
    f = os.path.join(G, f'{R}_launches.csv')
    if not os.path.exists(f):
        return
    rows = [r for r in csv.reader(open(f, errors='replace')) if len(r) > 5]
    h = rows[0]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(short(r[ki]), []).append(float(r[vi].replace(',', '')) * TIME.get(r[ui], 1e-3))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, f'{R}_launches.txt'), 'w') as o:
        o.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES), window of {len(rows) - 1} launches\n')
        o.write(f'# command: {CMD} (tools/make_profiles.sh)\n')
        o.write(f'{"kernel":44s} {"n":>4s} {"sum_us":>10s} {"avg_us":>9s} {"share":>7s}\n')
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            o.write(f'{k[:44]:44s} {len(v):4d} {sum(v):10.1f} {sum(v) / len(v):9.1f} {sum(v) / tot:7.3f}\n')
    os.replace(f, os.path.join(P, f'{R}_launches.csv'))
    print(open(os.path.join(P, f'{R}_launches.txt')).read())


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


STAGE = {'k_pyramid_strip': 'pyramid', 'k_level_strip<32, 1, 0>': 'blur', 'k_level_strip<32, 0, 1>': 'fast', 'k_fast_cells2': 'fast',
         'k_quadtree': 'quadtree', 'k_orient_describe2': 'describe'}


def stage_of(name):
    for k, v in STAGE.items():
        if k in name:
            return v
    return None


def step():
    rep = os.path.join(G, f'{R}_step.ncu-rep')
    if not os.path.exists(rep):
        return
    h, u, rows = raw(rep)
    ki = h.index('Kernel Name')
    groups = collections.OrderedDict()
    for r in rows:
        groups.setdefault(short(r[ki]), []).append(r)
    traffic = {}
    for name, rs in groups.items():
        fn = name.replace('<', '_').replace('>', '').replace(', ', '_').replace('::', '_').replace(' ', '')
        with open(os.path.join(P, f'{R}_{fn}.txt'), 'w') as o:
            o.write(f'# ncu --set full --clock-control none --import-source on, {len(rs)} launch(es) of {name} inside one device-resident half-batch ({FRAMES} C1 frames: an un-instrumented 256-frame step runs as two halves on two streams)\n')
            o.write(f'# command: {CMD} (tools/make_profiles.sh); times are cold-cache and serialised\n')
            for n, r in enumerate(rs):
                if len(rs) > 1:
                    o.write(f'## launch {n + 1} of {len(rs)}\n')
                for k in KEYS:
                    if k in h:
                        i = h.index(k)
                        o.write(f'{k:90s} {r[i]:>18s} {u[i]}\n')
        st = stage_of(name)
        if st:
            t = traffic.setdefault(st, {'dram_bytes_per_step': 0.0, 'kernel_us_per_step': 0.0, 'frames_per_step': FRAMES, 'kernels': []})
            for r in rs:
                rd, wr = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
                t['dram_bytes_per_step'] += float(r[rd].replace(',', '')) * UNIT[u[rd]] + float(r[wr].replace(',', '')) * UNIT[u[wr]]
                ti = h.index('gpu__time_duration.sum')
                t['kernel_us_per_step'] += float(r[ti].replace(',', '')) * TIME[u[ti]]
            t['kernels'].append(f'{name} x{len(rs)}')
            # issue-slot and pipe utilisation of the stage's longest launch: what actually bounds these integer kernels
            big = max(rs, key=lambda r: float(r[h.index('gpu__time_duration.sum')].replace(',', '')))
            pipes = t.setdefault('pipes', {})
            pipes[name] = {k2: float(big[h.index(k1)].replace(',', '')) for k2, k1 in (
                ('issue_active_pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                ('alu_pct', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
                ('fma_pct', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
                ('fmaheavy_pct_elapsed', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'),
                ('lsu_pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
                ('dram_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')) if k1 in h}
        print(name, len(rs), 'launches')
    json.dump(traffic, open(os.path.join(P, f'{R}_traffic.json'), 'w'), indent=1)
    print(json.dumps(traffic, indent=1))


def knn():
    rep = os.path.join(G, f'{R}_k_knn2_partial.ncu-rep')
    if not os.path.exists(rep):
        return
    h, u, rows = raw(rep)
    with open(os.path.join(P, f'{R}_k_knn2_partial.txt'), 'w') as o:
        o.write('# ncu --set full --clock-control none --import-source on -k regex:k_knn2_partial -c 1 (131072 queries x 262144 train rows)\n')
        for k in KEYS:
            if k in h:
                i = h.index(k)
                o.write(f'{k:90s} {rows[-1][i]:>18s} {u[i]}\n')


launches()
step()
knn()
