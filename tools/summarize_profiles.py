"""Turns gpurun_out/<round>_*.{csv,ncu-rep} (written by tools/make_profiles.sh on the GPU box) into the committed text
summaries under profiles/. Run here (needs the ncu CLI, no GPU)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else 'r01'
G = os.path.join(ROOT, 'gpurun_out')
P = os.path.join(ROOT, 'profiles')
os.makedirs(P, exist_ok=True)

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']


def launches():
    f = os.path.join(G, f'{R}_launches.csv')
    if not os.path.exists(f):
        return
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    h = rows[0]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split('(')[0].replace('<unnamed>::', ''), []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, f'{R}_launches.txt'), 'w') as o:
        o.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES), window of {len(rows) - 1} launches\n')
        o.write('# command: python bench.py --steps 3 --warmup 3 --batch 256 --skip-cpu --knn-steps 1 --knn-queries 131072 --knn-train-per-gpu 262144 (tools/profile_one.sh)\n')
        o.write(f'{"kernel":28s} {"n":>4s} {"sum_us":>10s} {"avg_us":>9s} {"share":>7s}\n')
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            o.write(f'{k[:28]:28s} {len(v):4d} {sum(v) / 1e3:10.1f} {sum(v) / len(v) / 1e3:9.1f} {sum(v) / tot:7.3f}\n')
    os.replace(f, os.path.join(P, f'{R}_launches.csv'))
    print(open(os.path.join(P, f'{R}_launches.txt')).read())


def full(kernel, frames_per_launch=None):
    rep = os.path.join(G, f'{R}_{kernel}.ncu-rep')
    if not os.path.exists(rep):
        return None
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[-1]
    d = {}
    with open(os.path.join(P, f'{R}_{kernel}.txt'), 'w') as o:
        o.write(f'# ncu --set full --clock-control none --import-source on -k regex:{kernel} -c 1 (one launch)\n')
        for k in KEYS:
            if k in h:
                i = h.index(k)
                o.write(f'{k:90s} {v[i]:>18s} {u[i]}\n')
                d[k] = (v[i], u[i])
    print(open(os.path.join(P, f'{R}_{kernel}.txt')).read())
    return d


def to_bytes(val, unit):
    x = float(val.replace(',', ''))
    return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


launches()
tj = os.path.join(P, f'{R}_traffic.json')
traffic = json.load(open(tj)) if os.path.exists(tj) else {}      # captures arrive one gpurun call at a time: keep what is already there
NOTES = {'k_guided_search': 'one search, 1000 keypoints x 1000 map points (tools/guided_probe.py), a cluster of 8 CTAs',
         'k_remap_to_l0': 'one launch of 256 frames 752x480 (bench.py remap block)',
         'k_stereo_match': 'one launch of 64 C2 stereo pairs (bench.py stereo block)',
         'k_bow_descend': 'one launch of 256 frames x 1000 descriptors through the k 10 / L 6 vocabulary (bench.py bow block)',
         'k_bow_finalize': 'one launch of 256 frames (bench.py bow block), one CTA per frame'}
for k in ('k_fast_cells', 'k_gauss7', 'k_pyramid_resize', 'k_orient_describe', 'k_quadtree', 'k_knn2_partial', 'k_guided_search', 'k_remap_to_l0',
          'k_stereo_match', 'k_bow_descend', 'k_bow_finalize', 'k_best_in_windows'):
    d = full(k)
    if d and 'dram__bytes_read.sum' in d:
        traffic[k] = {'dram_bytes_per_launch': to_bytes(*d['dram__bytes_read.sum']) + to_bytes(*d['dram__bytes_write.sum']),
                      'launch_us': float(d['gpu__time_duration.sum'][0].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}[d['gpu__time_duration.sum'][1]],
                      'grid': d['launch__grid_size'][0],
                      'note': NOTES.get(k, 'one launch of the bench default batch (256 frames); k_gauss7 / k_pyramid_resize: one level of it')}
if traffic:
    json.dump(traffic, open(os.path.join(P, f'{R}_traffic.json'), 'w'), indent=1)
    print(traffic)
