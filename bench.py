#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: ORB extraction frames/s at 640x480 / 1000 keypoints
(BASELINE.json configs[0], TUM1.yaml settings: scale 1.2, 8 levels, FAST 20/7), plus Hamming kNN Gpairs/s on the
per-GPU share of configs[4] (1 M queries x 1.25 M train rows per GPU, train-sharded, NCCL all-gather merge at N > 1).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code on the host cores

A step = one pass of the extractor over one batch of synthetic frames per GPU. One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from orb_slam2_refactored_b200 import synth  # noqa: E402

W, H, NFEATURES = 640, 480, 1000
METRIC = 'ORB extract frames/s @640x480 1000kp'


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def make_frames(n, seed0):
    """n distinct synthetic frames from 24 generated base frames (vertical flips / row rolls keep the statistics)."""
    base = [synth.image(seed0 + s, W, H) for s in range(min(n, 24))]
    out = np.empty((n, H, W), np.uint8)
    for i in range(n):
        b = base[i % len(base)]
        k = i // len(base)
        out[i] = b if k == 0 else np.roll(b[::-1] if k & 1 else b, 37 * k, axis=0)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '20',
                                          '-i', str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# the reference's CPU implementation on the host cores (oracle/_ref when it was built, else the restatement)
# ---------------------------------------------------------------------------------------------------------------------
def load_cpu_reference():
    from oracle import bindings
    try:
        bindings.build()
    except Exception as e:   # no compiler on the box: use what travelled
        log('oracle build skipped:', e)
    for kind, native in (('ref', True), ('ref', False), ('port', True), ('port', False)):
        try:
            return bindings.Oracle(kind, native=native), ('reference' if kind == 'ref' else 'port'), native
        except (FileNotFoundError, OSError):
            continue
    raise RuntimeError('no CPU oracle library available')


def cpu_extract_rate(o, frames, threads, per_thread):
    """frames/s of Extract with one frame per thread at a time (the reference runs left/right on 2 threads,
    src/System.cc:449-452); steady_clock around the calls as the example drivers do (mono_tum.cc:81-88)."""
    exs = [o.extractor(NFEATURES) for _ in range(threads)]
    for t in range(threads):
        exs[t].extract(frames[t % len(frames)])      # warm-up
    done = [0] * threads

    def work(t):
        for i in range(per_thread):
            exs[t].extract(frames[(t * per_thread + i) % len(frames)])
            done[t] += 1
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, dt


def cpu_config_rate(o, name, threads, seconds=4.0):
    """frames/s of the CPU Extract on frames of BASELINE.json config `name`, one frame per thread at a time, a bounded sample of about `seconds`."""
    c = synth.CONFIGS[name]
    frames = [synth.image(700 + s, c['w'], c['h']) for s in range(min(threads, 8))]
    exs = [o.extractor(c['nfeatures']) for _ in range(threads)]
    t0 = time.perf_counter()
    exs[0].extract(frames[0])
    one = time.perf_counter() - t0                      # single-thread time of one frame (also the warm-up of extractor 0)
    per_thread = max(1, int(seconds / max(one, 1e-3)))
    done = [0] * threads

    def work(t):
        for i in range(per_thread):
            exs[t].extract(frames[(t + i) % len(frames)])
            done[t] += 1
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, f'{sum(done)} frames {c["w"]}x{c["h"]} / {c["nfeatures"]} kp, one frame per thread at a time, {dt:.1f} s'


def cpu_stereo_rate(o, threads, seconds=4.0):
    """stereo pairs/s of Extract(left) + Extract(right) + ComputeStereoMatches on the CPU at the C2 shape, one pair per thread at a time."""
    c = synth.CONFIGS['C2']
    pairs = [synth.stereo_pair(800 + s, c['w'], c['h']) for s in range(min(threads, 4))]
    exs = [(o.extractor(c['nfeatures']), o.extractor(c['nfeatures'])) for _ in range(threads)]

    def one_pair(t, i):
        L, R = pairs[(t + i) % len(pairs)]
        eL, eR = exs[t]
        kl, dl = eL.extract(L); kr, dr = eR.extract(R)
        tb = eL.tables()
        o.stereo(kl, dl, eL.pyramid(), kr, dr, eR.pyramid(), tb[0], tb[1], c['camera'])
    t0 = time.perf_counter()
    one_pair(0, 0)
    one = time.perf_counter() - t0
    per_thread = max(1, int(seconds / max(one, 1e-3)))
    done = [0] * threads

    def work(t):
        for i in range(per_thread):
            one_pair(t, i)
            done[t] += 1
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t0
    return sum(done) / dt, f'{sum(done)} stereo pairs 1241x376 / 2000 kp (Extract x2 + ComputeStereoMatches), one pair per thread at a time, {dt:.1f} s'


def cpu_knn_rate(o, threads, nq=8192, nt=1000000):
    q = synth.descriptors(1, nq); t = synth.descriptors(2, nt)
    o.knn2(q[:64], t[:1000], 50, 0.6, threads=1)
    t0 = time.perf_counter()
    o.knn2(q, t, 50, 0.6, threads=threads)
    dt = time.perf_counter() - t0
    return nq * nt / dt / 1e9, dt, f'{nq} queries x {nt} train rows, {threads} threads'


def run_reference(args, rank, world, emit):
    if rank != 0:
        return
    o, kind, native = load_cpu_reference()
    threads = os.cpu_count() or 1
    frames = make_frames(max(24, threads), 0)
    for _ in range(max(args.warmup, 1)):
        r0, _ = cpu_extract_rate(o, frames, threads, 1)
    per_thread = max(2, int(1.5 * r0 / threads))                                   # each step is about 1.5 s of CPU work
    t_all, n_all = 0.0, 0
    for _ in range(args.steps):
        r, dt = cpu_extract_rate(o, frames, threads, per_thread)
        t_all += dt; n_all += threads * per_thread
    fps = n_all / t_all
    gp, _, knn_sample = cpu_knn_rate(o, threads)
    sample = f'{threads * per_thread} frames per step ({per_thread} per thread), {args.steps} steps'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_all / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': 'C1: ORBextractor::Extract on synthetic 640x480 gray frames, 1000 kp, scale 1.2, 8 levels, FAST 20/7',
                   'library': os.path.relpath(o.path, ROOT), 'flags': '-O3 -march=x86-64-v3' if native else '-O2 -ffp-contract=off',
                   'note': 'reference TUs compiled where they lie; OpenCV primitives are this repo\'s scalar restatements (no OpenCV on the box)'},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'knn': {'value': gp, 'unit': 'Gpairs/s', 'sample': knn_sample, 'cores': threads},
        'gpu_launches': 0,
    }
    emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# this repository's arm
# ---------------------------------------------------------------------------------------------------------------------
def _gpu_cpu_mask(pynvml, index, words):
    try:
        return tuple(int(w) for w in pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index), words))
    except Exception:
        return None


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: keep the process (and therefore its pinned staging buffers, first-touch) on the CPU cores NVML reports as
    local to that GPU, so that host<->device copies of different ranks do not cross the socket interconnect. Returns a note for `config`."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = tuple(int(w) for w in pynvml.nvmlDeviceGetCpuAffinity(h, words))
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        cpus = sorted(c for c in cpus if c in os.sched_getaffinity(0))
        world = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', 1)))
        if cpus:
            # ranks whose GPUs share a socket get DISJOINT slices of its cores: copy threads and the CUDA driver threads of different ranks
            # must not migrate over each other (pinned staging is first-touched after this call, so it lands on the same cores' memory)
            same = [r for r in range(world) if _gpu_cpu_mask(pynvml, r, words) == mask] if world > 1 else [local_rank]
            if len(same) > 1 and local_rank in same and len(cpus) >= 2 * len(same):
                per = len(cpus) // len(same)
                k = same.index(local_rank)
                cpus = cpus[k * per:(k + 1) * per]
            os.sched_setaffinity(0, cpus)
            return f'{len(cpus)} cores local to GPU {local_rank} ({cpus[0]}-{cpus[-1]}), disjoint from the other ranks on the socket'
    except Exception as e:     # affinity is an optimisation, never load-bearing
        return f'unavailable ({type(e).__name__})'
    return 'unavailable'


def run_b200(args, rank, world, local_rank, emit):
    affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else 'not set (single process: the CPU baseline uses every host core)'
    import torch
    import torch.distributed as dist
    from orb_slam2_refactored_b200 import api

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    api.lib()     # fails loudly if the CUDA library is missing
    hbm_peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    B = args.batch
    NB = 4                                     # rotating input batches: 4 x B x 300 KiB > 126 MB L2 for B >= 128
    host = make_frames(B * 2, 1000 * rank)     # two distinct pinned host batches for the end-to-end leg
    pinned = [torch.from_numpy(host[i * B:(i + 1) * B].copy()).pin_memory() for i in range(2)]
    d_batches = []
    for i in range(NB):
        src = pinned[i % 2]
        d = src.to(dev)
        if i >= 2:
            d = torch.roll(d, shifts=53 * i, dims=1).contiguous()
        d_batches.append(d)

    ex = api.ORBextractor(nfeatures=NFEATURES, device=local_rank)
    cap = None
    stream = torch.cuda.ExternalStream(ex.stream(), device=dev)
    outs = None

    def step(i):
        nonlocal outs
        if outs is None:
            outs = ex.extract_batch_device(d_batches[i % NB])
        else:
            ex.extract_batch_device(d_batches[i % NB], *outs)

    # ---- device-resident throughput. Two timed regions of K steps each, both bracketed by barrier + synchronize and timed by CUDA events
    # on the extractor's stream (the second lane is forked from and joined back into it):
    #   A  the call as a caller gets it: a batch runs as two half-batches on two streams with staggered stage orders -> `value`
    #   B  the same steps with the library's per-stage events on: one launch per stage on one stream, so that a stage's event pair brackets
    #      exactly its kernels -> the per-kernel launch durations of `roofline` (stage events around concurrent lanes would time nothing)
    barrier()      # inputs were produced on torch's stream; the extractor runs on its own
    for i in range(args.warmup):
        step(i)
    ex.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    frames_total = world * B * args.steps
    fps = frames_total / (ms * 1e-3)
    # region B
    ex.enable_stage_timing(True)
    for i in range(min(args.warmup, 3)):
        step(i)
    ex.synchronize()
    ex.stage_times()
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    u1.record(stream)
    u1.synchronize()
    t1 = time.time()
    barrier()
    ms_inst = max_over_ranks(u0.elapsed_time(u1))
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    stage_ms, calls = ex.stage_times()
    ex.enable_stage_timing(False)
    fps_inst = frames_total / (ms_inst * 1e-3)

    # keypoint counts of the last step, for the algorithmic byte count
    n_last = outs[2].cpu().numpy()
    n_mean = float(n_last.mean())
    sizes = ex.level_sizes()
    S = sum(w * h for w, h in sizes); P0 = sizes[0][0] * sizes[0][1]; P7 = sizes[-1][0] * sizes[-1][1]
    alg = {'pyramid': 2 * S - P0 - P7, 'fast': S, 'blur': 2 * S, 'describe': 1321 * n_mean, 'quadtree': 0.0}
    b_alg = 5 * S - P0 - P7 + 1321 * n_mean
    per_step = {k: v / max(calls, 1) for k, v in stage_ms.items()}
    dominant = max(per_step, key=per_step.get)
    # FAST is two kernels (dense bound pass over the level tiles, then one warp per cell); the blur of all levels is one launch
    launches_per_stage = {'pyramid': len(sizes) - 1, 'fast': 2, 'quadtree': 1, 'blur': 1, 'describe': 1}
    # roofline of the dominant stage, per launch: its algorithmic bytes for the whole batch / its device time
    dom_bytes_per_step = alg[dominant] * B
    dom_gbs = dom_bytes_per_step / (per_step[dominant] * 1e-3) / 1e9 if per_step[dominant] > 0 else 0.0
    traffic = None
    pipes = None
    tj = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if os.path.exists(tj):
        t = json.load(open(tj)).get(dominant)
        if t:   # dram__bytes_read + write of the stage's kernels in one captured step (ncu --set full, profiles/), scaled to this batch
            traffic = t['dram_bytes_per_step'] * B / float(t.get('frames_per_step') or 256)
            pipes = t.get('pipes')
    stage_kernels = {'pyramid': 'k_pyramid_strip x7', 'fast': 'k_level_strip<FAST> + k_fast_cells2 (+ k_fast_cells2_overflow: empty list on these frames)', 'quadtree': 'k_quadtree', 'blur': 'k_level_strip<BLUR>',
                     'describe': 'k_orient_describe2'}
    roofline = {'bound': 'hbm', 'kernel': dominant, 'kernels': stage_kernels[dominant], 'achieved': dom_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': dom_gbs / hbm_peak, 'traffic': traffic, 'peak_source': peak_src,
                'launches_per_step': launches_per_stage[dominant],
                'algorithmic_bytes_per_launch': dom_bytes_per_step / launches_per_stage[dominant],
                'avg_launch_ms': per_step[dominant] / launches_per_stage[dominant],
                'ncu_pipes': pipes,      # committed ncu capture (profiles/r02_*.txt), not this run: issue-slot / ALU / FMA / LSU utilisation of the stage's kernels
                'stages_ms_per_step': per_step,
                'stages_gbs': {k: (alg[k] * B / (per_step[k] * 1e-3) / 1e9 if per_step[k] > 0 else None) for k in per_step},
                'timed_region': 'B: K steps with per-stage events, one launch per stage on one stream (kernel durations); `value` and `frame` are region A',
                'frame': {'algorithmic_bytes_per_frame': b_alg, 'achieved': b_alg * fps / world / 1e9, 'frac': b_alg * fps / world / 1e9 / hbm_peak,
                          'frac_region_b': b_alg * fps_inst / world / 1e9 / hbm_peak}}

    # ---- end to end through the public host-buffer API: pinned H2D of the frames, D2H of keypoints + descriptors.
    # Two extractor instances on two host threads, each calling the synchronous orbx_extract_batch on its own batches — the
    # way the reference itself runs its two extractors (src/System.cc:449-452); their copies and kernels overlap on the GPU.
    import ctypes as C
    NH = args.e2e_handles
    exs = [ex] + [api.ORBextractor(nfeatures=NFEATURES, device=local_rank) for _ in range(NH - 1)]
    kcap = ex.max_keypoints()
    bufs = [(torch.empty((B, kcap, 28), dtype=torch.uint8).pin_memory().numpy(), torch.empty((B, kcap, 32), dtype=torch.uint8).pin_memory().numpy(),
             np.zeros(B, np.int32)) for _ in range(NH)]

    def e2e_step(hi, i):
        a = pinned[(hi + i) % 2].numpy()
        k, d, n = bufs[hi]
        api._check(api.lib().orbx_extract_batch(exs[hi]._h, C.c_void_p(a.ctypes.data), B, W, H, W, W * H, C.c_void_p(k.ctypes.data),
                                                C.c_void_p(d.ctypes.data), kcap, C.c_void_p(n.ctypes.data)))
    for hi in range(NH):
        for i in range(max(1, min(args.warmup, 2))):
            e2e_step(hi, i)
    e2e_steps = max(2, min(args.steps, 10))

    def e2e_worker(hi):
        for i in range(e2e_steps):
            e2e_step(hi, i)
    barrier()
    t_e2e = time.perf_counter()
    workers = [threading.Thread(target=e2e_worker, args=(hi,)) for hi in range(NH)]
    for t in workers: t.start()
    for t in workers: t.join()
    torch.cuda.synchronize(dev)
    ms_e2e = max_over_ranks((time.perf_counter() - t_e2e) * 1e3)      # the API is synchronous: wall clock around the calls is the end-to-end time
    barrier()
    e2e_fps = world * NH * B * e2e_steps / (ms_e2e * 1e-3)
    n_h = bufs[0][2]
    h2d = B * W * H
    d2h = B * kcap * 60 + 4 * B                       # the call downloads whole capacity rows (a strided copy per chunk), not n[f] rows
    # what the platform allows for exactly these transfers: the same bytes per step as plain pinned cudaMemcpyAsync on two streams (one per
    # direction), no kernels, all ranks at once — the ceiling of ANY end-to-end number on this host
    cs_in = [torch.cuda.Stream(dev) for _ in range(NH)]; cs_out = [torch.cuda.Stream(dev) for _ in range(NH)]
    d_in = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(NH)]
    d_out = torch.empty((B, kcap, 60), dtype=torch.uint8, device=dev)
    h_out = [torch.empty((B, kcap, 60), dtype=torch.uint8).pin_memory() for _ in range(NH)]

    def copy_step(i):
        for hi in range(NH):          # as many transfers per direction in flight as the end-to-end leg has extractor instances
            with torch.cuda.stream(cs_in[hi]):
                d_in[hi].copy_(pinned[(i + hi) % 2], non_blocking=True)
            with torch.cuda.stream(cs_out[hi]):
                h_out[hi].copy_(d_out, non_blocking=True)
    for i in range(2):
        copy_step(i)
    torch.cuda.synchronize(dev)
    barrier()
    t_c = time.perf_counter()
    csteps = e2e_steps
    for i in range(csteps):
        copy_step(i)
    torch.cuda.synchronize(dev)
    ms_copy = max_over_ranks((time.perf_counter() - t_c) * 1e3)
    barrier()
    copy_fps = world * NH * B * csteps / (ms_copy * 1e-3)
    copy_ceiling = {'value': copy_fps, 'unit': 'frames/s',
                    'gb_per_s_per_direction': {'h2d': world * NH * h2d * csteps / ms_copy / 1e6, 'd2h': world * NH * d2h * csteps / ms_copy / 1e6},
                    'note': f'pinned cudaMemcpyAsync of the same H2D and D2H bytes per step, {NH} streams per direction, no kernels, all ranks concurrently'}
    del d_in, d_out, h_out

    # ---- kNN: per-GPU share of configs[4]
    knn = None
    launches_knn = 0
    if not args.skip_knn:
        nq, nt = args.knn_queries, args.knn_train_per_gpu
        g = torch.Generator(device=dev); g.manual_seed(1234)
        dq = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=g)     # queries replicated on every rank
        nplant = max(1, nq // (8 * world))

        def make_shard(r):
            """Rank r's train shard: uniform random rows, plus nplant rows that are a query with ~16 of its 256 bits flipped (so that the
            TH_LOW / ratio test accepts matches) and, for every 5th of them, an exact copy at a second row (lowest index must win)."""
            gg = torch.Generator(device=dev); gg.manual_seed(99 + r)
            t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device=dev, generator=gg)
            k = torch.arange(nplant, device=dev)
            qi = (k * 131 + r * 7919) % nq
            tj = (k * 977 + 13) % nt
            flips = torch.randint(0, 256, (nplant, 32), dtype=torch.uint8, device=dev, generator=gg)
            for _ in range(3):
                flips &= torch.randint(0, 256, (nplant, 32), dtype=torch.uint8, device=dev, generator=gg)
            t[tj] = dq[qi] ^ flips
            dup = k[::5]
            t[(tj[dup] + nt // 2) % nt] = t[tj[dup]]
            return t
        dt = make_shard(rank)                                                                # this rank's train shard
        m = api.ORBmatcher(0.6, device=local_rank)
        gathered = torch.empty((world, nq), dtype=torch.int64, device=dev) if world > 1 else None
        part = torch.empty(nq, dtype=torch.int64, device=dev)

        def knn_step():
            if world == 1:
                return m.knn2_device(dq, dt)
            api.knn2_partial_device(dq, dt, rank * nt, part)
            dist.all_gather_into_tensor(gathered.view(-1), part)
            return api.knn2_merge_device(gathered, world, nq, 50, 0.6)
        knn_step()
        barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(args.knn_steps):
            res = knn_step()
        k1.record()
        k1.synchronize()
        barrier()
        ms_knn = max_over_ranks(k0.elapsed_time(k1)) / args.knn_steps
        gpairs = nq * (nt * world) / (ms_knn * 1e-3) / 1e9
        popc_peak = api.measure_popc_peak(local_rank)
        # check (every N): the first `nchk` queries of the timed result against ONE GPU scanning every shard in rank order — rank 0 rebuilds
        # the other ranks' shards from their seeds, scans each with the right index base and merges; all four outputs must be identical
        import hashlib
        nchk = min(nq, 32768)
        check = None
        if rank == 0:
            parts = torch.empty((world, nchk), dtype=torch.int64, device=dev)
            for r in range(world):
                sh = dt if r == 0 else make_shard(r)
                api.knn2_partial_device(dq[:nchk], sh, r * nt, parts[r])
                torch.cuda.synchronize(dev)
                del sh
            ref = api.knn2_merge_device(parts, world, nchk, 50, 0.6)
            torch.cuda.synchronize(dev)
            same = all(bool(torch.equal(a[:nchk], b)) for a, b in zip(res, ref))
            dig = hashlib.sha256(b''.join(a[:nchk].cpu().numpy().tobytes() for a in res)).hexdigest()[:16]
            check = {'result': 'ok' if same else 'MISMATCH', 'queries_checked': nchk, 'sha256_16': dig,
                     'against': f'single-GPU scan of all {world} shard(s) in rank order (orbx_knn2_partial_device per shard + orbx_knn2_merge_device)'}
            if not same:
                raise RuntimeError('kNN: the sharded result differs from the single-GPU scan of the same data')
        barrier()
        knn = {'value': gpairs, 'unit': 'Gpairs/s', 'ms_per_step': ms_knn, 'steps': args.knn_steps,
               'check': check['result'] if check else None, 'check_detail': check,
               'workload': f'C5 share: {nq} queries x {nt} train rows per GPU ({nt * world} total), uniform random 256-bit descriptors with '
                           f'{nplant} planted near-duplicates per shard',
               'roofline': {'bound': 'popc', 'achieved': 8 * gpairs / world * 1e9, 'peak': popc_peak, 'unit': 'POPC.32/s',
                            'frac': 8 * gpairs / world * 1e9 / popc_peak,
                            'peak_source': 'orbx_measure_popc_peak (register-operand POPC loop on all SMs, same run)'},
               'accepted_matches': int((res[3] >= 0).sum().item())}
        launches_knn = args.knn_steps * (2 if world == 1 else 2)
        # end to end on a bounded sample through the host-buffer API (uploads queries + train, downloads results)
        if world == 1:
            sq, st = 65536, min(nt, 1 << 20)
            hq = dq[:sq].cpu().numpy(); ht = dt[:st].cpu().numpy()
            m.knn2(hq[:1024], ht[:4096])
            t0k = time.perf_counter()
            m.knn2(hq, ht)
            dtk = time.perf_counter() - t0k
            knn['e2e'] = {'value': sq * st / dtk / 1e9, 'unit': 'Gpairs/s', 'h2d_bytes_per_step': 32 * (sq + st), 'd2h_bytes_per_step': 12 * sq,
                          'sample': f'{sq} x {st} through orbx_knn2 (pageable host buffers, allocation included)'}

    # ---- stereo: BASELINE.json configs[1] (KITTI shape 1241x376, 2000 kp, left + right Extract + ComputeStereoMatches), resident
    stereo = None
    if not args.skip_stereo:
        c2 = synth.CONFIGS['C2']
        SB = args.stereo_pairs
        base = [synth.stereo_pair(500 + 100 * rank + s, c2['w'], c2['h']) for s in range(8)]
        Ls = np.stack([base[i % 8][0] if (i // 8) % 2 == 0 else base[i % 8][0][::-1] for i in range(SB)])
        Rs = np.stack([base[i % 8][1] if (i // 8) % 2 == 0 else base[i % 8][1][::-1] for i in range(SB)])
        dL = torch.from_numpy(np.ascontiguousarray(Ls)).to(dev); dR = torch.from_numpy(np.ascontiguousarray(Rs)).to(dev)
        eL = api.ORBextractor(nfeatures=c2['nfeatures'], device=local_rank); eR = api.ORBextractor(nfeatures=c2['nfeatures'], device=local_rank)
        oL = eL.extract_batch_device(dL); oR = eR.extract_batch_device(dR)
        capS = oL[0].shape[1]
        d_ur = torch.empty((SB, capS), dtype=torch.float32, device=dev); d_dp = torch.empty((SB, capS), dtype=torch.float32, device=dev)
        cam = api._Camera(*[float(v) for v in c2['camera']])

        def stereo_step():
            eL.extract_batch_device(dL, *oL); eR.extract_batch_device(dR, *oR)
            api._check(api.lib().orbx_stereo_match_device(eL._h, eR._h, C.byref(cam), C.c_void_p(d_ur.data_ptr()), C.c_void_p(d_dp.data_ptr())))
        for _ in range(3):
            stereo_step()
        eL.synchronize(); eR.synchronize()
        barrier()
        sL = torch.cuda.ExternalStream(eL.stream(), device=dev)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ssteps = max(2, min(args.steps, 10))
        torch.cuda.synchronize(dev)
        s0.record(sL)
        for _ in range(ssteps):
            stereo_step()
        s1.record(sL)           # the match runs on the left extractor's stream after waiting for the right one
        s1.synchronize(); eR.synchronize()
        barrier()
        ms_st = max_over_ranks(s0.elapsed_time(s1)) / ssteps
        matched = int((d_dp[:, :] > 0).sum().item())
        szS = eL.level_sizes()
        S2 = sum(w_ * h_ for w_, h_ in szS)
        nS = float(oL[2].float().mean().item())
        b_alg2 = 2 * (5 * S2 - szS[0][0] * szS[0][1] - szS[-1][0] * szS[-1][1] + 1321 * nS)       # two extractions per pair; the match itself adds < 1 %
        pairs_s = world * SB / (ms_st * 1e-3)
        stereo = {'value': pairs_s, 'unit': 'stereo pairs/s', 'ms_per_step': ms_st, 'pairs_per_step_per_gpu': SB,
                  'workload': 'C2: 1241x376 stereo, 2000 kp per image, Extract left + right and ComputeStereoMatches, device-resident',
                  'matched_per_pair': matched / SB, 'keypoints_per_image': nS,
                  'roofline': {'bound': 'hbm', 'frame': {'algorithmic_bytes_per_pair': b_alg2, 'achieved': b_alg2 * pairs_s / world / 1e9, 'unit': 'GB/s',
                                                         'peak': hbm_peak, 'frac': b_alg2 * pairs_s / world / 1e9 / hbm_peak}}}
        if rank == 0 and world == 1 and not args.skip_cpu:
            try:
                o, kind, native = load_cpu_reference()
                r_, smp = cpu_stereo_rate(o, os.cpu_count() or 1)
                stereo['cpu_baseline'] = {'value': r_, 'unit': 'stereo pairs/s', 'cores': os.cpu_count() or 1, 'kind': kind, 'sample': smp}
            except Exception as e:
                stereo['cpu_baseline'] = {'unavailable': str(e)}
        del dL, dR, eL, eR, oL, oR

    # ---- BASELINE.json configs[2] and configs[3]: EuRoC-shape 752x480 / 1200 kp and 4K 3840x2160 / 8000 kp, device-resident extraction on
    #      every rank's own frames (frames sharded, no collective), frame-level roofline, CPU reference beside it at N = 1
    configs = None
    if not args.skip_configs:
        configs = {}
        for name, cb, csteps_ in (('C3', 512, 6), ('C4', args.c4_batch, 4)):
            c = synth.CONFIGS[name]
            base = np.stack([synth.image(2000 + 17 * rank + s_, c['w'], c['h']) for s_ in range(4)])
            d1 = torch.from_numpy(base).to(dev).repeat((cb + 3) // 4, 1, 1)[:cb].contiguous()
            d2 = torch.flip(d1, dims=[1]).contiguous()
            ec = api.ORBextractor(nfeatures=c['nfeatures'], device=local_rank)
            oc = ec.extract_batch_device(d1)
            for i in range(3):
                ec.extract_batch_device(d2 if i & 1 else d1, *oc)
            ec.synchronize()
            barrier()
            sc = torch.cuda.ExternalStream(ec.stream(), device=dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(sc)
            for i in range(csteps_):
                ec.extract_batch_device(d2 if i & 1 else d1, *oc)
            a1.record(sc)
            a1.synchronize()
            barrier()
            ms_c = max_over_ranks(a0.elapsed_time(a1)) / csteps_
            szc = ec.level_sizes()
            Sc = sum(w_ * h_ for w_, h_ in szc)
            nc = float(oc[2].float().mean().item())
            b_c = 5 * Sc - szc[0][0] * szc[0][1] - szc[-1][0] * szc[-1][1] + 1321 * nc
            fps_c = world * cb / (ms_c * 1e-3)
            blk = {'workload': f'{name}: {c["w"]}x{c["h"]}, {c["nfeatures"]} kp, 8 levels, device-resident, {cb} frames per step per GPU '
                               f'(two alternating batches of {cb * c["w"] * c["h"] / 1e6:.0f} MB + slabs: larger than L2)',
                   'value': fps_c, 'unit': 'frames/s', 'ms_per_step': ms_c, 'steps': csteps_, 'keypoints_per_frame': nc,
                   'roofline': {'bound': 'hbm', 'frame': {'algorithmic_bytes_per_frame': b_c, 'achieved': b_c * fps_c / world / 1e9, 'unit': 'GB/s',
                                                          'peak': hbm_peak, 'frac': b_c * fps_c / world / 1e9 / hbm_peak}}}
            if name == 'C3':
                blk['sequence_10k_frames_s'] = 10000.0 / fps_c       # configs[2]: a 10k-frame sequence sharded over the GPUs of this run
            if rank == 0 and world == 1 and not args.skip_cpu:
                try:
                    o, kind, native = load_cpu_reference()
                    r_, smp = cpu_config_rate(o, name, os.cpu_count() or 1, 3.0)
                    blk['cpu_baseline'] = {'value': r_, 'unit': 'frames/s', 'cores': os.cpu_count() or 1, 'kind': kind, 'sample': smp}
                except Exception as e:
                    blk['cpu_baseline'] = {'unavailable': str(e)}
            configs[name] = blk
            del d1, d2, ec, oc
        if stereo is not None:
            configs['C2'] = {k_: stereo[k_] for k_ in ('workload', 'value', 'unit', 'ms_per_step', 'roofline', 'cpu_baseline') if k_ in stereo}

    # ---- guided matchers: one tracking search per call through the C ABI, host buffers in and out. Latency-bound
    #      (one CTA per search), so it is reported as microseconds per call next to the reference text on one host core.
    guided = None
    if rank == 0 and not args.skip_guided:
        fr = synth.frame(1, n=1000)
        gf = api.Frame(fr['kps_un'], fr['desc'], fr['scale_factors'], fr['bounds'], fr['uright'], device=local_rank)
        gpts, gdesc = synth.local_map_points(1, fr, npts=1000)
        gcp, glp, glpts, gldesc = synth.last_frame_points(1, fr, synth.KITTI_CAMERA, npts=1000)
        gm = api.ORBmatcher(0.8, True, device=local_rank)

        def g_local():
            gf.mappoints[:] = -1
            return gm.SearchByProjection(gf, gpts, gdesc, 3.0)

        def g_last():
            gf.mappoints[:] = -1
            return gm.SearchByProjectionLastFrame(gf, synth.KITTI_CAMERA, gcp, glp, glpts, gldesc, 7.0, False)

        def per_call(fn, reps=200):
            for _ in range(10):
                fn()
            t0g = time.perf_counter()
            for _ in range(reps):
                fn()
            return (time.perf_counter() - t0g) / reps * 1e6
        guided = {'workload': '1000 keypoints (640x480), 1000 map points; SearchByProjection for local-map and motion-model tracking',
                  'unit': 'us per call', 'local_map': per_call(g_local), 'local_map_kernel': gf.last_stats()[1] * 1e3,
                  'last_frame': per_call(g_last), 'last_frame_kernel': gf.last_stats()[1] * 1e3, 'rounds': gf.last_rounds(),
                  'frame_assign': per_call(lambda: gf.assign(fr['kps_un'], fr['desc'], fr['scale_factors'], fr['bounds'], fr['uright'])),
                  'h2d_bytes_per_call': int(gpts.nbytes + gdesc.nbytes + 4 * len(fr['kps_un'])), 'd2h_bytes_per_call': int(4 * len(fr['kps_un']) + 16)}
        # the matchers of local mapping / loop closing (Fuse, SearchBySim3, SearchForTriangulation): search on the device, per call
        _, inv_sig = synth.sigma_tables(fr['scale_factors'])
        Sf, fpts, fdesc = synth.sim3_points(21, fr, synth.KITTI_CAMERA, npts=1000, scale=1.0)
        fstate = synth.fuse_map(1, fr, len(fpts))
        lsf = synth.log_scale_factor()
        sp = synth.sim3_pair(1, n=1000)
        sp['lsf'] = lsf
        sf1 = api.Frame(sp['f1']['kps_un'], sp['f1']['desc'], sp['f1']['scale_factors'], sp['f1']['bounds'], sp['f1']['uright'], device=local_rank)
        sf2 = api.Frame(sp['f2']['kps_un'], sp['f2']['desc'], sp['f2']['scale_factors'], sp['f2']['bounds'], sp['f2']['uright'], device=local_rank)
        tp = synth.triangulation_pair(1, n=1000)
        tf1 = api.Frame(tp['f1']['kps_un'], tp['f1']['desc'], tp['f1']['scale_factors'], tp['f1']['bounds'], tp['f1']['uright'], device=local_rank)
        tf2 = api.Frame(tp['f2']['kps_un'], tp['f2']['desc'], tp['f2']['scale_factors'], tp['f2']['bounds'], tp['f2']['uright'], device=local_rank)
        gm2 = api.ORBmatcher(0.6, True, device=local_rank)
        guided['mapping'] = {
            'workload': '1000 keypoints per key frame, 1000 map points; the search half of Fuse, SearchBySim3 (both directions + agreement) and '
                        'SearchForTriangulation, host buffers in and out',
            'unit': 'us per call',
            'fuse': per_call(lambda: gm2.FuseSearch(gf, synth.KITTI_CAMERA, (Sf[0], Sf[1]), lsf, inv_sig, fpts, fdesc, 3.0)),
            'search_by_sim3': per_call(lambda: gm2.SearchBySim3(sf1, sp['cam'], sp['pose1'], lsf, sf2, sp['cam'], sp['pose2'], lsf, sp['S12'], 7.5,
                                                                sp['pts1'], sp['desc1'], sp['pts2'], sp['desc2'])),
            'search_for_triangulation': per_call(lambda: gm2.SearchForTriangulation(tf1, tp['fv1'], tp['has1'], tf2, tp['fv2'], tp['has2'], tp['F12'],
                                                                                    tp['ep2'], tp['sigma_sq2'], False))}
        if world == 1 and not args.skip_cpu:
            try:
                o, kind, native = load_cpu_reference()
                if kind == 'reference':
                    guided['mapping']['cpu_baseline'] = {
                        'fuse': o.time_fuse(fr, synth.KITTI_CAMERA, (Sf[0], Sf[1]), lsf, inv_sig, fpts, fdesc, 3.0, fstate, 50) * 1e6,
                        'search_by_sim3': o.time_search_by_sim3(sp, 7.5, 50) * 1e6,
                        'search_for_triangulation': o.time_search_for_triangulation(tp, False, True, 50) * 1e6,
                        'unit': 'us per call', 'cores': 1, 'kind': kind,
                        'sample': '50 calls each of the whole reference function (search + map mutation); key frames, grids and map points built outside '
                                  'the timed region'}
            except Exception as e:
                guided['mapping']['cpu_baseline'] = {'unavailable': str(e)}
        if world == 1 and not args.skip_cpu:
            try:
                o, kind, native = load_cpu_reference()
                mp0 = np.full(len(fr['kps_un']), -1, np.int32)
                guided['cpu_baseline'] = {
                    'local_map': o.time_search_local_map(fr, mp0, gpts, gdesc, 3.0, 0.8, 100) * 1e6,
                    'last_frame': o.time_search_last_frame(fr, synth.KITTI_CAMERA, gcp, glp, mp0, glpts, gldesc, 7.0, False, 0.9, True, 100) * 1e6,
                    'unit': 'us per call', 'cores': 1, 'kind': kind,
                    'sample': '100 calls each; Frame, grid and map points built outside the timed region'}
            except Exception as e:
                guided['cpu_baseline'] = {'unavailable': str(e)}

    # ---- rectification remap (Examples/Stereo/stereo_euroc.cc:100-101) at the EuRoC frame size, device-resident: an L2-bound
    #      kernel (8 B of table per pixel out of L2, 1 B gathered + 1 B written per pixel of DRAM traffic)
    remap = None
    if rank == 0 and not args.skip_guided:
        RB, RW, RH = 256, 752, 480
        rmx, rmy = synth.rectification_maps(8, RW, RH)
        rex = api.ORBextractor(nfeatures=1200, device=local_rank)
        rex.SetRectification(rmx, rmy, (RH, RW))
        raw_host = np.stack([synth.image(900 + s, RW, RH) for s in range(8)])
        d_raw = torch.from_numpy(np.ascontiguousarray(raw_host[np.arange(RB) % 8])).to(dev)
        rpitch = (RW + 127) // 128 * 128
        d_rect = [torch.empty((RB, RH, rpitch), dtype=torch.uint8, device=dev) for _ in range(3)]     # 3 x 98 MB outputs + 92 MB in: > L2 per rotation
        rs = torch.cuda.ExternalStream(rex.stream(), device=dev)

        def remap_step(i):
            api._check(api.lib().orbx_rectify_batch_device(rex._h, C.c_void_p(d_raw.data_ptr()), RB, RW, RW * RH, C.c_void_p(d_rect[i % 3].data_ptr()),
                                                           rpitch, rpitch * RH))
        for i in range(3):
            remap_step(i)
        rex.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rsteps = 20
        r0.record(rs)
        for i in range(rsteps):
            remap_step(i)
        r1.record(rs)
        r1.synchronize()
        ms_r = r0.elapsed_time(r1) / rsteps
        alg = RB * RW * RH * 2 + RW * RH * 8          # one byte gathered + one written per pixel, the 8-byte table entries once per launch
        remap = {'workload': f'cv::remap INTER_LINEAR of {RB} frames {RW}x{RH} with one rectification table, device-resident',
                 'value': RB / (ms_r * 1e-3), 'unit': 'frames/s', 'ms_per_step': ms_r,
                 'roofline': {'bound': 'hbm', 'kernel': 'k_remap_to_l0', 'achieved': alg / (ms_r * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                              'frac': alg / (ms_r * 1e-3) / 1e9 / hbm_peak, 'algorithmic_bytes_per_pixel': 2,
                              'l2_bytes_per_pixel': 10,
                              'note': 'L2-bound, not HBM-bound: every pixel reads its 8-byte table entry, but the 2.9 MB table is shared by all frames of '
                                      'a launch and stays in L2 (ncu r01: 1.59 B/px of DRAM traffic); the HBM fraction counts 2 B/px + the table once'}}
        del d_rect, d_raw

    # ---- bag-of-words transform (Frame::ComputeBoW): the reference's vocabulary shape (k 10, L 6, levelsup 4) with
    #      synthetic nodes; per call through the C ABI with host buffers, and a device-resident batch behind Extract
    bow = None
    if rank == 0 and not args.skip_guided:
        vt = synth.vocabulary(7, 10, 6)
        gv = api.ORBVocabulary(device=local_rank).create(10, 6, vt['parent'], vt['is_leaf'], vt['desc'], vt['weights'])
        bfe = synth.vocabulary_features(3, vt, 1000)
        for _ in range(10):
            gv.transform(bfe, 4)
        t0b = time.perf_counter()
        for _ in range(200):
            gv.transform(bfe, 4)
        us_call = (time.perf_counter() - t0b) / 200 * 1e6
        BF, BCAP = 256, 1024
        d_bdesc = torch.from_numpy(np.ascontiguousarray(np.stack([synth.vocabulary_features(10 + f_ % 8, vt, BCAP) for f_ in range(8)])[np.arange(BF) % 8])).to(dev)
        d_bn = torch.full((BF,), 1000, dtype=torch.int32, device=dev)
        bs = torch.cuda.ExternalStream(ex.stream(), device=dev)
        for _ in range(3):
            gv.transform_batch_device(d_bdesc, d_bn, BF, BCAP, 4, stream=ex.stream())
        ex.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(bs)
        for _ in range(10):
            gv.transform_batch_device(d_bdesc, d_bn, BF, BCAP, 4, stream=ex.stream())
        b1.record(bs)
        b1.synchronize()
        ms_b = b0.elapsed_time(b1) / 10
        bow = {'workload': 'DBoW2 transform, vocabulary k 10 / L 6 (1.1 M nodes, synthetic), 1000 descriptors per frame, levelsup 4',
               'us_per_call': us_call, 'api': 'orbx_bow_transform (host descriptors in, BowVector + FeatureVector out)',
               'batch_frames_per_s': BF / (ms_b * 1e-3), 'batch_ms_per_step': ms_b,
               'batch_api': f'orbx_bow_transform_batch_device, {BF} frames per call, device-resident (2 launches)',
               'h2d_bytes_per_call': 32000, 'd2h_bytes_per_call': int(1000 * (4 + 8 + 4 + 4) + 16)}
        if world == 1 and not args.skip_cpu:
            try:
                from oracle import bindings as ob
                op = ob.Oracle('port', native=True)
                ovv = op.vocabulary(arrays=vt)
                bow['cpu_baseline'] = {'us_per_call': ovv.time_transform(bfe, 4, 50) * 1e6, 'cores': 1, 'kind': 'port',
                                       'sample': '50 transforms of the same 1000 descriptors, oracle/bow_oracle.cc (-O3 -march=x86-64-v3)'}
            except Exception as e:
                bow['cpu_baseline'] = {'unavailable': str(e)}
        del d_bdesc

    # ---- single-frame latency: what SystemImpl::Track* sees per image (one Extract call, host frame in, keypoints + descriptors out)
    latency = None
    if rank == 0 and not args.skip_guided:
        lex = api.ORBextractor(nfeatures=1000, device=local_rank)
        one = torch.from_numpy(host[:1].copy()).pin_memory().numpy()
        for _ in range(10):
            lex.ExtractBatch(one)
        t0l = time.perf_counter()
        for _ in range(100):
            lex.ExtractBatch(one)
        latency = {'workload': 'C1, one frame per call through ORBextractor.ExtractBatch (pinned host frame in, host keypoints + descriptors out)',
                   'ms_per_call': (time.perf_counter() - t0l) / 100 * 1e3}
        # the same call at the C ABI with the caller's arrays allocated once, as a C++ host makes it (no numpy allocation / slicing per call)
        lcap = lex.max_keypoints()
        lk = np.empty((1, lcap, 28), np.uint8); ld = np.empty((1, lcap, 32), np.uint8); ln = np.zeros(1, np.int32)
        largs = (lex._h, C.c_void_p(one.ctypes.data), 1, W, H, W, W * H, C.c_void_p(lk.ctypes.data), C.c_void_p(ld.ctypes.data), lcap, C.c_void_p(ln.ctypes.data))
        fn = api.lib().orbx_extract_batch
        for _ in range(10):
            api._check(fn(*largs))
        t0l = time.perf_counter()
        for _ in range(200):
            fn(*largs)
        latency['ms_per_call_c_abi'] = (time.perf_counter() - t0l) / 200 * 1e3
        latency['note'] = ('ms_per_call: through the Python mirror (allocates and slices the result arrays per call); ms_per_call_c_abi: orbx_extract_batch with '
                           'preallocated caller arrays. The call runs the levels as three groups (level 0 / 1-2 / 3-7) on parallel streams, replayed as one CUDA graph')

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own code on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        try:
            o, kind, native = load_cpu_reference()
            threads = os.cpu_count() or 1
            r0, _ = cpu_extract_rate(o, host[:48], threads, 2)                    # calibration
            per_thread = max(4, int(12.0 * r0 / threads))                          # about 12 s of CPU work
            r, dt_cpu = cpu_extract_rate(o, host[:48], threads, per_thread)
            cpu = {'value': r, 'unit': 'frames/s', 'cores': threads, 'kind': kind,
                   'sample': f'{threads * per_thread} frames of the same workload, one frame per thread at a time, {dt_cpu:.1f} s, '
                             f'{os.path.relpath(o.path, ROOT)} ({"-O3 -march=x86-64-v3" if native else "-O2"}; OpenCV primitives restated scalar)'}
            if knn is not None:
                gp, dtk, smp = cpu_knn_rate(o, threads)
                knn['cpu_baseline'] = {'value': gp, 'unit': 'Gpairs/s', 'cores': threads, 'kind': kind, 'sample': smp}
        except Exception as e:   # the baseline is reported, never load-bearing
            cpu = {'value': None, 'unit': 'frames/s', 'cores': 0, 'kind': 'port', 'sample': f'unavailable: {e}'}

    if rank == 0:
        # timed device-resident steps (+ the cell kernel's overflow launch, which finds its list empty on these frames) + timed kNN steps
        # region A runs every stage once per lane (two lanes from 64 frames per call); region B once per step
        lanes = 2 if B >= 64 else 1
        launches = args.steps * (lanes + 1) * (sum(launches_per_stage.values()) + 1) + launches_knn
        line = {
            'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
            'data': 'synthetic',
            'config': {'workload': 'C1: ORBextractor::Extract on synthetic 640x480 gray frames, 1000 kp, scale 1.2, 8 levels, FAST 20/7',
                       'frames_per_step_per_gpu': B, 'keypoints_per_frame': n_mean, 'cpu_affinity': affinity, 'parallelism': f'frames sharded over {world} GPU(s), no data-path collective',
                       'l2': f'inputs larger than L2: {NB} rotating device batches of {B} frames ({NB * B * W * H / 1e6:.0f} MB) + {B * 2.1:.0f} MB of pyramid/blur slabs per step'},
            'e2e': {'value': e2e_fps, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                    'copy_ceiling': copy_ceiling, 'frac_of_copy_ceiling': e2e_fps / copy_fps if copy_fps > 0 else None,
                    'handles': NH, 'api': f'orbx_extract_batch (pinned host frames in, keypoints + descriptors out), {NH} extractor instances on '
                                          f'{NH} host threads per GPU, one {B}-frame batch per call; wall clock around the synchronous calls'},
            'gpu_launches': launches,
            'instrumented': {'value': fps_inst, 'unit': 'frames/s', 'ms_per_step': ms_inst / args.steps,
                             'note': 'timed region B: the same K steps with the per-stage CUDA events on (one launch per stage on one stream instead of two '
                                     'half-batches with staggered stage orders on two streams); the roofline kernel durations come from here'},
            'clocks': clocks,
            'roofline': roofline,
            'cpu_baseline': cpu,
            'knn': knn,
            'configs': configs,
            'stereo': stereo,
            'guided': guided,
            'remap': remap,
            'latency': latency,
            'bow': bow,
        }
        emit(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=512, help='frames per step per GPU')
    ap.add_argument('--e2e-handles', type=int, default=2, help='extractor instances (host threads) of the end-to-end leg')
    ap.add_argument('--knn-steps', type=int, default=2)
    ap.add_argument('--knn-queries', type=int, default=1000000)
    ap.add_argument('--knn-train-per-gpu', type=int, default=1250000)
    ap.add_argument('--skip-knn', action='store_true')
    ap.add_argument('--skip-stereo', action='store_true')
    ap.add_argument('--skip-configs', action='store_true', help='skip the C3 / C4 blocks')
    ap.add_argument('--c4-batch', type=int, default=96, help='4K frames per step per GPU (the quadtree is one CTA per frame and level: small batches leave SMs idle)')
    ap.add_argument('--stereo-pairs', type=int, default=256, help='stereo pairs per step per GPU (C2 block)')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-guided', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local_rank = int(os.environ.get('LOCAL_RANK', 0))
    # stdout carries exactly one JSON line: anything libraries print meanwhile (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(line, flush=True)
        os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args, rank, world, emit)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_b200(args, rank, world, local_rank, emit)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
