"""Device-resident extraction throughput for the image shapes of BASELINE.json configs C1-C4 (synthetic frames)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
for name, B in (('C1', 512), ('C2', 256), ('C3', 512), ('C4', 32), ('C4', 128)):
    c = synth.CONFIGS[name]
    base = np.stack([synth.image(s, c['w'], c['h']) for s in range(4)])
    d = torch.from_numpy(base).cuda().repeat((B + 3) // 4, 1, 1)[:B].contiguous()
    d2 = torch.flip(d, dims=[1]).contiguous()
    ex = api.ORBextractor(nfeatures=c['nfeatures'])
    outs = ex.extract_batch_device(d)
    for i in range(3): ex.extract_batch_device(d if i & 1 else d2, *outs)
    ex.synchronize(); t = time.perf_counter()
    reps = 10
    for i in range(reps): ex.extract_batch_device(d if i & 1 else d2, *outs)
    ex.synchronize(); dt = (time.perf_counter() - t) / reps
    S = sum(w * h for w, h in ex.level_sizes()); P0 = c['w'] * c['h']; P7 = ex.level_sizes()[-1][0] * ex.level_sizes()[-1][1]
    n = float(outs[2].float().mean())
    balg = 5 * S - P0 - P7 + 1321 * n
    print(f"{name} {c['w']}x{c['h']} nf={c['nfeatures']} batch={B}: {B/dt:9.0f} frames/s  {dt*1e3/B*1e3:7.1f} us/frame  kp/frame={n:.0f}  B_alg={balg/1e6:.2f} MB  alg GB/s={balg*B/dt/1e9:.0f} ({balg*B/dt/1e9/6548.2*100:.1f}% of HBM copy peak)")
    os.environ['ORBX_LANES'] = '1'
    ex.enable_stage_timing(True); ex.stage_times()
    for i in range(4): ex.extract_batch_device(d if i & 1 else d2, *outs)
    ms, calls = ex.stage_times()
    ex.enable_stage_timing(False)
    print('    stage us/frame:', {k: round(v / calls / B * 1e3, 2) for k, v in ms.items()})
