"""Device-resident extraction throughput and per-stage times for one BASELINE.json config (default C1, 512 frames).
usage: stage_probe.py [C1|C2|C3|C4] [batch] — tuning knobs come from the environment (ORBX_STRIP_TH, ORBX_FUSE, ORBX_PYR_TH, ORBX_BLUR_SIDE, ORBX_LANES)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
name = sys.argv[1] if len(sys.argv) > 1 else 'C1'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
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
knobs = {k: v for k, v in os.environ.items() if k.startswith('ORBX_')}
print(f"{name} batch={B} {knobs}: {B/dt:9.0f} frames/s  {dt*1e3:7.3f} ms/step  kp/frame={n:.0f}  alg GB/s={balg*B/dt/1e9:.0f} ({balg*B/dt/1e9/6548.2*100:.1f}% of HBM copy peak)")
os.environ['ORBX_LANES'] = '1'
ex.enable_stage_timing(True); ex.stage_times()
for i in range(6): ex.extract_batch_device(d if i & 1 else d2, *outs)
ms, calls = ex.stage_times()
ex.enable_stage_timing(False)
print('    stage ms/step:', {k: round(v / calls, 4) for k, v in ms.items()}, ' sum', round(sum(ms.values()) / calls, 4))
