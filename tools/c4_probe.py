"""C4 (3840x2160, 8000 kp) device-resident extraction throughput and stage times; ORBX_QT_THREADS selects the quadtree CTA size."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
c = synth.CONFIGS['C4']; B = 32
base = np.stack([synth.image(s, c['w'], c['h']) for s in range(4)])
d = torch.from_numpy(base).cuda().repeat((B + 3) // 4, 1, 1)[:B].contiguous()
ex = api.ORBextractor(nfeatures=c['nfeatures'])
outs = ex.extract_batch_device(d)
for i in range(3): ex.extract_batch_device(d, *outs)
ex.synchronize(); t = time.perf_counter()
for i in range(8): ex.extract_batch_device(d, *outs)
ex.synchronize(); dt = (time.perf_counter() - t) / 8
ex.enable_stage_timing(True); ex.stage_times()
for i in range(4): ex.extract_batch_device(d, *outs)
ms, calls = ex.stage_times()
print(os.environ.get('ORBX_QT_THREADS', 'default (512)'), f'{B / dt:.0f} frames/s', {k: round(v / calls / B * 1e3, 2) for k, v in ms.items()})
