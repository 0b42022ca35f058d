"""Single-frame / small-batch latency of the host-buffer API (what SystemImpl::Track* would see per frame)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
import torch
for (w, h, nf, name) in ((640, 480, 1000, 'C1'), (1241, 376, 2000, 'C2'), (3840, 2160, 8000, 'C4')):
    ex = api.ORBextractor(nfeatures=nf)
    for F in (1, 2, 8):
        if name == 'C4' and F > 2: continue
        imgs = np.stack([synth.image(s, w, h) for s in range(F)])
        pin = torch.from_numpy(imgs).pin_memory().numpy()
        for _ in range(5): ex.ExtractBatch(pin)
        n = 50 if name != 'C4' else 10
        t = time.perf_counter()
        for _ in range(n): k, d = ex.ExtractBatch(pin)
        dt = (time.perf_counter() - t) / n
        print(f'{name} {w}x{h} nf={nf} frames={F}: {dt*1e3:.3f} ms per call, {dt*1e3/F:.3f} ms per frame, {len(k[0])} kp')
