"""Where a single-frame Extract call spends its time: GPU stage times (CUDA events) vs wall clock of the synchronous call."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
import torch
for (w, h, nf, name) in ((640, 480, 1000, 'C1'), (1241, 376, 2000, 'C2')):
    ex = api.ORBextractor(nfeatures=nf)
    img = synth.image(0, w, h)[None]
    pin = torch.from_numpy(img).pin_memory().numpy()
    for _ in range(10): ex.ExtractBatch(pin)
    n = 100
    t = time.perf_counter()
    for _ in range(n): ex.ExtractBatch(pin)
    wall = (time.perf_counter() - t) / n * 1e3
    ex.enable_stage_timing(True)
    for _ in range(3): ex.ExtractBatch(pin)
    st = ex.stage_times()
    print(name, f'wall {wall:.3f} ms per call;', 'stage ms of the last call:', {k: round(v, 4) for k, v in st.items()} if isinstance(st, dict) else st)
