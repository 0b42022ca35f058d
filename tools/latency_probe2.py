"""Wall clock of one-frame Extract calls through the C ABI (pinned host frame in, host keypoints + descriptors out), checked against
a batch extraction of the same frame (the one-frame path uses 8-row tiles, two keypoints per warp and a side-stream blur)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
for (w, h, nf, name) in ((640, 480, 1000, 'C1'), (1241, 376, 2000, 'C2')):
    ex = api.ORBextractor(nfeatures=nf)
    imgs = [torch.from_numpy(synth.image(s, w, h)[None]).pin_memory().numpy() for s in range(4)]
    ref = api.ORBextractor(nfeatures=nf)
    for i in range(12): k, d = ex.ExtractBatch(imgs[i % 4])
    for i in range(4):
        k, d = ex.ExtractBatch(imgs[i]); k2, d2 = ref.ExtractBatch(np.concatenate([imgs[i]] * 20))      # 20 > ORBX_SMALL_BATCH: the throughput path
        assert k[0].tobytes() == k2[7].tobytes() and np.array_equal(d[0], d2[7]), 'the one-frame path differs from the batch path'
    n = 200
    t = time.perf_counter()
    for i in range(n): ex.ExtractBatch(imgs[i % 4])
    print(name, {k_: v for k_, v in os.environ.items() if k_.startswith('ORBX_')}, f'{(time.perf_counter() - t) / n * 1e3:.4f} ms per call')
