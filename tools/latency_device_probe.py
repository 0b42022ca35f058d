"""Wall clock of a one-frame orbx_extract_batch_device call (frame and results stay on the device; launches issued call by call, single chain)."""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orb_slam2_refactored_b200 import api, synth
ex = api.ORBextractor(nfeatures=1000)
d = torch.from_numpy(synth.image(0, 640, 480)[None]).cuda()
outs = ex.extract_batch_device(d)
for _ in range(20): ex.extract_batch_device(d, *outs); ex.synchronize()
t = time.perf_counter()
for _ in range(200): ex.extract_batch_device(d, *outs); ex.synchronize()
print('device-resident one-frame call: %.4f ms' % ((time.perf_counter() - t) / 200 * 1e3))
