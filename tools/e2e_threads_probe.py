"""End-to-end host-buffer throughput with 1, 2, 3 extractor handles on as many host threads (like src/System.cc:449-452)."""
import os, sys, time, threading, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api
import bench
B = 256
host = bench.make_frames(B, 0)
for nh in (1, 2, 3):
    exs = [api.ORBextractor(nfeatures=1000) for _ in range(nh)]
    pins = [torch.from_numpy(np.roll(host, 13 * i, 1).copy()).pin_memory() for i in range(nh)]
    kcap = exs[0].max_keypoints()
    outs = [(torch.empty((B, kcap, 28), dtype=torch.uint8).pin_memory().numpy(), torch.empty((B, kcap, 32), dtype=torch.uint8).pin_memory().numpy(), np.zeros(B, np.int32)) for _ in range(nh)]
    def step(i):
        a = pins[i].numpy(); k, d, n = outs[i]
        api._check(api.lib().orbx_extract_batch(exs[i]._h, C.c_void_p(a.ctypes.data), B, 640, 480, 640, 640 * 480, C.c_void_p(k.ctypes.data), C.c_void_p(d.ctypes.data), kcap, C.c_void_p(n.ctypes.data)))
    for i in range(nh):
        for _ in range(3): step(i)
    reps = 12
    def work(i):
        for _ in range(reps): step(i)
    th = [threading.Thread(target=work, args=(i,)) for i in range(nh)]
    t = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt = time.perf_counter() - t
    print('handles', nh, '%.0f fps' % (nh * reps * B / dt))
