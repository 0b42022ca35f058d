"""Device-resident throughput with 1 vs 2 stream lanes (ORBX_LANES), batch 256."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api
import bench
B = 256
host = bench.make_frames(B, 0)
d = [torch.from_numpy(host).cuda(), torch.roll(torch.from_numpy(host).cuda(), 77, 1).contiguous()]
for lanes, chunk in ((1, 256), (2, 256), (1, 128), (1, 64), (2, 64), (1, 32), (2, 32), (1, 16)):
    os.environ['ORBX_LANES'] = str(lanes); os.environ['ORBX_DEV_CHUNK'] = str(chunk)
    ex = api.ORBextractor(nfeatures=1000)
    outs = ex.extract_batch_device(d[0])
    for i in range(3): ex.extract_batch_device(d[i & 1], *outs)
    ex.synchronize(); t = time.perf_counter()
    for i in range(20): ex.extract_batch_device(d[i & 1], *outs)
    ex.synchronize(); dt = (time.perf_counter() - t) / 20
    print('lanes', lanes, 'chunk', chunk, '%.3f ms/step %.0f fps' % (dt * 1e3, B / dt))
