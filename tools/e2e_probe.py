import sys, time, os, numpy as np, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api
import bench
B=256
host = bench.make_frames(B, 0)
pin = torch.from_numpy(host).pin_memory()
d = torch.empty_like(pin, device='cuda')
torch.cuda.synchronize()
for _ in range(3): d.copy_(pin, non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): d.copy_(pin, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print('H2D 78.6MB: %.2f ms  %.1f GB/s'%(dt*1e3, host.nbytes/dt/1e9))
out = torch.empty(16_000_000, dtype=torch.uint8).pin_memory(); dd = torch.empty(16_000_000, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): out.copy_(dd, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
print('D2H 16MB: %.2f ms  %.1f GB/s'%(dt*1e3, 16e6/dt/1e9))
for chunk in (256, 128, 64, 32, 16):
    os.environ['ORBX_CHUNK']=str(chunk)
    ex = api.ORBextractor(nfeatures=1000)
    kcap = ex.max_keypoints()
    kps_h = torch.empty((B, kcap, 28), dtype=torch.uint8).pin_memory().numpy()
    desc_h = torch.empty((B, kcap, 32), dtype=torch.uint8).pin_memory().numpy()
    n_h = np.zeros(B, np.int32)
    a = pin.numpy()
    def step():
        api._check(api.lib().orbx_extract_batch(ex._h, C.c_void_p(a.ctypes.data), B, 640, 480, 640, 640*480, C.c_void_p(kps_h.ctypes.data), C.c_void_p(desc_h.ctypes.data), kcap, C.c_void_p(n_h.ctypes.data)))
    for _ in range(3): step()
    t=time.perf_counter()
    for _ in range(10): step()
    dt=(time.perf_counter()-t)/10
    print('chunk',chunk,'%.2f ms/step  %.0f fps'%(dt*1e3, B/dt))
