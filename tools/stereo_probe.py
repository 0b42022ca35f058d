"""C2 (KITTI-shape stereo) step split: left + right extraction alone vs extraction + ComputeStereoMatches, device-resident.
usage: stereo_probe.py [pairs]"""
import os, sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
SB = int(sys.argv[1]) if len(sys.argv) > 1 else 64
c2 = synth.CONFIGS['C2']
base = [synth.stereo_pair(500 + s, c2['w'], c2['h']) for s in range(8)]
Ls = np.stack([base[i % 8][0] for i in range(SB)]); Rs = np.stack([base[i % 8][1] for i in range(SB)])
dL = torch.from_numpy(Ls).cuda(); dR = torch.from_numpy(Rs).cuda()
eL = api.ORBextractor(nfeatures=c2['nfeatures']); eR = api.ORBextractor(nfeatures=c2['nfeatures'])
oL = eL.extract_batch_device(dL); oR = eR.extract_batch_device(dR)
cap = oL[0].shape[1]
ur = torch.empty((SB, cap), dtype=torch.float32, device='cuda'); dp = torch.empty_like(ur)
cam = api._Camera(*[float(v) for v in c2['camera']])
def ext():
    eL.extract_batch_device(dL, *oL); eR.extract_batch_device(dR, *oR)
def full():
    ext()
    api._check(api.lib().orbx_stereo_match_device(eL._h, eR._h, C.byref(cam), C.c_void_p(ur.data_ptr()), C.c_void_p(dp.data_ptr())))
for name, fn in (('extract L+R', ext), ('extract + match', full)):
    for _ in range(3): fn()
    eL.synchronize(); eR.synchronize(); t = time.perf_counter()
    for _ in range(10): fn()
    eL.synchronize(); eR.synchronize(); dt = (time.perf_counter() - t) / 10
    print(f'{name:18s} {SB} pairs: {dt * 1e3:7.3f} ms/step  {SB / dt:9.0f} pairs/s')
