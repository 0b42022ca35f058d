"""Parity fuzzer for ComputeStereoMatches (run on a GPU box): random stereo pairs (size, disparity, noise, keypoint budget, camera)
through Extract left + right and the resident stereo matcher against the CPU oracle, uright / depth byte for byte.
usage: fuzz_stereo.py [seconds] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
from oracle import bindings

bindings.build()
O = bindings.Oracle('port')
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = time.time(); n = 0; bad = 0; matched = 0
while time.time() - t0 < budget:
    r = np.random.RandomState(seed0 * 7919 + n)
    n += 1
    w = int(r.randint(200, 1300)); h = int(r.randint(120, min(w, 500) + 1))
    nf = int(r.choice([300, 1000, 2000, 3000]))
    disp = int(r.randint(0, 60)); noise = int(r.randint(0, 9))
    cam = list(synth.KITTI_CAMERA if r.randint(0, 2) else synth.EUROC_CAMERA)
    L, R = synth.stereo_pair(int(r.randint(0, 1 << 30)), w, h, disparity=disp, noise=noise)
    if r.randint(0, 4) == 0:                     # a low-texture pair: few keypoints, many rows without candidates
        L = (L.astype(np.int32) // 8 + 100).astype(np.uint8); R = (R.astype(np.int32) // 8 + 100).astype(np.uint8)
    try:
        eL = api.ORBextractor(nfeatures=nf); eR = api.ORBextractor(nfeatures=nf)
        kl, dl = eL.ExtractBatch(L[None]); kr, dr = eR.ExtractBatch(R[None])
        ur, dp = api.ComputeStereoMatchesResident(eL, eR, cam)
    except api.OrbxError:
        continue
    oL, oR = O.extractor(nf), O.extractor(nf)
    okl, odl = oL.extract(L); okr, odr = oR.extract(R)
    sc, inv, _, _ = oL.tables()
    rc, wu, wd = O.stereo(okl, odl, oL.pyramid(), okr, odr, oR.pyramid(), sc, inv, cam)
    m = len(okl)
    matched += int((wd > 0).sum())
    if kl[0].tobytes() != okl.tobytes() or ur[0, :m].tobytes() != wu.tobytes() or dp[0, :m].tobytes() != wd.tobytes():
        bad += 1
        print(f'MISMATCH case {n - 1} seed0 {seed0}: {w}x{h} nf={nf} disparity={disp} noise={noise}: keypoints equal {kl[0].tobytes() == okl.tobytes()}, '
              f'uright differs at {int((ur[0, :m] != wu).sum())}, depth at {int((dp[0, :m] != wd).sum())} of {m}', flush=True)
print(f'{n} stereo pairs compared, {matched} matches in total, {time.time() - t0:.0f} s, {bad} mismatches')
sys.exit(1 if bad else 0)
