"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): one VGA frame, one wide stereo pair, a small kNN."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
ex = api.ORBextractor(nfeatures=1000)
k, d = ex.ExtractBatch(np.stack([synth.image(0, 640, 480), synth.image(1, 640, 480)]))
L, R = synth.stereo_pair(2, 700, 240)
eL, eR = api.ORBextractor(nfeatures=400), api.ORBextractor(nfeatures=400)
kl, dl = eL.Extract(L); kr, dr = eR.Extract(R)
ur, dp = api.ComputeStereoMatchesResident(eL, eR, synth.EUROC_CAMERA)
q, t = synth.planted_descriptors(1, 700, 3001)
r = api.ORBmatcher(0.6).knn2(q, t)
print('ok', len(k[0]), len(k[1]), int((dp[0] > 0).sum()), int((r[3] >= 0).sum()))
