import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from orb_slam2_refactored_b200 import api, synth
ex = api.ORBextractor(nfeatures=1000)
img = synth.image(0, 640, 480)[None]
for _ in range(3): ex.ExtractBatch(img)
