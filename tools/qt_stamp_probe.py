"""Phase stamps (%globaltimer) of the quadtree kernel for level 0 of frame 0: ORBX_QT_STAMPS=1 python tools/qt_stamp_probe.py [C1|C4] [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from orb_slam2_refactored_b200 import api, synth
name = sys.argv[1] if len(sys.argv) > 1 else 'C1'
F = int(sys.argv[2]) if len(sys.argv) > 2 else 1
c = synth.CONFIGS[name]
ex = api.ORBextractor(nfeatures=c['nfeatures'])
img = np.stack([synth.image(s, c['w'], c['h']) for s in range(F)])
for _ in range(3): ex.ExtractBatch(img)
print('candidates per level:', [len(ex.debug_candidates(0, s)) for s in range(8)], 'selected:', [len(ex.debug_selected(0, s)) for s in range(8)])
