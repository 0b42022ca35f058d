"""Device-resident throughput of a 512-frame batch (two stream lanes) under the scheduling knobs of orbx_extract_batch_device: ORBX_BLUR_POS
(where each lane runs its blur), ORBX_LEAD (which lane's first kernel reaches the GPU first), ORBX_PYR_ONE, ORBX_PRIO, ORBX_QT_PAD, ORBX_SPLIT.
One subprocess per setting because the knobs are read once per process.
usage: skew_probe.py [KEY=V,KEY=V ...]      (no arguments: the grid of blur positions x lead x pyramid launch form)"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRAMES = '/tmp/skew_probe_frames.npy'
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import numpy as np, torch
from orb_slam2_refactored_b200 import api
from orb_slam2_refactored_b200 import synth
cfg = os.environ.get('PROBE_CFG', 'C1'); B = int(os.environ.get('PROBE_B', '512'))
c = synth.CONFIGS[cfg]
if cfg == 'C1':
    host = np.load(%r); host = np.concatenate([host, host[:, ::-1]])[:B] if B > len(host) else host[:B]
else:
    base = np.stack([synth.image(2000 + s_, c['w'], c['h']) for s_ in range(4)])
    host = np.tile(base, ((B + 3) // 4, 1, 1))[:B]
t0 = torch.from_numpy(host).cuda()
d = [t0, torch.roll(t0, 77, 1).contiguous(), torch.flip(t0, dims=[1]).contiguous()]
ex = api.ORBextractor(nfeatures=c['nfeatures'])
outs = ex.extract_batch_device(d[0])
for i in range(4): ex.extract_batch_device(d[i %% 3], *outs)
ex.synchronize()
ts = []
for rep in range(3):
    t = time.perf_counter()
    n = 15 if cfg != 'C4' else 4
    for i in range(n): ex.extract_batch_device(d[i %% 3], *outs)
    ex.synchronize(); ts.append((time.perf_counter() - t) / n)
best = min(ts)
print('%%.4f ms/step %%.0f fps  reps(ms): %%s' %% (best * 1e3, B / best, ' '.join('%%.3f' %% (t * 1e3) for t in ts)))
''' % (ROOT, FRAMES)
if not os.path.exists(FRAMES):
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    np.save(FRAMES, bench.make_frames(512, 0))
if len(sys.argv) > 1:
    settings = [dict(kv.split('=') for kv in a.split(',') if kv) for a in sys.argv[1:]]
else:
    settings = [dict()]
    for one in '01':
        for lead in '12':
            for a in '0123':
                for b in '0123':
                    settings.append(dict(ORBX_PYR_ONE=one, ORBX_LEAD=lead, ORBX_BLUR_POS=a + b))
res = []
for s in settings:
    env = dict(os.environ); env.update(s)
    r = subprocess.run([sys.executable, '-c', CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().split('\n')[-1] if r.returncode == 0 else 'FAILED ' + r.stderr[-400:]
    print(' '.join('%s=%s' % (k.replace('ORBX_', ''), v) for k, v in s.items()) or '(defaults)', line, flush=True)
    if r.returncode == 0: res.append((float(line.split()[0]), s))
res.sort(key=lambda x: x[0])
print('best:')
for t, s in res[:8]: print('  %.4f' % t, s)
