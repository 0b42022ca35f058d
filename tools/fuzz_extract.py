"""Parity fuzzer (run on a GPU box): Extract on random images of many kinds / sizes / parameters through the C ABI against the CPU
oracle (restatement), bit for bit; on a difference, says which stage differs first. usage: fuzz_extract.py [seconds] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_refactored_b200 import api, synth
from oracle import bindings

bindings.build()
O = bindings.Oracle('port')
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0


def make(r, w, h):
    kind = r.randint(0, 9)
    if kind == 0:
        lo = r.randint(0, 200); hi = lo + r.randint(2, 256 - lo)
        return 'noise[%d,%d)' % (lo, hi), r.randint(lo, hi, (h, w)).astype(np.uint8)
    if kind == 1:
        b = r.randint(1, 12)
        img = np.kron(r.randint(0, 256, ((h + b - 1) // b, (w + b - 1) // b)), np.ones((b, b)))[:h, :w]
        return 'blocks%d' % b, img.astype(np.uint8)
    if kind == 2:
        img = (r.randint(0, 2, (h, w)) * 255)
        return 'binary', img.astype(np.uint8)
    if kind == 3:
        g = np.linspace(0, 255, w)[None, :] + r.randint(-8, 9, (h, w))
        return 'ramp+noise', np.clip(g, 0, 255).astype(np.uint8)
    if kind == 4:
        yy, xx = np.mgrid[0:h, 0:w]
        p = r.randint(1, 5)
        return 'checker%d' % p, ((((xx // p) + (yy // p)) & 1) * r.randint(30, 256)).astype(np.uint8)
    if kind == 5:
        img = synth.image(r.randint(0, 1 << 30), w, h).astype(np.int32)
        img[r.randint(0, h):, :] = 255 if r.randint(0, 2) else 0      # a saturated band
        return 'synth+band', img.astype(np.uint8)
    if kind == 6:
        n = r.randint(5, 200)
        img = np.full((h, w), r.randint(0, 256), np.int32)
        for _ in range(n):
            x, y, s = r.randint(0, w), r.randint(0, h), r.randint(1, 30)
            img[y:y + s, x:x + s] = r.randint(0, 256)
        return 'rects%d' % n, img.astype(np.uint8)
    if kind == 7:
        a = r.randint(1, 12)
        img = 128 + r.randint(-a, a + 1, (h, w))
        return 'lownoise%d' % a, img.astype(np.uint8)
    return 'synth', synth.image(r.randint(0, 1 << 30), w, h)


t0 = time.time(); n = 0; bad = 0; ran = 0
while time.time() - t0 < budget:
    r = np.random.RandomState(seed0 * 100003 + n)
    n += 1
    w = int(r.randint(96, 1000)); h = int(r.randint(80, min(w, 720) + 1))
    nl = int(r.randint(1, 9)); sf = float(r.choice([1.1, 1.2, 1.2, 1.2, 1.3, 1.5, 2.0]))
    nf = int(r.choice([100, 500, 1000, 1000, 2000, 5000]))
    ini = int(r.randint(2, 60)); mn = int(r.randint(1, ini + 1))
    name, img = make(r, w, h)
    try:
        ex = api.ORBextractor(nfeatures=nf, scaleFactor=sf, nlevels=nl, iniThFAST=ini, minThFAST=mn)
        k, d = ex.Extract(img)
    except api.OrbxError as e:
        continue            # refused geometry (a level under 62 px ...): the reference divides by zero there
    ran += 1
    oe = O.extractor(nf, sf, nl, ini, mn)
    ok, od = oe.extract(img)
    if k.tobytes() != ok.tobytes() or not np.array_equal(d, od):
        bad += 1
        print(f'MISMATCH case {n - 1} seed0 {seed0}: {name} {w}x{h} nf={nf} sf={sf} nl={nl} th={ini}/{mn}: {len(k)} vs {len(ok)} keypoints', flush=True)
        for s, pl in enumerate(oe.pyramid()):
            if not np.array_equal(ex.GetImagePyramid()[s], pl): print('   pyramid differs at level', s); break
            c = O.detect_fast(pl, ini, mn)
            want = np.stack([c['x'], c['y'], c['response']], 1).astype(np.int32) if len(c) else np.zeros((0, 3), np.int32)
            if not np.array_equal(ex.debug_candidates(0, s), want): print('   FAST candidates differ at level', s); break
        if len(k) == len(ok):
            print('   kp fields equal:', {f: bool(np.array_equal(k[f], ok[f])) for f in k.dtype.names}, 'desc rows differing', int((d != od).any(1).sum()))
print(f'{n} cases drawn, {ran} extracted and compared (the rest refused: a level under 62 px), {time.time() - t0:.0f} s, {bad} mismatches')
sys.exit(1 if bad else 0)
