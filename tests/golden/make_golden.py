"""Generates the committed golden fixtures. Run in the build container (needs cv2 4.13.0 and /root/reference):

    python tests/golden/make_golden.py

primitives.npz   outputs of cv2 4.13.0 (IPP off) for the four un-vendored primitives the reference calls
                 (cv::resize / cv::FAST / cv::GaussianBlur / cv::fastAtan2) on small synthetic inputs;
pipeline.npz     outputs of the REFERENCE's own translation units (oracle/_ref, built from
                 /root/reference/src/ORBextractor.cc and ORBmatcher.cc line ranges) for Extract, the stages,
                 ComputeStereoMatches and the best/second scan on small synthetic inputs.
bow.npz          outputs of the reference's own DBoW2 text (transform: BowVector + FeatureVector) on the small cases of tests/bow_cases.py.
guided.npz       outputs of the reference's guided matchers (src/Frame.cc, src/ORBmatcher.cc line ranges) on seeded scenes.
Inputs are regenerated from orb_slam2_refactored_b200/synth.py (numpy RandomState); a CRC of every input is stored so
that a silent change of the generator is caught.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cv2  # noqa: E402

from oracle import bindings  # noqa: E402
from orb_slam2_refactored_b200 import synth  # noqa: E402

cv2.ipp.setUseIPP(False)
assert cv2.__version__ == '4.13.0', cv2.__version__


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def fast_cv(img, th):
    k = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(img)
    return np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)


UNDISTORT_CASES = {
    'tum1': ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
    'k4': ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),          # EuRoC cam0, 4 coefficients
    'strong': ((300.0, 310.0, 320.0, 240.0), (-0.9, 2.5, 0.01, -0.02, -4.0, 0.3, -0.2, 0.1)),                     # 8 coefficients, hits icdist < 0
}


def adversarial_maps(seed, w, h):
    """Maps of a different size than the source with coordinates up to 10 px outside it and many exact x.5/32 ties."""
    r = np.random.RandomState(seed)
    dh, dw = h - 17, w - 11
    mx = (r.rand(dh, dw) * (w + 20) - 10).astype(np.float32)
    my = (r.rand(dh, dw) * (h + 20) - 10).astype(np.float32)
    mx[::3, ::5] = (np.floor(mx[::3, ::5] * 32) + 0.5) / 32      # sx * 32 lands on .5: cvRound goes to even
    my[1::4, ::7] = (np.floor(my[1::4, ::7] * 32) + 0.5) / 32
    return mx.astype(np.float32), my.astype(np.float32)


def primitives():
    out = {}
    img = synth.image(42, 160, 120)
    noise = np.random.RandomState(7).randint(0, 256, (97, 131)).astype(np.uint8)
    blocks = np.kron(np.random.RandomState(8).randint(0, 256, (12, 16)), np.ones((8, 8))).astype(np.uint8)
    out['img_crc'] = crc(img); out['noise_crc'] = crc(noise); out['blocks_crc'] = crc(blocks)
    for name, src in (('img', img), ('noise', noise)):
        for (dw, dh) in ((133, 100), (111, 83), (80, 61), (159, 119)):
            if dw <= src.shape[1]:
                out[f'resize_{name}_{dw}x{dh}'] = cv2.resize(src, (dw, dh))
        out[f'blur_{name}'] = cv2.GaussianBlur(src, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    for name, src in (('img', img), ('noise', noise), ('blocks', blocks)):
        for th in (20, 7):
            out[f'fast_{name}_{th}'] = fast_cv(src, th)
    rc = np.random.RandomState(11)
    for ch in (3, 4):
        col = rc.randint(0, 256, (61, 83, ch)).astype(np.uint8)
        out[f'color{ch}_crc'] = crc(col)
        out[f'gray{ch}_rgb'] = cv2.cvtColor(col, cv2.COLOR_RGB2GRAY if ch == 3 else cv2.COLOR_RGBA2GRAY)
        out[f'gray{ch}_bgr'] = cv2.cvtColor(col, cv2.COLOR_BGR2GRAY if ch == 3 else cv2.COLOR_BGRA2GRAY)
    # cv::remap(INTER_LINEAR), float maps: a rectification warp and an adversarial map (out of range, 1/64-pixel rounding ties)
    for name, src in (('img', img), ('noise', noise)):
        hh, ww = src.shape
        mx, my = synth.rectification_maps(5, ww, hh)
        out[f'remap_rect_{name}'] = cv2.remap(src, mx, my, cv2.INTER_LINEAR)
        ax, ay = adversarial_maps(6, ww, hh)
        out[f'remap_adv_{name}'] = cv2.remap(src, ax, ay, cv2.INTER_LINEAR)
        out[f'remap_{name}_crc'] = np.array([crc(mx), crc(my), crc(ax), crc(ay)])
    # cv::undistortPoints(pts, K, dist, None, K) with the TUM1 and a strong synthetic distortion (Examples/Monocular/TUM1.yaml:9-18)
    ru = np.random.RandomState(12)
    upts = np.stack([ru.rand(4000) * 660 - 10, ru.rand(4000) * 500 - 10], 1).astype(np.float32)
    out['undist_pts_crc'] = crc(upts)
    for tag, (cam, dist) in UNDISTORT_CASES.items():
        K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], np.float32)
        out[f'undist_{tag}'] = cv2.undistortPoints(upts.reshape(-1, 1, 2), K, np.array(dist, np.float32), None, K).reshape(-1, 2)
    r = np.random.RandomState(3)
    y = r.randint(-60000, 60000, 4000).astype(np.float32); x = r.randint(-60000, 60000, 4000).astype(np.float32)
    y[:50] = 0; x[50:100] = 0; y[100] = 0; x[100] = 0
    out['atan_y'] = y; out['atan_x'] = x
    out['atan_deg'] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    v = np.concatenate([np.arange(-8, 9) + 0.5, r.uniform(-1000, 1000, 200)]).astype(np.float32)
    out['round_in'] = v
    out['round_rne'] = np.rint(v).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, 'primitives.npz'), **out)
    print('primitives.npz', {k: getattr(val, 'shape', None) for k, val in list(out.items())[:6]}, '...')


def pipeline():
    bindings.build()
    ref = bindings.Oracle('ref')
    out = {}
    # Extract on two small frames (all stages) and one C1 frame (final outputs only)
    for tag, (seed, w, h, nf) in {'small': (1, 400, 300, 500), 'wide': (2, 700, 240, 400)}.items():
        img = synth.image(seed, w, h)
        e = ref.extractor(nf)
        kps, desc = e.extract(img)
        out[f'{tag}_args'] = np.array([seed, w, h, nf], np.int32)
        out[f'{tag}_crc'] = crc(img)
        out[f'{tag}_kps'] = kps.view(np.uint8).reshape(-1, 28); out[f'{tag}_desc'] = desc
        pyr = e.pyramid()
        quotas = ref.quotas(nf, 1.2, 8)
        out[f'{tag}_quotas'] = quotas
        for s in (0, 3, 7):
            c = ref.detect_fast(pyr[s])
            out[f'{tag}_cand{s}'] = c.view(np.int32).reshape(-1, 3)
            out[f'{tag}_sel{s}'] = ref.quadtree(c, pyr[s].shape[1], pyr[s].shape[0], int(quotas[s])).view(np.int32).reshape(-1, 3)
            out[f'{tag}_pyrcrc{s}'] = crc(pyr[s])
    img = synth.image(0, 640, 480)
    kps, desc = ref.extractor(1000).extract(img)
    out['c1_crc'] = crc(img); out['c1_kps'] = kps.view(np.uint8).reshape(-1, 28); out['c1_desc'] = desc
    # stereo on a small pair, EuRoC camera
    L, R = synth.stereo_pair(5, 480, 320)
    eL, eR = ref.extractor(600), ref.extractor(600)
    kl, dl = eL.extract(L); kr, dr = eR.extract(R)
    sc, inv, _, _ = eL.tables()
    rc, ur, dp = ref.stereo(kl, dl, eL.pyramid(), kr, dr, eR.pyramid(), sc, inv, synth.EUROC_CAMERA)
    out['stereo_crc'] = np.array([crc(L), crc(R)]); out['stereo_uright'] = ur; out['stereo_depth'] = dp
    out['stereo_nl'] = np.int32(len(kl))
    # best/second scan
    q, t = synth.planted_descriptors(9, 300, 3000)
    idx, best, second, match = ref.knn2(q, t, 50, 0.6)
    out['knn_crc'] = np.array([crc(q), crc(t)]); out['knn_idx'] = idx; out['knn_best'] = best; out['knn_second'] = second
    out['knn_match'] = match
    # constants of SURVEY §8(a) E0
    out['scale_factors'] = sc
    np.savez_compressed(os.path.join(HERE, 'pipeline.npz'), **out)
    print('pipeline.npz: c1', len(kps), 'kps; stereo matched', int((dp > 0).sum()), '; knn accepted', int((match >= 0).sum()))


def guided():
    """guided.npz: outputs of the reference text of FeaturesGrid / SearchByProjection x2 / SearchForInitialization (oracle/_ref, rule
    guided_gen.cc) on the small cases of tests/guided_cases.py."""
    sys.path.insert(0, os.path.dirname(HERE))
    import guided_cases as gc
    ref = bindings.Oracle('ref')
    out = {}
    for name, kind, c in gc.cases(small=True):
        res = gc.run_oracle(ref, kind, c)
        for k, v in res.items():
            out[f'{name}.{k}'] = v
        fr = c.get('frame') or c.get('f2') or c['scene']['f2']
        out[f'{name}.crc'] = np.array([crc(fr['kps_un']), crc(fr['desc'])])
    np.savez_compressed(os.path.join(HERE, 'guided.npz'), **out)
    print('guided.npz:', len(out), 'arrays')


def bow():
    """bow.npz: outputs of the reference's DBoW2 text (oracle/_ref, rule bow_gen.cc) on the small cases of tests/bow_cases.py; the
    vocabulary goes through a text file and the reference's own loadFromTextFile."""
    import tempfile
    sys.path.insert(0, os.path.dirname(HERE))
    import bow_cases as bc
    from orb_slam2_refactored_b200 import synth
    ref = bindings.Oracle('ref')
    out = {}
    for name in bc.SMALL:
        voc, feats, levelsup = bc.make(name)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, 'voc.txt')
            synth.write_vocabulary_text(voc, path)
            wi, wv, fv = ref.vocabulary(path=path).transform(feats, levelsup)
        for k, v in bc.flatten(((wi, wv), fv)).items():
            out[f'{name}_{k}'] = v
        out[f'{name}_features'] = feats
    np.savez_compressed(os.path.join(HERE, 'bow.npz'), **out)
    print('bow.npz:', len(out), 'arrays')


if __name__ == '__main__':
    if 'bow' in sys.argv[1:]:
        bow()
        sys.exit(0)
    primitives()
    pipeline()
    guided()
    bow()
    for f in ('primitives.npz', 'pipeline.npz', 'guided.npz', 'bow.npz'):
        print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')
