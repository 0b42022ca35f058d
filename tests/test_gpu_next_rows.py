"""GPU parity of the SURVEY §8(f) rows built so far: ConvertToGray (#3, standalone and fused into Extract),
ComputeStereoFromRGBD (#3) and the N x N Hamming site of MapPoint::ComputeDistinctiveDescriptors (#4). Bit-exact bar."""
import numpy as np
import pytest

from orb_slam2_refactored_b200 import synth

pytestmark = pytest.mark.gpu


def _color(seed, w, h, ch):
    g = synth.image(seed, w, h).astype(np.int16)
    r = np.random.RandomState(seed)
    img = np.stack([np.clip(g + r.randint(-40, 41, g.shape), 0, 255) for _ in range(ch)], -1).astype(np.uint8)
    return img


@pytest.mark.parametrize('ch,rgb', [(3, True), (3, False), (4, True), (4, False)])
def test_convert_to_gray(orbx, oracle_port, ch, rgb):
    img = _color(ch * 2 + rgb, 643, 481, ch)     # width not a multiple of 4
    assert np.array_equal(orbx.ConvertToGray(img, rgb), oracle_port.convert_to_gray(img, rgb))
    mono = img[..., 0]
    assert orbx.ConvertToGray(mono, rgb) is mono                                    # ch == 1: dst = src (src/System.cc:129-133)


def test_extract_color_fused(orbx, oracle_port):
    imgs = np.stack([_color(s, 640, 480, 3) for s in range(3)])
    ex = orbx.ORBextractor(nfeatures=1000)
    k, d = ex.ExtractBatchColor(imgs, RGB=False)
    e = oracle_port.extractor(1000)
    for f in range(len(imgs)):
        ok, od = e.extract(oracle_port.convert_to_gray(imgs[f], False))
        assert k[f].tobytes() == ok.tobytes() and np.array_equal(d[f], od)
    assert np.array_equal(ex.GetImagePyramid(1)[0], oracle_port.convert_to_gray(imgs[1], False))


def test_stereo_from_rgbd(orbx, oracle_port):
    img = synth.image(9, 640, 480)
    ex = orbx.ORBextractor(nfeatures=1000)
    kps, _ = ex.Extract(img)
    un = kps.copy(); un['x'] += np.float32(0.37)
    r = np.random.RandomState(4)
    dm = r.uniform(-0.5, 6.0, (480, 640)).astype(np.float32)
    dm[r.rand(480, 640) < 0.2] = 0
    ur, dp = orbx.ComputeStereoFromRGBD(kps, un, dm, synth.KITTI_CAMERA)
    wu, wd = oracle_port.stereo_from_rgbd(kps, un, dm, synth.KITTI_CAMERA)
    assert ur.tobytes() == wu.tobytes() and dp.tobytes() == wd.tobytes()
    assert 0 < (wd > 0).sum() < len(kps)


def test_distinctive_descriptors(orbx, oracle_port):
    r = np.random.RandomState(5)
    sets = []
    for n in (1, 2, 3, 4, 7, 16, 33, 64, 150, 0, 5):
        base = synth.descriptors(n + 1, 1)[0]
        d = np.repeat(base[None], n, 0)
        for i in range(n):          # observations of one map point: noisy copies of one descriptor (+ exact duplicates -> ties)
            bits = np.unpackbits(d[i]); flip = r.choice(256, r.randint(0, 40), replace=False); bits[flip] ^= 1; d[i] = np.packbits(bits)
        if n > 4:
            d[n - 1] = d[1]
        sets.append(d)
    got = orbx.ComputeDistinctiveDescriptors(sets)
    want = np.array([oracle_port.distinctive_index(d) if len(d) else -1 for d in sets], np.int32)
    assert np.array_equal(got, want)


def test_remap_standalone(orbx, oracle_port):
    """cv::remap(INTER_LINEAR) of the rectification step (Examples/Stereo/stereo_euroc.cc:100-101): a rectification warp at the EuRoC
    frame size, and an adversarial map (other output size, taps outside the source, 1/64-pixel rounding ties). Also against cv2's
    own output stored in tests/golden/primitives.npz."""
    import os
    from test_oracle_golden import adversarial_maps, _inputs, G
    img = synth.image(3, 752, 480)
    mx, my = synth.rectification_maps(3, 752, 480)
    assert np.array_equal(orbx.Remap(img, mx, my), oracle_port.remap(img, mx, my))
    ax, ay = adversarial_maps(4, 752, 480)
    assert np.array_equal(orbx.Remap(img, ax, ay), oracle_port.remap(img, ax, ay))
    prim = np.load(os.path.join(G, 'primitives.npz'))
    i = _inputs()
    for name in ('img', 'noise'):
        hh, ww = i[name].shape
        rx, ry = synth.rectification_maps(5, ww, hh)
        assert np.array_equal(orbx.Remap(i[name], rx, ry), prim[f'remap_rect_{name}'])
        bx, by = adversarial_maps(6, ww, hh)
        assert np.array_equal(orbx.Remap(i[name], bx, by), prim[f'remap_adv_{name}'])


def test_extract_rectified_fused(orbx, oracle_port):
    """Raw EuRoC-size frames in, remap fused into the upload, Extract on the rectified image: equals remap-then-Extract of the oracle;
    level 0 of the pyramid is the rectified image. The maps crop the output to 736x464 to exercise a size change."""
    raw = np.stack([synth.image(20 + s, 752, 480) for s in range(3)])
    mx, my = synth.rectification_maps(8, 752, 480)
    mx, my = np.ascontiguousarray(mx[8:472, 8:744]), np.ascontiguousarray(my[8:472, 8:744])
    ex = orbx.ORBextractor(nfeatures=1200)
    ex.SetRectification(mx, my, raw.shape[1:])
    k, d = ex.ExtractBatchRectified(raw)
    e = oracle_port.extractor(1200)
    for f in range(len(raw)):
        rect = oracle_port.remap(raw[f], mx, my)
        ok, od = e.extract(rect)
        assert len(ok) > 800
        assert k[f].tobytes() == ok.tobytes() and np.array_equal(d[f], od)
    assert np.array_equal(ex.GetImagePyramid(2)[0], oracle_port.remap(raw[2], mx, my))
    with pytest.raises(orbx.OrbxError):
        orbx.ORBextractor(nfeatures=500).ExtractBatchRectified(raw)        # no maps set
    with pytest.raises(orbx.OrbxError):
        ex.ExtractBatchRectified(raw[:, :400])                              # not the frame size the maps were set for


def test_undistort_keypoints(orbx, oracle_port):
    """UndistortKeyPoints (src/System.cc:153-174) on the keypoints of a real extraction with the TUM1 distortion, and on scattered points
    with 4-, 5- and 8-coefficient models against cv2's own output (tests/golden/primitives.npz). Bit-exact: FP64 on the device."""
    import os
    from test_oracle_golden import UNDISTORT_CASES, undistort_input, G
    prim = np.load(os.path.join(G, 'primitives.npz'))
    cam, dist = UNDISTORT_CASES['tum1']
    cam6 = cam + (40.0, 40.0 / cam[0])
    kps, _ = orbx.ORBextractor(nfeatures=1000).Extract(synth.image(4, 640, 480))
    got = orbx.UndistortKeyPoints(kps, cam6, dist)
    assert got.tobytes() == oracle_port.undistort_keypoints(kps, cam6, dist).tobytes()
    assert np.abs(got['x'] - kps['x']).max() > 1.0 and np.array_equal(got['angle'], kps['angle'])
    assert orbx.UndistortKeyPoints(kps, cam6, [0, 0, 0, 0, 0]).tobytes() == kps.tobytes()       # distCoeffs(0) == 0: dst = src
    pts = undistort_input()
    scattered = np.zeros(len(pts), kps.dtype)
    scattered['x'], scattered['y'] = pts[:, 0], pts[:, 1]
    for tag, (c, d) in UNDISTORT_CASES.items():
        out = orbx.UndistortKeyPoints(scattered, c + (40.0, 0.1), d)
        assert np.array_equal(np.stack([out['x'], out['y']], 1), prim[f'undist_{tag}']), tag
