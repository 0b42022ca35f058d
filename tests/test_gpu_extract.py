"""GPU parity of the extractor against the CPU oracle, stage by stage and end to end, through the C ABI.
Bit-exact bar for pyramid pixels, FAST candidates (set AND order), selected keypoints (set AND order), blur pixels
and descriptors; angles within 1e-3 degree (they are in fact expected bit-equal)."""
import numpy as np
import pytest

from orb_slam2_refactored_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_stages(o, img, nfeatures):
    e = o.extractor(nfeatures)
    kps, desc = e.extract(img)
    pyr = e.pyramid()
    quotas = o.quotas(nfeatures, 1.2, 8)
    cands = [o.detect_fast(p) for p in pyr]
    sels = [o.quadtree(c, p.shape[1], p.shape[0], int(q)) for c, p, q in zip(cands, pyr, quotas)]
    blurs = [o.gaussian7(p) for p in pyr]
    return kps, desc, pyr, cands, sels, blurs


def _as_xyr(c):
    return np.stack([c['x'], c['y'], c['response']], 1).astype(np.int32) if len(c) else np.zeros((0, 3), np.int32)


@pytest.mark.parametrize('cfg,seed', [('C1', 0), ('C1', 3), ('C2', 1), ('C3', 2)])
def test_stages_bit_exact(orbx, oracle_port, cfg, seed):
    c = synth.CONFIGS[cfg]
    img = synth.image(seed, c['w'], c['h'])
    ex = orbx.ORBextractor(nfeatures=c['nfeatures'])
    kps, desc = ex.Extract(img)
    okps, odesc, opyr, ocands, osels, oblurs = _oracle_stages(oracle_port, img, c['nfeatures'])

    pyr = ex.GetImagePyramid()
    for s in range(8):
        assert pyr[s].shape == opyr[s].shape
        assert np.array_equal(pyr[s], opyr[s]), f'pyramid level {s} differs in {(pyr[s] != opyr[s]).sum()} px'
    for s in range(8):
        got = ex.debug_candidates(0, s)
        assert np.array_equal(got, _as_xyr(ocands[s])), f'FAST candidates differ at level {s}: {len(got)} vs {len(ocands[s])}'
    for s in range(8):
        got = ex.debug_selected(0, s)
        assert np.array_equal(got, _as_xyr(osels[s])), f'quadtree selection differs at level {s}: {len(got)} vs {len(osels[s])}'
    for s in range(8):
        assert np.array_equal(ex.debug_blurred(0, s), oblurs[s]), f'blur level {s} differs'

    assert len(kps) == len(okps)
    for name in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
        assert np.array_equal(kps[name], okps[name]), name
    dang = np.abs(kps['angle'] - okps['angle'])
    dang = np.minimum(dang, 360 - dang)
    assert dang.max() <= 1e-3, f'angle differs by {dang.max()} deg'          # tolerance stated by BASELINE.json north_star
    same_angle = kps['angle'] == okps['angle']
    assert np.array_equal(desc[same_angle], odesc[same_angle])               # bit-exact wherever the angle agrees
    assert same_angle.all(), f'{(~same_angle).sum()} angles not bit-equal'
    assert np.array_equal(desc, odesc)


def test_batch_equals_single_and_ref(orbx, oracle_port, oracle_ref):
    c = synth.CONFIGS['C1']
    imgs = np.stack([synth.image(s, c['w'], c['h']) for s in range(6)])
    ex = orbx.ORBextractor(nfeatures=c['nfeatures'])
    kb, db = ex.ExtractBatch(imgs)
    e = oracle_ref.extractor(c['nfeatures'])
    for f in range(len(imgs)):
        okps, odesc = e.extract(imgs[f])
        assert kb[f].tobytes() == okps.tobytes(), f'frame {f} keypoints'
        assert np.array_equal(db[f], odesc), f'frame {f} descriptors'
    # a smaller batch on the same handle reuses the plan
    k1, d1 = ex.Extract(imgs[4])
    assert k1.tobytes() == kb[4].tobytes() and np.array_equal(d1, db[4])


def test_device_resident_batch(orbx, oracle_final):
    import torch
    c = synth.CONFIGS['C1']
    imgs = np.stack([synth.image(s + 10, c['w'], c['h']) for s in range(4)])
    ex = orbx.ORBextractor(nfeatures=c['nfeatures'])
    d = torch.from_numpy(imgs).cuda()
    kps, desc, n = ex.extract_batch_device(d)
    ex.synchronize()
    n = n.cpu().numpy()
    kps = kps.cpu().numpy().view(np.uint8).reshape(len(imgs), -1, 28)
    desc = desc.cpu().numpy()
    e = oracle_final.extractor(c['nfeatures'])
    for f in range(len(imgs)):
        okps, odesc = e.extract(imgs[f])
        assert n[f] == len(okps)
        assert kps[f, :n[f]].tobytes() == okps.tobytes()
        assert np.array_equal(desc[f, :n[f]], odesc)


@pytest.mark.parametrize('pad,shift', [(16, 0), (10, 0), (32, 4), (0, 0)])
def test_device_resident_layouts(orbx, oracle_final, pad, shift):
    """orbx_extract_batch_device reads 16-byte aligned frames IN PLACE as level 0 (pitch / frame stride / base multiples of 16) and copies
    every other layout into its padded buffer first: (16, 0) in place with pitch != width, (10, 0) odd pitch -> copy, (32, 4) base off by 4 bytes
    -> copy, (0, 0) contiguous -> in place. Same keypoints and descriptors either way, and the level-0 view of GetImagePyramid is the frame."""
    import torch
    c = synth.CONFIGS['C1']
    w, h = c['w'], c['h']
    imgs = np.stack([synth.image(40 + s, w, h) for s in range(3)])
    big = torch.zeros((3, h, w + pad), dtype=torch.uint8, device='cuda')
    view = big[:, :, shift:shift + w]
    view.copy_(torch.from_numpy(imgs).cuda())
    ex = orbx.ORBextractor(nfeatures=c['nfeatures'])
    kps, desc, n = ex.extract_batch_device(view)
    ex.synchronize()
    n = n.cpu().numpy(); kps = kps.cpu().numpy().view(np.uint8).reshape(3, -1, 28); desc = desc.cpu().numpy()
    e = oracle_final.extractor(c['nfeatures'])
    for f in range(3):
        okps, odesc = e.extract(imgs[f])
        assert n[f] == len(okps) and kps[f, :n[f]].tobytes() == okps.tobytes() and np.array_equal(desc[f, :n[f]], odesc), f'frame {f}'
    assert np.array_equal(ex.GetImagePyramid(1)[0], imgs[1])
    # a host-buffer call on the same handle afterwards (its level 0 lives in the handle's own buffer again)
    k1, d1 = ex.Extract(imgs[2])
    okps, odesc = e.extract(imgs[2])
    assert k1.tobytes() == okps.tobytes() and np.array_equal(d1, odesc)


def test_strided_and_misaligned_input(orbx, oracle_final):
    # a sub-matrix view (row stride != width, base not 16-byte aligned), as cv::Mat ROIs are
    c = synth.CONFIGS['C1']
    big = synth.image(5, c['w'] + 13, c['h'] + 4)
    view = big[2:2 + c['h'], 5:5 + c['w']]
    ex = orbx.ORBextractor(nfeatures=500)
    import ctypes as C
    from orb_slam2_refactored_b200.api import KP_DTYPE, _check, _p, lib
    cap = ex.max_keypoints()
    kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); n = C.c_int()
    _check(lib().orbx_extract(ex._h, C.c_void_p(view.ctypes.data), c['w'], c['h'], view.strides[0], _p(kps), _p(desc), cap, C.byref(n)))
    okps, odesc = oracle_final.extractor(500).extract(np.ascontiguousarray(view))
    assert n.value == len(okps) and kps[:n.value].tobytes() == okps.tobytes() and np.array_equal(desc[:n.value], odesc)


def test_flat_image_yields_nothing(orbx):
    # N == 0 path (src/ORBextractor.cc:778-782): outputs untouched, n = 0
    ex = orbx.ORBextractor(nfeatures=500)
    k, d = ex.Extract(np.full((480, 640), 77, np.uint8))
    assert len(k) == 0 and len(d) == 0


def test_tie_blocks_image(orbx, oracle_final):
    # flat 8x8 blocks: adjacent equal scores suppress each other under strict NMS, so cells retry at minTh (App. A.4)
    r = np.random.RandomState(3)
    img = np.kron(r.randint(0, 256, (60, 80)), np.ones((8, 8))).astype(np.uint8)
    ex = orbx.ORBextractor(nfeatures=1000)
    k, d = ex.Extract(img)
    ok, od = oracle_final.extractor(1000).extract(img)
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)


@pytest.mark.parametrize('lo,hi', [(0, 256), (100, 119), (96, 136), (90, 150)])
def test_noise_images_overflow_the_cell_list(orbx, oracle_port, lo, hi):
    # white noise flags most pixels of a cell: more than the cell kernel's list holds (FT_LIST_CAP), at iniTh ((0, 256)), only at
    # minTh ((100, 119): no pixel can pass iniTh), or only once the minTh pixels join ((96, 136), (90, 150)) — the chunked passes
    r = np.random.RandomState(hi)
    img = r.randint(lo, hi, (240, 320)).astype(np.uint8)
    ex = orbx.ORBextractor(nfeatures=1500)
    k, d = ex.Extract(img)
    o = oracle_port.extractor(1500)
    ok, od = o.extract(img)
    for s, pl in enumerate(o.pyramid()):
        want = _as_xyr(oracle_port.detect_fast(pl))
        got = ex.debug_candidates(0, s)
        assert np.array_equal(got, want), f'FAST candidates differ at level {s}: {len(got)} vs {len(want)}'
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    # a throughput launch (more than 16 frames) runs the cell kernel with its short list and hands these cells to the overflow launch
    kb, db = ex.ExtractBatch(np.stack([img, img[::-1].copy()] * 10))
    ok2, od2 = oracle_port.extractor(1500).extract(img[::-1].copy())
    for f in range(20):
        wk, wd = (ok, od) if f % 2 == 0 else (ok2, od2)
        assert kb[f].tobytes() == wk.tobytes() and np.array_equal(db[f], wd), f'frame {f} of the batch'


def test_handle_survives_a_refused_size(orbx, oracle_final):
    # good extract, then a size the plan builder refuses, then the first size again on the SAME handle: the size cache must not
    # match a plan that the failed rebuild zeroed (illegal address and a sticky context error before the fix)
    c = synth.CONFIGS['C1']
    img = synth.image(21, c['w'], c['h'])
    ex = orbx.ORBextractor(nfeatures=c['nfeatures'])
    k0, d0 = ex.Extract(img)
    with pytest.raises(orbx.OrbxError) as e:
        ex.Extract(np.zeros((100, 100), np.uint8))
    assert e.value.status == orbx.ORBX_ERR_INVALID
    k1, d1 = ex.Extract(img)
    ok, od = oracle_final.extractor(c['nfeatures']).extract(img)
    assert k0.tobytes() == ok.tobytes() and k1.tobytes() == ok.tobytes() and np.array_equal(d0, od) and np.array_equal(d1, od)
    # and a different valid size right after the failure
    img2 = synth.image(22, 752, 480)
    with pytest.raises(orbx.OrbxError):
        ex.Extract(np.zeros((640, 300), np.uint8))
    k2, d2 = ex.Extract(img2)
    ok2, od2 = oracle_final.extractor(c['nfeatures']).extract(img2)
    assert k2.tobytes() == ok2.tobytes() and np.array_equal(d2, od2)


def test_stereo_result_buffers_follow_the_last_extract(orbx):
    # orbx_stereo_match writes last_frames x last_cap floats: the mirror sizes its arrays from orbx_last_result_shape, whatever cap the
    # device-resident extract was given (a cap above max_keypoints() overran the host arrays)
    import torch
    c = synth.CONFIGS['C3']
    L, R = synth.stereo_pair(31, c['w'], c['h'])
    eL = orbx.ORBextractor(nfeatures=c['nfeatures']); eR = orbx.ORBextractor(nfeatures=c['nfeatures'])
    dL = torch.from_numpy(np.stack([L, L[::-1].copy()])).cuda(); dR = torch.from_numpy(np.stack([R, R[::-1].copy()])).cuda()
    lib = orbx.lib()
    import ctypes as C
    orbx._check(lib.orbx_plan(eL._h, c['w'], c['h'], 2))
    cap = eL.max_keypoints() + 37                     # a roomier layout than the plan's own
    outs = []
    for e, d in ((eL, dL), (eR, dR)):
        k = torch.empty((2, cap, 7), dtype=torch.float32, device='cuda'); ds = torch.empty((2, cap, 32), dtype=torch.uint8, device='cuda')
        n = torch.empty((2,), dtype=torch.int32, device='cuda')
        e.extract_batch_device(d, k, ds, n)
        outs.append((k, ds, n))
    ur, dp = orbx.ComputeStereoMatchesResident(eL, eR, c['camera'])
    assert ur.shape == (2, cap) and dp.shape == (2, cap)
    # the same pair through the plan's own layout gives the same matches
    kl, dl = eL.ExtractBatch(np.stack([L, L[::-1].copy()])); kr, dr = eR.ExtractBatch(np.stack([R, R[::-1].copy()]))
    ur2, dp2 = orbx.ComputeStereoMatchesResident(eL, eR, c['camera'])
    for f in range(2):
        n = len(kl[f])
        assert ur[f, :n].tobytes() == ur2[f, :n].tobytes() and dp[f, :n].tobytes() == dp2[f, :n].tobytes()


def test_input_contract(orbx):
    ex = orbx.ORBextractor(nfeatures=500)
    with pytest.raises(orbx.OrbxError) as e:
        ex.Extract(np.zeros((100, 100), np.uint8))       # level 7 would be 28 px: the reference divides by zero
    assert e.value.status == orbx.ORBX_ERR_INVALID
    with pytest.raises(orbx.OrbxError):
        ex.Extract(np.zeros((640, 300), np.uint8))       # portrait: cvRound(w/h) == 0 in the reference
    with pytest.raises(orbx.OrbxError):
        orbx.ORBextractor(nfeatures=500, minThFAST=0)


def test_getters_match_reference_tables(orbx, oracle_port):
    ex = orbx.ORBextractor(nfeatures=1000)
    t = oracle_port.extractor(1000).tables()
    for a, b in zip((ex.GetScaleFactors(), ex.GetInverseScaleFactors(), ex.GetScaleSigmaSquares(), ex.GetInverseScaleSigmaSquares()), t):
        assert a.tobytes() == b.tobytes()
    assert ex.GetLevels() == 8 and ex.GetScaleFactor() == np.float32(1.2)
    assert list(ex.GetFeatureQuotas()) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(orbx.ORBextractor(nfeatures=8000).GetFeatureQuotas()) == [1737, 1448, 1207, 1005, 838, 698, 582, 485]


def test_plan_rebuild_and_odd_batches(orbx, oracle_final):
    # one handle, changing image sizes and batch sizes (plan rebuild, chunk pipeline with a ragged last chunk)
    ex = orbx.ORBextractor(nfeatures=700)
    e = oracle_final.extractor(700)
    # (320 x 240 and 1241 x 376 as throughput launches: levels lower than one 64-row blur tile, cells taller than 32 rows)
    for (w, h, frames) in ((640, 480, 3), (752, 480, 1), (640, 480, 70), (400, 300, 5), (320, 240, 20), (1241, 376, 18), (640, 480, 2)):
        imgs = np.stack([synth.image(1000 + w + s, w, h) for s in range(min(frames, 6))])
        if frames > len(imgs):
            imgs = np.concatenate([imgs] * (frames // len(imgs) + 1))[:frames]
        k, d = ex.ExtractBatch(imgs)
        assert len(k) == frames
        ref = {}
        for f in range(frames):
            key = f % 6
            if key not in ref:
                ref[key] = e.extract(imgs[f])
            assert k[f].tobytes() == ref[key][0].tobytes() and np.array_equal(d[f], ref[key][1]), (w, h, frames, f)


def test_two_extractors_on_two_threads(orbx, oracle_final):
    # src/System.cc:449-452 runs the left and right extractor on two std::threads
    import threading
    c = synth.CONFIGS['C3']
    L, R = synth.stereo_pair(31, c['w'], c['h'])
    exs = [orbx.ORBextractor(nfeatures=c['nfeatures']) for _ in range(2)]
    out = [None, None]

    def work(i, img):
        for _ in range(5):
            out[i] = exs[i].Extract(img)
    th = [threading.Thread(target=work, args=(0, L)), threading.Thread(target=work, args=(1, R))]
    for t in th: t.start()
    for t in th: t.join()
    for i, img in enumerate((L, R)):
        ok, od = oracle_final.extractor(c['nfeatures']).extract(img)
        assert out[i][0].tobytes() == ok.tobytes() and np.array_equal(out[i][1], od)


@pytest.mark.parametrize('w,h,nf,scale,nlevels,ini,mn', [
    (640, 480, 1000, 1.1, 8, 20, 7),        # slow pyramid: every source row feeds two output rows
    (640, 480, 1000, 1.5, 4, 20, 7),
    (1241, 376, 1500, 2.0, 3, 20, 7),       # the steepest pyramid the resize tiles take
    (800, 600, 300, 1.2, 12, 20, 7),        # 12 levels (ORBX_MAX_LEVELS), small quota: the quadtree stops early everywhere
    (752, 480, 5000, 1.2, 8, 12, 5),        # more keypoints than corners on the upper levels, low thresholds
    (640, 480, 1000, 1.2, 8, 40, 30),       # high thresholds: many empty cells
    (640, 480, 1000, 1.2, 1, 20, 7),        # a single level
    (333, 251, 500, 1.3, 5, 20, 7),         # odd sizes, widths that are no multiple of 4
])
def test_parameter_sweep(orbx, oracle_port, w, h, nf, scale, nlevels, ini, mn):
    """Extract with non-default ORBextractor::Parameters (include/ORBextractor.h:38-50): keypoints and descriptors bit-equal, and the
    pyramid levels too (the resize tables depend on the scale factor)."""
    ex = orbx.ORBextractor(nfeatures=nf, scaleFactor=scale, nlevels=nlevels, iniThFAST=ini, minThFAST=mn)
    e = oracle_port.extractor(nf, scale, nlevels, ini, mn)
    for seed in (31, 32):
        img = synth.image(seed, w, h)
        k, d = ex.Extract(img)
        ok, od = e.extract(img)
        assert len(ok) > 50
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od), (w, h, nf, scale, nlevels, seed)
    for a, b in zip(ex.GetImagePyramid(), e.pyramid()):
        assert np.array_equal(a, b)
    assert np.array_equal(ex.GetFeatureQuotas(), oracle_port.quotas(nf, scale, nlevels))


def test_one_frame_graph_replay_between_other_calls(orbx, oracle_final):
    # A frame at a time through the host-buffer API replays a captured graph of the grouped launches (levels 0 / 1-2 / 3.. on parallel
    # streams). The replay must follow whatever else the handle did in between: other images, a small batch (another graph), a
    # device-resident batch that reads the caller's buffer in place (level-0 tensor maps re-encoded), a throughput batch, a re-plan.
    import torch
    c = synth.CONFIGS['C1']
    imgs = [synth.image(40 + s, c['w'], c['h']) for s in range(5)]
    small = synth.image(77, 400, 300)
    ex = orbx.ORBextractor(nfeatures=c['nfeatures'])
    e = oracle_final.extractor(c['nfeatures'])
    ref = [e.extract(im) for im in imgs]
    ref_small = e.extract(small)

    def check(got, want, what):
        assert got[0].tobytes() == want[0].tobytes() and np.array_equal(got[1], want[1]), what
    for rnd in range(2):
        for i in (0, 1, 0, 2):
            check(ex.Extract(imgs[i]), ref[i], ('one frame', rnd, i))
        k, d = ex.ExtractBatch(np.stack([imgs[3], imgs[4], imgs[1]]))
        for j, i in enumerate((3, 4, 1)):
            check((k[j], d[j]), ref[i], ('batch of three', rnd, i))
        check(ex.Extract(imgs[2]), ref[2], ('after the small batch', rnd))
        dv = torch.from_numpy(np.stack([imgs[(s + rnd) % 5] for s in range(20)])).cuda()
        kk, dd, nn = ex.extract_batch_device(dv)
        ex.synchronize()
        n0 = int(nn[3].item())
        assert kk[3, :n0].cpu().numpy().tobytes() == ref[(3 + rnd) % 5][0].tobytes(), ('device batch', rnd)
        check(ex.Extract(imgs[4]), ref[4], ('after the device-resident batch', rnd))
        k, d = ex.ExtractBatch(np.stack([imgs[s % 5] for s in range(70)]))
        check((k[66], d[66]), ref[1], ('throughput batch', rnd))
        check(ex.Extract(imgs[0]), ref[0], ('after the throughput batch', rnd))
        check(ex.Extract(small), ref_small, ('re-plan', rnd))
        check(ex.Extract(imgs[1]), ref[1], ('after the re-plan', rnd))
