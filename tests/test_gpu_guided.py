"""GPU parity of the guided matchers through the C ABI: FeaturesGrid (src/Frame.cc:63-145), SearchByProjection for
local-map and motion-model tracking (src/ORBmatcher.cc:315-382, 1279-1362), SearchForInitialization (:614-694), SearchByBoW x2 (:406-516, 696-766) and CheckOrientation
(:249-309). Bit-exact against the CPU oracle and against the committed outputs of the reference text."""
import os

import numpy as np
import pytest

import guided_cases as gc
from orb_slam2_refactored_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_grid_layout(orbx, oracle_port):
    """AssignFeatures: cell = round-half-away(invW * (x - minx)), out-of-grid keypoints dropped, push_back order kept."""
    for seed, margin in ((0, 0.0), (1, 3.7), (2, 40.0)):
        fr = synth.frame(seed, n=2500, bounds_margin=margin)
        if seed == 2:   # undistorted keypoints may leave the image (src/Frame.cc:94-96)
            fr['kps_un']['x'][::17] -= 120.0
            fr['kps_un']['y'][::13] += 90.0
        f = gc.make_frame(orbx, fr)
        start, items = f.grid()
        b = fr['bounds']
        invW = np.float32(64) / np.float32(b[1] - b[0]); invH = np.float32(48) / np.float32(b[3] - b[2])
        vx = (invW * (fr['kps_un']['x'] - b[0])).astype(np.float32); vy = (invH * (fr['kps_un']['y'] - b[2])).astype(np.float32)
        cx = (np.sign(vx) * np.floor(np.abs(vx) + np.float32(0.5))).astype(np.int64)
        cy = (np.sign(vy) * np.floor(np.abs(vy) + np.float32(0.5))).astype(np.int64)
        inside = (cx >= 0) & (cx < 64) & (cy >= 0) & (cy < 48)
        cell = cx * 48 + cy
        want = [np.flatnonzero(inside & (cell == c)) for c in range(64 * 48)]
        assert start[-1] == inside.sum() == len(items)
        for c in range(64 * 48):
            assert np.array_equal(items[start[c]:start[c + 1]], want[c]), (seed, c)
        if seed == 2:
            assert inside.sum() < len(inside)


def test_all_cases_against_oracle(orbx, oracle_port):
    kinds, max_rounds = {}, 0
    for name, kind, c in gc.cases():
        want = gc.run_oracle(oracle_port, kind, c)
        got = gc.run_product(orbx, kind, c)
        for k, v in want.items():
            assert np.array_equal(got[k], v), (name, k)
        kinds[kind] = kinds.get(kind, 0) + 1
        max_rounds = max(max_rounds, got.get('_rounds', 0))
        if name == 'local_ladder':
            assert got['_rounds'] >= 12 and list(got['mp'][:12]) == list(range(12))
    assert set(kinds) == {'grid', 'local', 'last', 'init', 'bow', 'reloc', 'sim3', 'fuse', 'fuse_sim3', 'sim3_search', 'triang'} and max_rounds >= 12


def test_reference_golden(orbx):
    g = np.load(os.path.join(G, 'guided.npz'))
    n = 0
    for name, kind, c in gc.cases(small=True):
        got = gc.run_product(orbx, kind, c)
        for k in got:
            if not k.startswith('_'):
                assert np.array_equal(g[f'{name}.{k}'], got[k]), (name, k)
                n += 1
    assert n >= 25


def test_frame_reuse_and_state_codes(orbx, oracle_port):
    """One device frame serves several searches (TrackWithMotionModel then SearchLocalPoints); the second search starts from the
    map points the first one stored."""
    fr = synth.frame(7, n=2000)
    f = gc.make_frame(orbx, fr)
    cp, lp, lpts, ldesc = synth.last_frame_points(7, fr, synth.KITTI_CAMERA)
    m = orbx.ORBmatcher(0.9, True)
    n1 = m.SearchByProjectionLastFrame(f, synth.KITTI_CAMERA, cp, lp, lpts, ldesc, 7.0, False)
    w1, mp1 = oracle_port.search_last_frame(fr, synth.KITTI_CAMERA, cp, lp, np.full(2000, -1, np.int32), lpts, ldesc, 7.0, False, 0.9, True)
    assert n1 == w1 and np.array_equal(f.mappoints, mp1)
    # the local-map search sees those as "some other map point": with observations when the last-frame point had them
    state = np.where(mp1 >= 0, np.where((lpts['flags'][np.maximum(mp1, 0)] & 2) != 0, -2, -3), -1).astype(np.int32)
    f.mappoints[:] = state
    pts, desc = synth.local_map_points(8, fr, npts=2500)
    n2 = orbx.ORBmatcher(0.8, True).SearchByProjection(f, pts, desc, 3.0)
    w2, mp2 = oracle_port.search_local_map(fr, state, pts, desc, 3.0, 0.8)
    assert n2 == w2 and np.array_equal(f.mappoints, mp2)


def test_large_frame_and_list_regrowth(orbx, oracle_port):
    """8000 keypoints (C4) and wide windows: the candidate arrays outgrow their first allocation and the search is repeated."""
    fr = synth.frame(9, n=8000, w=3840, h=2160)
    pts, desc = synth.local_map_points(9, fr, npts=6000)
    f = gc.make_frame(orbx, fr)
    f.mappoints[:] = -1
    n = orbx.ORBmatcher(0.8, True).SearchByProjection(f, pts, desc, 3.0)
    w, mp = oracle_port.search_local_map(fr, np.full(8000, -1, np.int32), pts, desc, 3.0, 0.8)
    assert n == w and np.array_equal(f.mappoints, mp)
    small = synth.frame(10, n=3000, w=320, h=240)
    pts, desc = synth.local_map_points(10, small, npts=500)
    f = gc.make_frame(orbx, small)
    n = orbx.ORBmatcher(0.8, True).SearchByProjection(f, pts, desc, 40.0)     # ~500 candidates per point
    w, mp = oracle_port.search_local_map(small, np.full(3000, -1, np.int32), pts, desc, 40.0, 0.8)
    assert n == w and np.array_equal(f.mappoints, mp)


def test_contract_errors(orbx):
    fr = synth.frame(3, n=100)
    with pytest.raises(orbx.OrbxError):
        orbx.Frame(fr['kps_un'], fr['desc'], fr['scale_factors'], (0.0, 0.0, 0.0, 480.0))      # empty bounds: division by zero upstream
    f = gc.make_frame(orbx, fr)
    pts, desc = synth.local_map_points(3, fr, npts=10)
    pts['scale_level'][0] = 9
    pts['flags'][0] = 1
    with pytest.raises(orbx.OrbxError):
        orbx.ORBmatcher(0.8).SearchByProjection(f, pts, desc, 3.0)


def test_resident_chain_extract_undistort_frame_search(orbx, oracle_port):
    """Extract -> UndistortKeyPoints -> Frame (grid) -> SearchByProjection with keypoints and descriptors staying on the GPU
    (orbx_extract_batch_device, orbx_undistort_keypoints_device, orbx_frame_assign_device): same frame.mappoints as the oracle fed with
    host copies of the same data."""
    import ctypes as C
    import torch
    from test_oracle_golden import UNDISTORT_CASES
    cam4, dist = UNDISTORT_CASES['tum1']
    cam = cam4 + (40.0, 40.0 / cam4[0])
    img = synth.image(77, 640, 480)
    ex = orbx.ORBextractor(nfeatures=1000)
    d_img = torch.from_numpy(img[None]).cuda()
    d_kps, d_desc, d_n = ex.extract_batch_device(d_img)
    ex.synchronize()
    n = int(d_n[0].item())
    d_un = torch.empty_like(d_kps[0])
    dcoef = np.array(dist, np.float32)
    orbx._check(orbx.lib().orbx_undistort_keypoints_device(C.c_void_p(d_kps[0].data_ptr()), n, C.byref(orbx._Camera(*cam)), dcoef.ctypes.data, len(dcoef),
                                                           C.c_void_p(d_un.data_ptr()), None))
    torch.cuda.synchronize()
    sf = ex.GetScaleFactors()
    bounds = (-12.5, 655.25, -9.0, 491.5)                       # an undistorted image's bounds (src/System.cc:178-195)
    seedfr = synth.frame(0, n=4)
    f = orbx.Frame(seedfr['kps_un'], seedfr['desc'], seedfr['scale_factors'], seedfr['bounds'])
    f.assign_device(d_un, d_desc[0], n, sf, bounds)
    # host copies of the same data for the oracle
    kps_un = d_un[:n].cpu().numpy().view(orbx.KP_DTYPE).reshape(-1)
    kps_h, desc_h = ex.Extract(img)
    assert kps_un.tobytes() == oracle_port.undistort_keypoints(kps_h, cam, dist).tobytes()
    fr = dict(kps_un=kps_un, desc=d_desc[0, :n].cpu().numpy(), uright=None, bounds=bounds, nlevels=len(sf), scale_factors=sf)
    assert np.array_equal(fr['desc'], desc_h)
    pts, pdesc = synth.local_map_points(5, fr, npts=900)
    got = orbx.ORBmatcher(0.8).SearchByProjection(f, pts, pdesc, 3.0)
    want, wmp = oracle_port.search_local_map(fr, np.full(n, -1, np.int32), pts, pdesc, 3.0, 0.8)
    assert got == want and np.array_equal(f.mappoints, wmp) and want > 200


def test_mapping_matchers_degenerate_inputs(orbx, oracle_port):
    """Fuse / SearchBySim3 / SearchForTriangulation on empty and disjoint inputs: same answers as the oracle, no work lost or invented."""
    fr = synth.frame(70, n=500)
    f = gc.make_frame(orbx, fr)
    m = orbx.ORBmatcher(0.6, True)
    _, inv_sig = synth.sigma_tables(fr['scale_factors'])
    S, pts, desc = synth.sim3_points(70, fr, gc.TUM_CAMERA, npts=300, scale=1.0)
    # no points at all, then points that are all flagged off
    bi, bd = m.FuseSearch(f, gc.TUM_CAMERA, (S[0], S[1]), synth.log_scale_factor(), inv_sig, pts[:0], desc[:0], 3.0)
    assert len(bi) == 0 and len(bd) == 0
    off = pts.copy(); off['flags'] = 0
    bi, bd = m.FuseSearch(f, gc.TUM_CAMERA, (S[0], S[1]), synth.log_scale_factor(), inv_sig, off, desc, 3.0)
    assert (bi == -1).all() and (bd == 256).all()
    bi, bd = m.FuseSim3Search(f, gc.TUM_CAMERA, S, synth.log_scale_factor(), off, desc, 4.0)
    assert (bi == -1).all() and (bd == 256).all()
    # a radius so small that no window holds a keypoint
    bi, bd = m.FuseSim3Search(f, gc.TUM_CAMERA, S, synth.log_scale_factor(), pts, desc, 1e-4)
    assert (bi == -1).all()
    # triangulation: disjoint vocabulary nodes, then only_stereo on monocular key frames
    tp = synth.triangulation_pair(3, n=400)
    f1, f2 = gc.make_frame(orbx, tp['f1']), gc.make_frame(orbx, tp['f2'])
    ids2, st2, ix2 = tp['fv2']
    n, m12 = m.SearchForTriangulation(f1, tp['fv1'], tp['has1'], f2, ((ids2 + 1000000).astype(np.uint32), st2, ix2), tp['has2'], tp['F12'], tp['ep2'],
                                      tp['sigma_sq2'], False)
    assert n == 0 and (m12 == -1).all()
    mono = dict(tp); mono['f1'] = dict(tp['f1']); mono['f2'] = dict(tp['f2'])
    mono['f1']['uright'] = np.full(400, -1, np.float32); mono['f2']['uright'] = np.full(400, -1, np.float32)
    g1, g2 = gc.make_frame(orbx, mono['f1']), gc.make_frame(orbx, mono['f2'])
    n, m12 = m.SearchForTriangulation(g1, mono['fv1'], mono['has1'], g2, mono['fv2'], mono['has2'], mono['F12'], mono['ep2'], mono['sigma_sq2'], True)
    wn, wm = oracle_port.search_for_triangulation(mono, True, True)
    assert n == wn == 0 and np.array_equal(m12, wm)
    # every keypoint of both key frames already holds a map point: nothing to triangulate
    full = np.ones(400, np.uint8)
    n, m12 = m.SearchForTriangulation(f1, tp['fv1'], full, f2, tp['fv2'], full, tp['F12'], tp['ep2'], tp['sigma_sq2'], False)
    assert n == 0 and (m12 == -1).all()
    # SearchBySim3 with no eligible map point on either side
    sp = synth.sim3_pair(4, n=300); sp['lsf'] = synth.log_scale_factor()
    s1, s2 = gc.make_frame(orbx, sp['f1']), gc.make_frame(orbx, sp['f2'])
    p1 = sp['pts1'].copy(); p2 = sp['pts2'].copy(); p1['flags'] = 0; p2['flags'] = 0
    n, m12, m1, m2 = orbx.ORBmatcher(0.75, True).SearchBySim3(s1, sp['cam'], sp['pose1'], sp['lsf'], s2, sp['cam'], sp['pose2'], sp['lsf'], sp['S12'], 7.5,
                                                             p1, sp['desc1'], p2, sp['desc2'])
    assert n == 0 and (m12 == -1).all() and (m1 == -1).all() and (m2 == -1).all()
