"""Seeded cases for the guided matchers, shared by the golden generator, the oracle tests and the GPU parity tests.
Every case is (name, kind, inputs); run_oracle / run_product evaluate it and return a dict of arrays that must agree bit for bit."""
import numpy as np

from orb_slam2_refactored_b200 import synth

TUM_CAMERA = (517.306408, 516.469215, 318.643040, 255.313989, 40.0, 40.0 / 517.306408)   # Examples/Monocular/TUM1.yaml + RGB-D bf


def grid_queries(seed, n=400):
    r = np.random.RandomState(seed + 9)
    q = []
    for lo, hi in [(-1, -1), (0, 0), (2, 3), (1, -1), (0, 7), (-1, 2), (5, 4)]:
        for _ in range(n // 7):
            q.append((r.rand() * 760 - 60, r.rand() * 600 - 60, r.rand() * 60 + 0.5, lo, hi))
    q += [(-500.0, 10.0, 5.0, -1, -1), (10.0, 5000.0, 5.0, -1, -1), (320.0, 240.0, 2000.0, -1, -1), (0.0, 0.0, 0.0, -1, -1)]
    return np.array(q, np.float64)


def cases(small=False):
    """small=True: the subset stored as golden vectors."""
    out = []
    seeds = (0, 1) if small else range(6)
    for seed in seeds:
        fr = synth.frame(seed, n=900 if small else 1500, bounds_margin=3.7 if seed % 2 else 0.0, stereo=seed % 3 != 2)
        out.append((f'grid{seed}', 'grid', dict(frame=fr, queries=grid_queries(seed, 140 if small else 400))))
        mp0 = synth.initial_frame_mappoints(seed, len(fr['kps_un']))
        pts, desc = synth.local_map_points(seed, fr, npts=700 if small else 1200)
        out.append((f'local{seed}', 'local', dict(frame=fr, mp=mp0, pts=pts, desc=desc, th=3.0 if seed % 3 else 5.0, nnratio=0.8)))
        cam = synth.KITTI_CAMERA if seed % 2 else TUM_CAMERA
        for dz in ((0.0, 1.0, -1.0) if not small else (0.0, 1.0)):
            cp, lp, lpts, ldesc = synth.last_frame_points(seed, fr, cam, npts=700 if small else 1200, dz=dz)
            for mono in (False, True):
                if small and mono and dz:
                    continue
                out.append((f'last{seed}_dz{dz:+.0f}_m{int(mono)}', 'last',
                            dict(frame=fr, cam=cam, cur_pose=cp, last_pose=lp, mp=mp0, pts=lpts, desc=ldesc, th=15.0 if mono else 7.0,
                                 monocular=mono, check=seed != 3)))
        f1, f2, prev = synth.initialization_pair(seed, n=800 if small else 1500)
        for win in ((100,) if small else (100, 20)):
            out.append((f'init{seed}_w{win}', 'init', dict(f1=f1, f2=f2, prev=prev, window=win, nnratio=0.9, check=seed != 4)))
        kpose, kpts, kdesc = synth.keyframe_points(seed, fr, cam, npts=700 if small else 1200)
        out.append((f'reloc{seed}', 'reloc', dict(frame=fr, cam=cam, pose=kpose, mp=mp0, pts=kpts, desc=kdesc, th=10.0 if seed % 2 else 3.0,
                                                  orb_dist=100 if seed % 2 else 64, check=seed != 1)))
        S, spts, sdesc = synth.sim3_points(seed, fr, cam, npts=700 if small else 1200)
        m0 = np.where(mp0 == -1, -1, -2).astype(np.int32)
        out.append((f'sim3_{seed}', 'sim3', dict(frame=fr, cam=cam, sim3=S, matched=m0, pts=spts, desc=sdesc, th=10 if seed % 2 else 4)))
        f1, fv1, va1, f2, fv2, va2 = synth.bow_pair(seed, n=800 if small else 1500)
        out.append((f'bow{seed}_kf_frame', 'bow', dict(f1=f1, fv1=fv1, valid1=va1, f2=f2, fv2=fv2, valid2=None, nnratio=0.7, check=seed != 5)))
        out.append((f'bow{seed}_kf_kf', 'bow', dict(f1=f1, fv1=fv1, valid1=va1, f2=f2, fv2=fv2, valid2=va2, nnratio=0.8 if seed % 2 else 0.75,
                                                    check=seed != 2)))
        # ---- the matchers of local mapping / loop closing whose per-point search is independent
        Sf, fpts, fdesc = synth.sim3_points(seed + 20, fr, cam, npts=700 if small else 1200, scale=1.0)
        fpts = fpts.copy(); fpts['flags'] = (np.random.RandomState(seed).rand(len(fpts)) < 0.93).astype(np.int32)   # a few null entries
        frf = dict(fr)                                         # most keypoints monocular here, so that both chi-square gates (:934-945) pass and fail
        if fr['uright'] is not None:
            frf['uright'] = np.where(np.random.RandomState(seed + 3).rand(len(fr['uright'])) < 0.7, np.float32(-1), fr['uright']).astype(np.float32)
        out.append((f'fuse{seed}', 'fuse', dict(frame=frf, cam=cam, pose=(Sf[0], Sf[1]), pts=fpts, desc=fdesc, th=3.0 if seed % 2 else 5.0,
                                                state=synth.fuse_map(seed, fr, len(fpts)))))
        out.append((f'fuse_sim3_{seed}', 'fuse_sim3', dict(frame=fr, cam=cam, sim3=S, pts=spts, desc=sdesc, th=4.0 if seed % 2 else 6.0,
                                                          state=synth.fuse_map(seed + 1, fr, len(spts)))))
        sp = synth.sim3_pair(seed, n=700 if small else 1200)
        sp['lsf'] = synth.log_scale_factor()
        r5 = np.random.RandomState(seed + 5)
        for P in (sp['pts1'], sp['pts2']):
            P['flags'] |= (4 * (r5.rand(len(P)) < 0.04)).astype(np.int32)          # present but bad
        sp['pts1']['flags'] |= (2 * (r5.rand(len(sp['pts1'])) < 0.05)).astype(np.int32)   # already in matches12 on entry
        out.append((f'sim3_search{seed}', 'sim3_search', dict(scene=sp, th=7.5 if seed % 2 else 4.0)))
        tp = synth.triangulation_pair(seed, n=800 if small else 1500)
        for only_stereo in ((False,) if small else (False, True)):
            out.append((f'triang{seed}_s{int(only_stereo)}', 'triang', dict(scene=tp, only_stereo=only_stereo, check=seed != 2)))
    if not small:
        # crowded: most points aim at a keypoint some other point wants too, and the alternatives are close: long dependency chains
        fr = synth.frame(40, n=600, w=320, h=240)
        pts, desc = synth.local_map_points(40, fr, npts=3000, dup=0.8, max_flips=40)
        pts['flags'] |= 2
        out.append(('local_crowded', 'local', dict(frame=fr, mp=np.full(600, -1, np.int32), pts=pts, desc=desc, th=6.0, nnratio=0.9)))
        cp, lp, lpts, ldesc = synth.last_frame_points(41, fr, TUM_CAMERA, npts=3000, dup=0.8, max_flips=40)
        out.append(('last_crowded', 'last', dict(frame=fr, cam=TUM_CAMERA, cur_pose=cp, last_pose=lp, mp=np.full(600, -1, np.int32), pts=lpts,
                                                 desc=ldesc, th=15.0, monocular=True, check=True)))
        # ladder: K keypoints on one spot at distances 0, 1, 2, ... from one descriptor, K + 3 identical points: point i must end on
        # keypoint i, which takes one round per rung (the worst case of the fixpoint)
        K = 12
        fr = synth.frame(60, n=400)
        base = fr['desc'][0].copy()
        for j in range(K):
            fr['kps_un']['x'][j], fr['kps_un']['y'][j], fr['kps_un']['octave'][j] = 200.0 + 0.1 * j, 150.0, 0
            bits = np.unpackbits(base)
            bits[:j] ^= 1
            fr['desc'][j] = np.packbits(bits)
        if fr['uright'] is not None:
            fr['uright'][:K] = -1.0
        pts = np.zeros(K + 3, synth.TRACK_POINT_DTYPE)
        pts['proj_x'], pts['proj_y'], pts['proj_xr'], pts['view_cos'], pts['scale_level'], pts['flags'] = 200.5, 150.0, -1.0, 0.9, 0, 3
        out.append(('local_ladder', 'local', dict(frame=fr, mp=np.full(400, -1, np.int32), pts=pts, desc=np.tile(base, (K + 3, 1)), th=3.0,
                                                  nnratio=1.0)))
        # degenerate inputs
        fr = synth.frame(50, n=300)
        pts, desc = synth.local_map_points(50, fr, npts=50)
        pts['flags'] &= ~1
        out.append(('local_all_inactive', 'local', dict(frame=fr, mp=np.full(300, -1, np.int32), pts=pts, desc=desc, th=3.0, nnratio=0.8)))
        out.append(('local_no_points', 'local', dict(frame=fr, mp=synth.initial_frame_mappoints(50, 300), pts=pts[:0], desc=desc[:0], th=3.0,
                                                     nnratio=0.8)))
        one = synth.frame(51, n=1)
        p1, d1 = synth.local_map_points(51, one, npts=5, max_flips=10)
        p1['flags'] = 3
        out.append(('local_one_keypoint', 'local', dict(frame=one, mp=np.full(1, -1, np.int32), pts=p1, desc=d1, th=5.0, nnratio=0.8)))
    return out


def run_oracle(o, kind, c):
    if kind == 'grid':
        res = o.grid_queries(c['frame'], c['queries'])
        return dict(counts=np.array([len(r) for r in res], np.int32), indices=np.concatenate(res + [np.empty(0, np.int32)]).astype(np.int32))
    if kind == 'local':
        n, mp = o.search_local_map(c['frame'], c['mp'], c['pts'], c['desc'], c['th'], c['nnratio'])
        return dict(n=np.int32(n), mp=mp)
    if kind == 'last':
        n, mp = o.search_last_frame(c['frame'], c['cam'], c['cur_pose'], c['last_pose'], c['mp'], c['pts'], c['desc'], c['th'], c['monocular'],
                                    0.9, c['check'])
        return dict(n=np.int32(n), mp=mp)
    if kind == 'reloc':
        n, mp = o.search_keyframe_projection(c['frame'], c['cam'], c['pose'], synth.log_scale_factor(), c['mp'], c['pts'], c['desc'], c['th'],
                                             c['orb_dist'], c['check'])
        return dict(n=np.int32(n), mp=mp)
    if kind == 'sim3':
        n, m = o.search_sim3_projection(c['frame'], c['cam'], c['sim3'], synth.log_scale_factor(), c['matched'], c['pts'], c['desc'], c['th'])
        return dict(n=np.int32(n), matched=m)
    if kind == 'bow':
        n, m2 = o.search_by_bow(c['f1'], c['fv1'], c['valid1'], c['f2'], c['fv2'], c['valid2'], c['nnratio'], c['check'])
        return dict(n=np.int32(n), m2=m2)
    if kind == 'fuse':
        _, inv_sig = synth.sigma_tables(c['frame']['scale_factors'])
        n, kf_mp, nobs, bad, in_kf, log = o.fuse(c['frame'], c['cam'], c['pose'], synth.log_scale_factor(), inv_sig, c['pts'], c['desc'], c['th'], c['state'])
        return dict(n=np.int32(n), kf_mp=kf_mp, nobs=nobs, bad=bad, in_kf=in_kf, log=log)
    if kind == 'fuse_sim3':
        n, kf_mp, nobs, bad, rep, log = o.fuse_sim3(c['frame'], c['cam'], c['sim3'], synth.log_scale_factor(), c['pts'], c['desc'], c['th'], c['state'])
        return dict(n=np.int32(n), kf_mp=kf_mp, nobs=nobs, bad=bad, replace=rep, log=log)
    if kind == 'sim3_search':
        n, m12 = o.search_by_sim3(c['scene'], c['th'])
        return dict(n=np.int32(n), m12=m12)
    if kind == 'triang':
        n, m12 = o.search_for_triangulation(c['scene'], c['only_stereo'], c['check'])
        return dict(n=np.int32(n), m12=m12)
    n, m12, prev = o.search_for_initialization(c['f1'], c['f2'], c['prev'], c['window'], c['nnratio'], c['check'])
    return dict(n=np.int32(n), m12=m12, prev=prev)


def make_frame(api, fr, device=0):
    return api.Frame(fr['kps_un'], fr['desc'], fr['scale_factors'], fr['bounds'], fr['uright'], device=device)


def run_product(api, kind, c, device=0):
    """The same through the C ABI (orb_slam2_refactored_b200.api). Also returns the number of rounds under key '_rounds'."""
    if kind == 'grid':
        f = make_frame(api, c['frame'], device)
        res = f.GetFeaturesInAreaBatch(c['queries'])
        return dict(counts=np.array([len(r) for r in res], np.int32), indices=np.concatenate(res + [np.empty(0, np.int32)]).astype(np.int32))
    if kind == 'local':
        f = make_frame(api, c['frame'], device)
        f.mappoints[:] = c['mp']
        n = api.ORBmatcher(c['nnratio'], True, device).SearchByProjection(f, c['pts'], c['desc'], c['th'])
        return dict(n=np.int32(n), mp=f.mappoints.copy(), _rounds=f.last_rounds())
    if kind == 'last':
        f = make_frame(api, c['frame'], device)
        f.mappoints[:] = c['mp']
        n = api.ORBmatcher(0.9, c['check'], device).SearchByProjectionLastFrame(f, c['cam'], c['cur_pose'], c['last_pose'], c['pts'], c['desc'],
                                                                                c['th'], c['monocular'])
        return dict(n=np.int32(n), mp=f.mappoints.copy(), _rounds=f.last_rounds())
    if kind == 'reloc':
        f = make_frame(api, c['frame'], device)
        f.mappoints[:] = c['mp']
        n = api.ORBmatcher(0.9, c['check'], device).SearchByProjectionKeyFrame(f, c['cam'], c['pose'], synth.log_scale_factor(), c['pts'], c['desc'],
                                                                               c['th'], c['orb_dist'])
        return dict(n=np.int32(n), mp=f.mappoints.copy(), _rounds=f.last_rounds())
    if kind == 'sim3':
        f = make_frame(api, c['frame'], device)
        f.mappoints[:] = c['matched']
        n = api.ORBmatcher(0.75, True, device).SearchByProjectionSim3(f, c['cam'], c['sim3'], synth.log_scale_factor(), c['pts'], c['desc'], c['th'])
        return dict(n=np.int32(n), matched=f.mappoints.copy(), _rounds=f.last_rounds())
    if kind in ('fuse', 'fuse_sim3'):
        # the library does the search of every point; :876 / :1002-1003 and the mutation of :956-976 / :1069-1084 are replayed here, in order,
        # on the same one-key-frame map model the oracle uses (oracle/ref_guided_decl.h)
        f = make_frame(api, c['frame'], device)
        st = {k: v.copy() for k, v in c['state'].items()}
        kf_mp, nobs, bad = st['kf_mp'], st['nobs'], st['bad']
        npts = len(c['pts'])
        where = np.full(npts, -1, np.int64)
        held0 = np.flatnonzero(kf_mp >= 0)
        where[kf_mp[held0]] = held0
        log = []
        ur = c['frame']['uright']

        def add_observation(p, idx):
            log.extend((2, p, idx)); where[p] = idx
            nobs[p] += 2 if (ur is not None and ur[idx] >= 0) else 1
            kf_mp[idx] = p

        def replace(p, other):
            log.extend((1, p, other))
            if p == other:
                return
            bad[p] = 1
            if where[p] >= 0:
                if where[other] < 0:
                    kf_mp[where[p]] = other; where[other] = where[p]; nobs[other] += nobs[p]
                else:
                    kf_mp[where[p]] = -1
        m = api.ORBmatcher(0.6, True, device)
        n = 0
        if kind == 'fuse':
            _, inv_sig = synth.sigma_tables(c['frame']['scale_factors'])
            pts = c['pts'].copy()
            bi, bd = m.FuseSearch(f, c['cam'], c['pose'], synth.log_scale_factor(), inv_sig, pts, c['desc'], c['th'])
            for i in range(npts):
                if not (pts['flags'][i] & 1) or bad[i] or where[i] >= 0 or bd[i] > api.TH_LOW:
                    continue
                held = kf_mp[bi[i]]
                if held >= 0:
                    if not bad[held]:
                        if nobs[held] > nobs[i]:
                            replace(i, held)
                        else:
                            replace(held, i)
                else:
                    add_observation(i, bi[i])
                n += 1
            return dict(n=np.int32(n), kf_mp=kf_mp, nobs=nobs, bad=bad, in_kf=(where >= 0).astype(np.uint8), log=np.array(log, np.int32), _rounds=1)
        found = np.zeros(npts, bool)
        found[kf_mp[held0][bad[kf_mp[held0]] == 0]] = True
        pts = c['pts'].copy()
        pts['flags'] = 1                                       # search everything; bad / alreadyFound are tested at replay time
        bi, bd = m.FuseSim3Search(f, c['cam'], c['sim3'], synth.log_scale_factor(), pts, c['desc'], c['th'])
        rep = np.full(npts, -1, np.int32)
        for i in range(npts):
            if bad[i] or found[i] or bd[i] > api.TH_LOW:
                continue
            held = kf_mp[bi[i]]
            if held >= 0:
                if not bad[held]:
                    rep[i] = held
            else:
                add_observation(i, bi[i])
            n += 1
        return dict(n=np.int32(n), kf_mp=kf_mp, nobs=nobs, bad=bad, replace=rep, log=np.array(log, np.int32), _rounds=1)
    if kind == 'sim3_search':
        s = c['scene']
        f1, f2 = make_frame(api, s['f1'], device), make_frame(api, s['f2'], device)
        N1, N2 = f1.N, f2.N
        p1, p2 = s['pts1'].copy(), s['pts2'].copy()
        am1 = (p1['flags'] & 2) != 0
        am2 = np.zeros(N2, bool)
        am2[(np.flatnonzero(am1) * 7) % N2] = True
        p1['flags'] = (((p1['flags'] & 5) != 0) & ((p1['flags'] & 4) == 0) & ~am1).astype(np.int32)     # present, not bad, not already matched
        p2['flags'] = (((p2['flags'] & 5) != 0) & ((p2['flags'] & 4) == 0) & ~am2).astype(np.int32)
        n, m12, m1, m2 = api.ORBmatcher(0.75, True, device).SearchBySim3(f1, s['cam'], s['pose1'], s['lsf'], f2, s['cam'], s['pose2'], s['lsf'], s['S12'],
                                                                         c['th'], p1, s['desc1'], p2, s['desc2'])
        out = np.where(m12 >= 0, m12, np.where(am1, (np.arange(N1) * 7) % N2, -1)).astype(np.int32)      # matches12 keeps its entries from before
        return dict(n=np.int32(n), m12=out, _rounds=1)
    if kind == 'triang':
        s = c['scene']
        f1, f2 = make_frame(api, s['f1'], device), make_frame(api, s['f2'], device)
        n, m12 = api.ORBmatcher(0.6, c['check'], device).SearchForTriangulation(f1, s['fv1'], s['has1'], f2, s['fv2'], s['has2'], s['F12'], s['ep2'],
                                                                                s['sigma_sq2'], c['only_stereo'])
        return dict(n=np.int32(n), m12=m12.copy(), _rounds=1)
    f1, f2 = make_frame(api, c['f1'], device), make_frame(api, c['f2'], device)
    if kind == 'bow':
        n, m2 = api.ORBmatcher(c['nnratio'], c['check'], device).SearchByBoW(f1, c['fv1'], c['valid1'], f2, c['fv2'], c['valid2'])
        return dict(n=np.int32(n), m2=m2.copy(), _rounds=f2.last_rounds())
    prev = c['prev'].copy()
    n, m12 = api.ORBmatcher(c['nnratio'], c['check'], device).SearchForInitialization(f1, f2, prev, c['window'])
    return dict(n=np.int32(n), m12=m12.copy(), prev=prev, _rounds=f2.last_rounds())
