"""Synthetic inputs for tests and bench.py (numpy only, deterministic, no cv2 needed at run time).

The image recipe follows SURVEY.md App. C in spirit: band-limited noise with a horizontal contrast ramp,
so that one frame holds strong corners (iniThFAST fires), weak corners (only the minThFAST retry fires,
src/ORBextractor.cc:526-530) and flat cells (nothing fires), and every pyramid level reaches its quota.
"""
import numpy as np


def _gauss_sep(x, sigma):
    r = int(4 * sigma + 0.5)
    t = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (t / sigma) ** 2)
    k /= k.sum()
    p = np.pad(x, ((0, 0), (r, r)), mode='reflect')
    y = sum(k[i] * p[:, i:i + x.shape[1]] for i in range(2 * r + 1))
    p = np.pad(y, ((r, r), (0, 0)), mode='reflect')
    return sum(k[i] * p[i:i + x.shape[0], :] for i in range(2 * r + 1))


def image(seed, w, h):
    """uint8 (h, w) frame."""
    r = np.random.RandomState(seed)
    x = r.randint(0, 256, (h, w)).astype(np.float64)
    y = _gauss_sep(x, 2.0)
    y = (y - y.mean()) / y.std()
    g = np.linspace(0.05, 1.0, w)[None, :]
    return np.clip(y * 28.0 * g + 128.0, 0, 255).astype(np.uint8)


def stereo_pair(seed, w, h, disparity=12, noise=3):
    """Left frame and a right frame = left shifted by `disparity` px plus small independent noise (an exact shift
    makes most SADs zero, the median zero, and ComputeStereoMatches' outlier cut (src/ORBmatcher.cc:231-246) then
    discards every match — a degenerate case that has its own test)."""
    left = image(seed, w, h)
    fill = image(seed + 1000000, w, h)
    right = np.empty_like(left)
    right[:, :w - disparity] = left[:, disparity:]
    right[:, w - disparity:] = fill[:, w - disparity:]
    if noise:
        r = np.random.RandomState(seed + 2000000)
        right = np.clip(right.astype(np.int16) + r.randint(-noise, noise + 1, right.shape), 0, 255).astype(np.uint8)
    return left, right


def descriptors(seed, n):
    """n x 32 uniform random bytes."""
    return np.random.RandomState(seed).randint(0, 256, (n, 32)).astype(np.uint8)


def planted_descriptors(seed, nq, nt, max_flips=60, dup_every=7):
    """Query/train sets where train[perm[i]] is query[i] with 0..max_flips random bit flips (uniform random
    256-bit vectors sit at distance 128 +- 8, so only planted pairs exercise TH_LOW / ratio acceptance), and
    every dup_every-th planted row is duplicated at a second, higher index to exercise lowest-index ties."""
    r = np.random.RandomState(seed)
    q = r.randint(0, 256, (nq, 32)).astype(np.uint8)
    t = r.randint(0, 256, (nt, 32)).astype(np.uint8)
    m = min(nq, nt // 2)
    perm = r.permutation(nt)[:2 * m]
    for i in range(m):
        bits = np.unpackbits(q[i])
        k = r.randint(0, max_flips + 1)
        flip = r.choice(256, k, replace=False)
        bits[flip] ^= 1
        t[perm[i]] = np.packbits(bits)
        if i % dup_every == 0:
            t[perm[m + i]] = t[perm[i]]
    return q, t


# camera blocks of the reference's example configs (fx, fy, cx, cy, bf, baseline = bf / fx, src/System.cc:47-57)
KITTI_CAMERA = (718.856, 718.856, 607.1928, 185.2157, 386.1448, 386.1448 / 718.856)       # Examples/Stereo/KITTI00-02.yaml
EUROC_CAMERA = (435.2046959714599, 435.2046959714599, 367.4517211914062, 252.2008514404297, 47.90639384423901,
                47.90639384423901 / 435.2046959714599)                                      # Examples/Stereo/EuRoC.yaml

CONFIGS = {
    'C1': dict(w=640, h=480, nfeatures=1000),
    'C2': dict(w=1241, h=376, nfeatures=2000, camera=KITTI_CAMERA),
    'C3': dict(w=752, h=480, nfeatures=1200, camera=EUROC_CAMERA),
    'C4': dict(w=3840, h=2160, nfeatures=8000),
}


# ---- guided-matcher scenes: a frame as Tracking sees it plus the map points projected into it ----
KP_DTYPE = np.dtype([('x', '<f4'), ('y', '<f4'), ('size', '<f4'), ('angle', '<f4'), ('response', '<f4'), ('octave', '<i4'),
                     ('class_id', '<i4')])
TRACK_POINT_DTYPE = np.dtype([('proj_x', '<f4'), ('proj_y', '<f4'), ('proj_xr', '<f4'), ('view_cos', '<f4'), ('scale_level', '<i4'),
                              ('flags', '<i4')])
LAST_POINT_DTYPE = np.dtype([('xw', '<f4', (3,)), ('octave', '<i4'), ('angle', '<f4'), ('flags', '<i4')])


def scale_factors(nlevels=8, factor=1.2):
    """float32 cumulative products, as ORBextractor::Init builds them (src/ORBextractor.cc:728-737)."""
    sf = np.ones(nlevels, np.float32)
    for i in range(1, nlevels):
        sf[i] = np.float32(sf[i - 1] * np.float32(factor))
    return sf


def _flip_bits(r, d, k):
    bits = np.unpackbits(d)
    bits[r.choice(256, k, replace=False)] ^= 1
    return np.packbits(bits)


def frame(seed, n=1500, w=640, h=480, nlevels=8, stereo=True, level0_share=None, bounds_margin=0.0):
    """A synthetic Frame: keypoints on level-integer positions times the level scale (as Extract emits them), octaves drawn like the
    per-level quotas, uniform angles, random descriptors, uright for ~70 % of the keypoints when stereo."""
    r = np.random.RandomState(seed)
    sf = scale_factors(nlevels)
    p = (1.0 / 1.2) ** np.arange(nlevels)
    if level0_share is not None:
        p[0] = level0_share * p[1:].sum() / (1.0 - level0_share)
    octv = r.choice(nlevels, n, p=p / p.sum()).astype(np.int32)
    kps = np.zeros(n, KP_DTYPE)
    lw = np.maximum((w / sf[octv]).astype(np.int32) - 32, 1)
    lh = np.maximum((h / sf[octv]).astype(np.int32) - 32, 1)
    kps['x'] = ((16 + (r.rand(n) * lw).astype(np.int32)).astype(np.float32) * sf[octv]).astype(np.float32)
    kps['y'] = ((16 + (r.rand(n) * lh).astype(np.int32)).astype(np.float32) * sf[octv]).astype(np.float32)
    kps['size'] = np.float32(31.0) * sf[octv]
    kps['angle'] = (r.rand(n) * 360.0).astype(np.float32)
    kps['response'] = r.randint(8, 120, n).astype(np.float32)
    kps['octave'] = octv
    kps['class_id'] = -1
    ur = None
    if stereo:
        ur = np.where(r.rand(n) < 0.7, kps['x'] - (1.0 + 60.0 * r.rand(n)).astype(np.float32), np.float32(-1.0)).astype(np.float32)
    m = np.float32(bounds_margin)
    return dict(kps_un=kps, desc=r.randint(0, 256, (n, 32)).astype(np.uint8), uright=ur,
                bounds=(np.float32(-m), np.float32(w + m), np.float32(-m * 0.5), np.float32(h + m * 0.5)), nlevels=nlevels, scale_factors=sf)


def initial_frame_mappoints(seed, n, p_seen=0.05, p_unseen=0.03):
    """frame.mappoints on entry: -1 null, -2 a map point with observations, -3 one without (see include/orbx.h)."""
    r = np.random.RandomState(seed + 77)
    u = r.rand(n)
    return np.where(u < p_seen, -2, np.where(u < p_seen + p_unseen, -3, -1)).astype(np.int32)


def local_map_points(seed, fr, npts=1200, dup=0.2, max_flips=70):
    """Map points in view of `fr` for SearchByProjection(Frame&, mappoints, th): each aims at a keypoint (20 % at one that another point
    already aims at, so the greedy 'already matched' state matters), with a jittered projection and a descriptor a few bits away."""
    r = np.random.RandomState(seed + 1)
    kps, n = fr['kps_un'], len(fr['kps_un'])
    sf = fr['scale_factors']
    tgt = r.randint(0, n, npts)
    d = r.rand(npts) < dup
    tgt[d] = tgt[r.randint(0, npts, d.sum())]
    pts = np.zeros(npts, TRACK_POINT_DTYPE)
    octv = kps['octave'][tgt]
    pts['scale_level'] = np.clip(octv + r.choice([0, 0, 1, 1, 2, -1], npts), 0, fr['nlevels'] - 1)
    jit = (r.randn(npts, 2) * 2.0).astype(np.float32) * sf[pts['scale_level']][:, None]
    pts['proj_x'] = kps['x'][tgt] + jit[:, 0]
    pts['proj_y'] = kps['y'][tgt] + jit[:, 1]
    ur = fr['uright'][tgt] if fr['uright'] is not None else np.full(npts, -1.0, np.float32)
    pts['proj_xr'] = np.where(r.rand(npts) < 0.85, ur + jit[:, 0], ur + (r.randn(npts) * 40).astype(np.float32)).astype(np.float32)
    pts['view_cos'] = (0.99 + 0.01 * r.rand(npts)).astype(np.float32)
    pts['flags'] = (r.rand(npts) < 0.92).astype(np.int32) | ((r.rand(npts) < 0.8).astype(np.int32) << 1)
    desc = np.stack([_flip_bits(r, fr['desc'][t], r.randint(0, max_flips + 1)) for t in tgt])
    return pts, desc


def last_frame_points(seed, fr, cam, npts=1200, dz=0.0, dup=0.2, max_flips=70):
    """The last frame's tracked keypoints for SearchByProjection(currFrame, lastFrame, th, monocular): world points that project near
    keypoints of `fr` under cur_pose. dz moves the last camera along z so that the forward / backward branches trigger
    (src/ORBmatcher.cc:1286-1288). Returns (cur_pose, last_pose, pts, desc); poses are (R 3x3, t 3) float32."""
    r = np.random.RandomState(seed + 2)
    kps, n = fr['kps_un'], len(fr['kps_un'])
    fx, fy, cx, cy = cam[:4]
    a = 0.02 * r.randn(3)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    U, _, Vt = np.linalg.svd(np.eye(3) + K)
    R = (U @ Vt).astype(np.float32)
    t = (0.1 * r.randn(3)).astype(np.float32)
    last_R = np.eye(3, dtype=np.float32)
    last_t = np.array([0.01, -0.02, dz], np.float32)
    tgt = r.randint(0, n, npts)
    d = r.rand(npts) < dup
    tgt[d] = tgt[r.randint(0, npts, d.sum())]
    z = 2.0 + 28.0 * r.rand(npts)
    z[r.rand(npts) < 0.03] *= -1.0
    octv = np.clip(kps['octave'][tgt] + r.choice([0, 0, 0, 1, -1], npts), 0, fr['nlevels'] - 1).astype(np.int32)
    jit = r.randn(npts, 2) * 3.0 * fr['scale_factors'][octv][:, None]
    u = kps['x'][tgt] + jit[:, 0]
    v = kps['y'][tgt] + jit[:, 1]
    far = r.rand(npts) < 0.03
    u[far] += 2000.0
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    Xw = (Xc - t.astype(np.float64)) @ R.astype(np.float64)      # R^T (Xc - t)
    pts = np.zeros(npts, LAST_POINT_DTYPE)
    pts['xw'] = Xw.astype(np.float32)
    pts['octave'] = octv
    rot = 25.0
    ang = kps['angle'][tgt] + rot + 4.0 * r.randn(npts)
    wild = r.rand(npts) < 0.15
    ang[wild] = 360.0 * r.rand(wild.sum())
    pts['angle'] = np.mod(ang, 360.0).astype(np.float32)
    pts['flags'] = (r.rand(npts) < 0.9).astype(np.int32) | ((r.rand(npts) < 0.8).astype(np.int32) << 1)
    desc = np.stack([_flip_bits(r, fr['desc'][t], r.randint(0, max_flips + 1)) for t in tgt])
    return (R, t), (last_R, last_t), pts, desc


def initialization_pair(seed, n=1500, w=640, h=480, max_flips=60):
    """Two monocular frames for SearchForInitialization: frame 2 holds frame 1's keypoints moved by a few pixels (descriptors a few bits
    away, angles turned by a common rotation) in shuffled order plus unrelated ones; level 0 carries most keypoints."""
    r = np.random.RandomState(seed + 3)
    f1 = frame(seed, n, w, h, stereo=False, level0_share=0.6)
    f2 = frame(seed + 500, n, w, h, stereo=False, level0_share=0.6)
    m = int(0.7 * n)
    src = r.permutation(n)[:m]
    dst = r.permutation(n)[:m]
    k1, k2 = f1['kps_un'], f2['kps_un']
    k2['x'][dst] = k1['x'][src] + (r.randn(m) * 6.0).astype(np.float32)
    k2['y'][dst] = k1['y'][src] + (r.randn(m) * 6.0).astype(np.float32)
    k2['octave'][dst] = k1['octave'][src]
    ang = k1['angle'][src] + 40.0 + 5.0 * r.randn(m)
    wild = r.rand(m) < 0.15
    ang[wild] = 360.0 * r.rand(wild.sum())
    k2['angle'][dst] = np.mod(ang, 360.0).astype(np.float32)
    for s, d in zip(src, dst):
        f2['desc'][d] = _flip_bits(r, f1['desc'][s], r.randint(0, max_flips + 1))
    # near-duplicates in frame 2 so that two frame-1 keypoints compete for one frame-2 keypoint and matches get revoked (:668-672)
    for s in src[: m // 6]:
        j = r.randint(0, n)
        k1['x'][j], k1['y'][j], k1['octave'][j] = k1['x'][s] + np.float32(1.5), k1['y'][s] - np.float32(1.0), k1['octave'][s]
        f1['desc'][j] = _flip_bits(r, f1['desc'][s], r.randint(0, 12))
    prev = np.stack([k1['x'], k1['y']], 1).astype(np.float32)
    return f1, f2, prev


def rectification_maps(seed, w, h, strength=1.0):
    """Float32 (map_x, map_y) of an undistort + rectify warp like the ones cv::initUndistortRectifyMap hands to cv::remap in
    Examples/Stereo/stereo_euroc.cc:88-101: pinhole + radial/tangential distortion + a small rotation. Near the borders the maps leave
    the source image, which exercises the constant border."""
    r = np.random.RandomState(seed + 31)
    fx = fy = 0.61 * w
    cx, cy = 0.49 * w + r.randn(), 0.52 * h + r.randn()
    k1, k2, p1, p2 = -0.28 * strength, 0.07 * strength, 2e-4 * strength, 2e-5 * strength
    a = 0.01 * strength * r.randn(3)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    U, _, Vt = np.linalg.svd(np.eye(3) + K)
    R = U @ Vt
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    X = np.stack([(u - 0.5 * w) / fx, (v - 0.5 * h) / fy, np.ones_like(u)], -1) @ R      # rays of the rectified camera in the old one
    x, y = X[..., 0] / X[..., 2], X[..., 1] / X[..., 2]
    r2 = x * x + y * y
    d = 1 + k1 * r2 + k2 * r2 * r2
    xd = x * d + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * d + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return (fx * xd + cx).astype(np.float32), (fy * yd + cy).astype(np.float32)


def feature_vector(node_of):
    """DBoW2::FeatureVector of a frame as CSR (node_ids ascending, start, indices): node_of[i] = vocabulary node of keypoint i;
    indices inside a node ascend, as FeatureVector::addFeature fills them when transform() walks the descriptors in order."""
    node_of = np.asarray(node_of, np.uint32)
    ids = np.unique(node_of)
    order = np.argsort(node_of, kind='stable')
    counts = np.searchsorted(node_of[order], ids, side='right')
    start = np.concatenate([[0], counts]).astype(np.int32)
    return ids.astype(np.uint32), start, order.astype(np.uint32)


def bow_pair(seed, n=1500, w=640, h=480, nodes=90, max_flips=60):
    """Two frames with feature vectors for SearchByBoW: 70 % of frame 2's keypoints are moved copies of frame 1's (descriptor a few
    bits away, same vocabulary node with probability 0.9, angle turned by a common rotation); near-duplicates inside a node make
    several keypoints of frame 1 compete for one keypoint of frame 2. Returns f1, fv1, valid1, f2, fv2, valid2."""
    r = np.random.RandomState(seed + 4)
    f1 = frame(seed, n, w, h, stereo=False)
    f2 = frame(seed + 700, n, w, h, stereo=False)
    node_ids = np.sort(r.choice(100000, nodes, replace=False)).astype(np.uint32)
    n1 = node_ids[r.randint(0, nodes, n)]
    n2 = node_ids[r.randint(0, nodes, n)]
    m = int(0.7 * n)
    src = r.permutation(n)[:m]
    dst = r.permutation(n)[:m]
    k1, k2 = f1['kps_un'], f2['kps_un']
    ang = k1['angle'][src] + 30.0 + 5.0 * r.randn(m)
    wild = r.rand(m) < 0.15
    ang[wild] = 360.0 * r.rand(wild.sum())
    k2['angle'][dst] = np.mod(ang, 360.0).astype(np.float32)
    for s, d in zip(src, dst):
        f2['desc'][d] = _flip_bits(r, f1['desc'][s], r.randint(0, max_flips + 1))
    same = r.rand(m) < 0.9
    n2[dst[same]] = n1[src[same]]
    for s in src[: m // 5]:      # a second keypoint of frame 1, in the same node, that looks like s
        j = r.randint(0, n)
        f1['desc'][j] = _flip_bits(r, f1['desc'][s], r.randint(0, 10))
        n1[j] = n1[s]
    valid1 = (r.rand(n) < 0.85).astype(np.uint8)
    valid2 = (r.rand(n) < 0.85).astype(np.uint8)
    return f1, feature_vector(n1), valid1, f2, feature_vector(n2), valid2


KF_POINT_DTYPE = np.dtype([('xw', '<f4', (3,)), ('min_distance', '<f4'), ('max_distance', '<f4'), ('angle', '<f4'), ('flags', '<i4')])


def log_scale_factor(factor=1.2):
    """pyramid.logScaleFactor as src/System.cc:143 computes it: log (double) of the float scale factor, stored in a float."""
    return np.float32(np.log(np.float64(np.float32(factor))))


def keyframe_points(seed, fr, cam, npts=1200, dup=0.2, max_flips=70):
    """Map points of a key frame for the relocalisation search SearchByProjection(Frame&, KeyFrame*, alreadyFound, th, ORBdist): world
    points that project near keypoints of `fr` under the returned pose, with the distance-invariance range of a point first seen at
    the keypoint's octave (some out of range, some behind the camera, some already found). Returns (pose, pts, desc)."""
    r = np.random.RandomState(seed + 5)
    (R, t), _, lp, desc = last_frame_points(seed + 50, fr, cam, npts=npts, dup=dup, max_flips=max_flips)
    pts = np.zeros(npts, KF_POINT_DTYPE)
    pts['xw'] = lp['xw']
    Ow = -(R.astype(np.float64).T @ t.astype(np.float64))
    dist = np.linalg.norm(lp['xw'].astype(np.float64) - Ow, axis=1)
    lvl = np.clip(lp['octave'] + r.choice([0, 0, 1, -1], npts), 0, fr['nlevels'] - 1)
    sf = fr['scale_factors'].astype(np.float64)
    pts['max_distance'] = (dist * sf[lvl] * (1.0 + 0.05 * r.randn(npts))).astype(np.float32)        # dist * levelScaleFactor (MapPoint::UpdateNormalAndDepth)
    pts['min_distance'] = (pts['max_distance'] / sf[fr['nlevels'] - 1]).astype(np.float32)
    far = r.rand(npts) < 0.05
    pts['max_distance'][far] *= np.float32(0.3)
    pts['angle'] = lp['angle']
    pts['flags'] = (r.rand(npts) < 0.85).astype(np.int32)
    return (R, t), pts, desc


SIM3_POINT_DTYPE = np.dtype([('xw', '<f4', (3,)), ('normal', '<f4', (3,)), ('min_distance', '<f4'), ('max_distance', '<f4'), ('flags', '<i4')])


def sim3_points(seed, fr, cam, npts=1200, scale=1.07):
    """Candidate map points for the loop-closing search SearchByProjection(keyframe, Scw, mappoints, matched, th): the points of
    keyframe_points() expressed in a world that is `scale` times larger, so that Scw = (R, scale * t, scale) maps them onto the key
    frame's keypoints; normals point roughly at the camera (some beyond 60 degrees). Returns (sim3, pts, desc)."""
    r = np.random.RandomState(seed + 6)
    (R, t), kp, desc = keyframe_points(seed, fr, cam, npts=npts)
    s = np.float32(scale)
    pts = np.zeros(npts, SIM3_POINT_DTYPE)
    pts['xw'] = kp['xw']
    pts['min_distance'], pts['max_distance'] = kp['min_distance'], kp['max_distance']
    pts['flags'] = kp['flags']
    Ow = -(R.astype(np.float64).T @ t.astype(np.float64))
    to_cam = pts['xw'].astype(np.float64) - Ow
    nrm = to_cam / np.linalg.norm(to_cam, axis=1, keepdims=True)
    tilt = r.randn(npts, 3) * 0.5
    tilt[r.rand(npts) < 0.15] *= 6.0
    nrm = nrm + tilt
    pts['normal'] = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    return (R, (t * s).astype(np.float32), s), pts, desc


def vocabulary(seed, k=10, L=6, prune=0.0, stop=0.1, min_leaf_level=2, hier_flips=40):
    """A synthetic ORB vocabulary tree in the node order DBoW2's HKmeansStep produces (the children of a node get consecutive ids, then each
    child is expanded depth-first), which is also the line order of the text file loadFromTextFile reads. prune > 0 drops children and
    ends branches early (never above min_leaf_level, where the reference's node id would be unset). A child's descriptor is its
    parent's with hier_flips random bit flips, so that the walk down the tree is decided by close calls and ties. Weights are small
    non-negative integers (this fork reads the weight with atoi), a share `stop` of the words has weight 0 (stopped).
    Returns dict(k, L, scoring, weighting, parent, is_leaf, desc, weights, text_weights)."""
    r = np.random.RandomState(seed)
    parent, leaf, level = [], [], []
    stack = [(0, 0)]
    nxt = 1
    while stack:
        node, lvl = stack.pop()
        nchild = k if (prune == 0.0 or r.rand() > prune) else max(1, int(r.randint(1, k + 1)))
        ids = list(range(nxt, nxt + nchild))
        nxt += nchild
        expand = []
        for c in ids:
            parent.append(node); level.append(lvl + 1)
            is_leaf = lvl + 1 >= L or (prune > 0.0 and lvl + 1 >= min_leaf_level and r.rand() < prune)
            leaf.append(is_leaf)
            if not is_leaf:
                expand.append((c, lvl + 1))
        stack.extend(reversed(expand))
    n = len(parent)
    parent = np.asarray(parent, np.int32); leaf = np.asarray(leaf, np.uint8)
    bits = np.zeros((n + 1, 256), np.uint8)
    bits[0] = r.randint(0, 2, 256)
    flips = r.randint(0, 256, (n, hier_flips))
    level = np.asarray(level)
    for lvl in range(1, L + 1):                               # parents before children, one level at a time
        idx = np.flatnonzero(level == lvl)
        if len(idx) == 0:
            break
        mask = np.zeros((len(idx), 256), np.uint8)
        mask[np.arange(len(idx))[:, None], flips[idx]] = 1    # a position drawn twice flips once
        bits[idx + 1] = bits[parent[idx]] ^ mask
    desc = np.packbits(bits[1:], axis=1, bitorder='little')
    w = r.randint(1, 10, n).astype(np.float64)
    w[r.rand(n) < stop] = 0.0
    frac = r.randint(0, 100, n)
    text = [f'{int(w[i])}.{frac[i]:02d}' if frac[i] % 2 else f'{int(w[i])}' for i in range(n)]
    return dict(k=k, L=L, scoring=0, weighting=0, parent=parent, is_leaf=leaf, desc=desc, weights=w, text_weights=text)


def scatter_vocabulary(voc, seed):
    """The same tree with the nodes of every level in random order (parents still precede their children, as loadFromTextFile needs):
    siblings no longer have consecutive ids, children lists keep their relative order by new id."""
    r = np.random.RandomState(seed)
    n = len(voc['parent'])
    level = np.zeros(n + 1, np.int64)
    for i in range(n):
        level[i + 1] = level[voc['parent'][i]] + 1
    order = np.lexsort((r.rand(n), level[1:]))             # old index (0-based) of the node that becomes new node k + 1
    new_id = np.zeros(n + 1, np.int64)
    new_id[order + 1] = np.arange(1, n + 1)
    out = dict(voc)
    out['parent'] = new_id[voc['parent'][order]].astype(np.int32)
    out['is_leaf'] = voc['is_leaf'][order]
    out['desc'] = voc['desc'][order]
    out['weights'] = voc['weights'][order]
    out['text_weights'] = [voc['text_weights'][i] for i in order]
    return out


def write_vocabulary_text(voc, path):
    """ORBvoc.txt layout: header `k L scoring weighting`, then one line per node: parent id, leaf flag, 32 descriptor bytes, weight."""
    with open(path, 'w') as f:
        f.write(f"{voc['k']} {voc['L']} {voc['scoring']} {voc['weighting']}\n")
        d = voc['desc']
        for i in range(len(voc['parent'])):
            f.write(f"{voc['parent'][i]} {int(voc['is_leaf'][i])} " + ' '.join(map(str, d[i])) + f" {voc['text_weights'][i]} \n")


def vocabulary_features(seed, voc, n=1500, exact=0.1, near=0.5, flips=30):
    """Descriptors to push through a vocabulary: random ones, copies of word descriptors (distance 0) and noisy copies."""
    r = np.random.RandomState(seed)
    d = r.randint(0, 256, (n, 32)).astype(np.uint8)
    words = np.flatnonzero(voc['is_leaf'])
    for i in range(n):
        u = r.rand()
        if u < exact + near:
            d[i] = voc['desc'][words[r.randint(len(words))]]
            if u >= exact:
                d[i] = _flip_bits(r, d[i], r.randint(1, flips + 1))
    return d


def sigma_tables(scale_factors):
    """pyramid.sigmaSq / invSigmaSq as ORBextractor::Init computes them (src/ORBextractor.cc:733-736): float products."""
    sf = np.asarray(scale_factors, np.float32)
    sig = (sf * sf).astype(np.float32)
    return sig, (np.float32(1.0) / sig).astype(np.float32)


def fuse_map(seed, fr, npts, p_occupied=0.5):
    """The mutable state Fuse (src/ORBmatcher.cc:868-974) reads between points: which keypoints of the key frame hold a map point (an index
    into the point list, so that Replace has both parties), every point's observation count and bad flag. Points that sit in the key
    frame are `in_kf`. Returns dict(kf_mp [N], nobs [npts], bad [npts], in_kf [npts])."""
    r = np.random.RandomState(seed + 77)
    n = len(fr['kps_un'])
    kf_mp = np.full(n, -1, np.int32)
    holders = r.permutation(npts)[: int(min(npts, n) * p_occupied)]
    slots = r.permutation(n)[: len(holders)]
    kf_mp[slots] = holders
    in_kf = np.zeros(npts, np.uint8)
    in_kf[holders] = 1                               # the holders are candidates too: IsInKeyFrame skips them (:875)
    nobs = r.randint(0, 9, npts).astype(np.int32)
    bad = (r.rand(npts) < 0.05).astype(np.uint8)
    return dict(kf_mp=kf_mp, nobs=nobs, bad=bad, in_kf=in_kf)


def sim3_pair(seed, n=1200, w=640, h=480, cam=(517.3, 516.5, 318.6, 255.3, 40.0, 0.0773), scale=1.1, max_flips=50):
    """Two key frames for SearchBySim3 (src/ORBmatcher.cc:1084-1277). Key frame k has pose_k and one map point per keypoint (some null /
    already matched / bad). For a set of pairs (i1, i2) the map point of kf1's keypoint i1, taken through pose1 and S21, lands near kf2's
    keypoint i2 and looks like it, and the other way round, so that part of the two directed searches agree; the rest disagree, miss the
    image, the depth range or the octave. Returns dict(f1, f2, pose1, pose2, S12, pts1, desc1, pts2, desc2)."""
    r = np.random.RandomState(seed + 31)
    f1 = frame(seed + 300, n, w, h, stereo=False)
    f2 = frame(seed + 301, n, w, h, stereo=False)
    fx, fy, cx, cy = [np.float64(c) for c in cam[:4]]

    def rot(ax, ang):
        ax = np.asarray(ax, np.float64) / np.linalg.norm(ax)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    R1, t1 = rot(r.randn(3), 0.2), r.randn(3) * 0.5
    R2, t2 = rot(r.randn(3), 0.3), r.randn(3) * 0.5
    R12, t12, s12 = rot(r.randn(3), 0.05), r.randn(3) * 0.05, np.float64(scale)
    sf = f1['scale_factors'].astype(np.float64)
    nl = f1['nlevels']
    pts = [np.zeros(n, KF_POINT_DTYPE), np.zeros(n, KF_POINT_DTYPE)]
    desc = [r.randint(0, 256, (n, 32)).astype(np.uint8), r.randint(0, 256, (n, 32)).astype(np.uint8)]
    m = int(0.7 * n)
    a = r.permutation(n)[:m]
    b = r.permutation(n)[:m]
    frames = (f1, f2)
    # direction 0: point of kf1[a] -> kf2 keypoint b; direction 1: point of kf2[b'] -> kf1 keypoint a' (b', a' = same pairs for 60 %, shuffled rest)
    b2 = b.copy(); a2 = a.copy()
    sh = r.rand(m) < 0.4
    a2[sh] = a[sh][r.permutation(sh.sum())]
    for d, (src, dst) in enumerate(((a, b), (b2, a2))):
        fs, ft = frames[d], frames[1 - d]
        kt = ft['kps_un'][dst]
        z = 2.0 + 6.0 * r.rand(m)
        u = kt['x'].astype(np.float64) + r.randn(m) * 1.5
        v = kt['y'].astype(np.float64) + r.randn(m) * 1.5
        Xt = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)              # camera coordinates of the target key frame
        if d == 0:      # Xc2 = S21.Map(Xc1)  =>  Xc1 = S12.Map(Xc2) = s12 R12 Xc2 + t12
            Xs = (s12 * (R12 @ Xt.T)).T + t12
            Rs, ts = R1, t1
        else:           # Xc1 = S12.Map(Xc2)  =>  Xc2 = S21.Map(Xc1) = (1/s12) R12^T (Xc1 - t12)
            Xs = ((R12.T @ (Xt - t12).T) / s12).T
            Rs, ts = R2, t2
        Xw = (Rs.T @ (Xs - ts).T).T
        lvl = np.clip(kt['octave'] + r.choice([0, 0, 0, 1, -1, 2], m), 0, nl - 1)
        dist = np.linalg.norm(Xt, axis=1)
        P = pts[d]
        P['xw'][src] = Xw.astype(np.float32)
        P['max_distance'][src] = (dist * sf[lvl] * (1.0 + 0.03 * r.randn(m))).astype(np.float32)
        P['min_distance'][src] = (P['max_distance'][src] / sf[nl - 1]).astype(np.float32)
        far = r.rand(m) < 0.05
        P['max_distance'][src[far]] *= np.float32(0.3)
        P['flags'][src] = (r.rand(m) < 0.9).astype(np.int32)
        for s_, t_ in zip(src, dst):
            desc[d][s_] = _flip_bits(r, ft['desc'][t_], r.randint(0, max_flips + 1))
    f32 = lambda x: np.asarray(x, np.float32)
    return dict(f1=f1, f2=f2, cam=cam, pose1=(f32(R1), f32(t1)), pose2=(f32(R2), f32(t2)), S12=(f32(R12), f32(t12), np.float32(s12)),
                pts1=pts[0], desc1=desc[0], pts2=pts[1], desc2=desc[1])


def triangulation_pair(seed, n=1500, w=640, h=480, fy=516.5):
    """Two key frames for SearchForTriangulation (src/ORBmatcher.cc:768-866): the bow_pair() scene with the matched keypoints of frame 2
    moved onto (or near, or off) the epipolar line of their partner under a sideways-translation fundamental matrix with a small
    perturbation, some stereo keypoints, map points on part of both frames and an epipole inside the image.
    Returns dict(f1, fv1, has1, f2, fv2, has2, F12, ep2, sigma_sq2)."""
    r = np.random.RandomState(seed + 57)
    f1, fv1, va1, f2, fv2, va2 = bow_pair(seed, n, w, h)
    k1, k2 = f1['kps_un'], f2['kps_un']
    # re-pair by descriptor: every keypoint of frame 2 whose descriptor was copied from frame 1 gets its row
    d1 = np.unpackbits(f1['desc'], axis=1).astype(np.int16); d2 = np.unpackbits(f2['desc'], axis=1).astype(np.int16)
    for j in range(0, n, 3):
        dist = np.abs(d1 - d2[j]).sum(1)
        i = int(np.argmin(dist))
        if dist[i] <= 60:
            k2['y'][j] = k1['y'][i] + np.float32(r.choice([0.0, 0.5, 1.5, 2.5, 6.0]) * r.choice([-1, 1]) * f2['scale_factors'][k2['octave'][j]])
            k1['y'][np.flatnonzero(dist <= 60)] = k1['y'][i]      # its look-alikes in frame 1 share the line: several idx1 may end on this idx2
    F = np.array([[0, 0, 0], [0, 0, -1.0 / fy], [0, 1.0 / fy, 0]], np.float64)
    F = F + r.randn(3, 3) * np.array([[1e-7, 1e-7, 1e-5], [1e-7, 1e-7, 1e-5], [1e-5, 1e-5, 1e-4]])
    ur1 = np.full(n, -1, np.float32); ur2 = np.full(n, -1, np.float32)
    s1 = r.rand(n) < 0.3; s2 = r.rand(n) < 0.3
    ur1[s1] = k1['x'][s1] - 5.0; ur2[s2] = k2['x'][s2] - 5.0
    f1['uright'], f2['uright'] = ur1, ur2
    sig, _ = sigma_tables(f2['scale_factors'])
    j = r.randint(0, n)
    ep = np.array([k2['x'][j] + 3.0, k2['y'][j] - 2.0], np.float32)
    return dict(f1=f1, fv1=fv1, has1=(r.rand(n) < 0.3).astype(np.uint8), f2=f2, fv2=fv2, has2=(r.rand(n) < 0.3).astype(np.uint8),
                F12=F.astype(np.float32), ep2=ep, sigma_sq2=sig)
