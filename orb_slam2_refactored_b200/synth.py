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
