"""TEST INFRASTRUCTURE — ctypes bindings of the CPU oracle (oracle/oracle_api.h). Not the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Two libraries export the same API under different prefixes:
  ref  -> oracle/_ref/liborb_ref.so      (the reference's own TUs; built where /root/reference exists)
  port -> oracle/_build/liborb_oracle.so (stand-alone restatement, orb_oracle.cc)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

KP_DTYPE = np.dtype([('x', '<f4'), ('y', '<f4'), ('size', '<f4'), ('angle', '<f4'), ('response', '<f4'),
                     ('octave', '<i4'), ('class_id', '<i4')])
CAND_DTYPE = np.dtype([('x', '<i4'), ('y', '<i4'), ('response', '<i4')])
assert KP_DTYPE.itemsize == 28 and CAND_DTYPE.itemsize == 12


TRACK_POINT_DTYPE = np.dtype([('proj_x', '<f4'), ('proj_y', '<f4'), ('proj_xr', '<f4'), ('view_cos', '<f4'), ('scale_level', '<i4'),
                              ('flags', '<i4')])
LAST_POINT_DTYPE = np.dtype([('xw', '<f4', (3,)), ('octave', '<i4'), ('angle', '<f4'), ('flags', '<i4')])
SIM3_POINT_DTYPE = np.dtype([('xw', '<f4', (3,)), ('normal', '<f4', (3,)), ('min_distance', '<f4'), ('max_distance', '<f4'), ('flags', '<i4')])
KF_POINT_DTYPE = np.dtype([('xw', '<f4', (3,)), ('min_distance', '<f4'), ('max_distance', '<f4'), ('angle', '<f4'), ('flags', '<i4')])
assert TRACK_POINT_DTYPE.itemsize == 24 and LAST_POINT_DTYPE.itemsize == 24 and KF_POINT_DTYPE.itemsize == 28


class Bounds(C.Structure):
    _fields_ = [(n, C.c_float) for n in ('minx', 'maxx', 'miny', 'maxy')]


class FrameView(C.Structure):
    _fields_ = [('n', C.c_int32), ('kps_un', C.c_void_p), ('desc', C.c_void_p), ('uright', C.c_void_p), ('bounds', Bounds),
                ('nlevels', C.c_int32), ('scale_factors', C.c_void_p)]


class FeatureVector(C.Structure):
    _fields_ = [('nnodes', C.c_int32), ('node_ids', C.c_void_p), ('start', C.c_void_p), ('indices', C.c_void_p)]


class Sim3(C.Structure):
    _fields_ = [('R', C.c_float * 9), ('t', C.c_float * 3), ('s', C.c_float)]


class Pose(C.Structure):
    _fields_ = [('R', C.c_float * 9), ('t', C.c_float * 3)]


class Camera(C.Structure):
    _fields_ = [(n, C.c_float) for n in ('fx', 'fy', 'cx', 'cy', 'bf', 'baseline')]


def build(native=True):
    """Compile the oracle libraries: the restatement always, the reference's own TUs only where /root/reference is
    present (elsewhere the prebuilt oracle/_ref/*.so that travelled with the snapshot is used). native=True also
    builds the -O3 timing variants used by bench.py's CPU baseline."""
    have_ref = os.path.isdir('/root/reference/src')
    targets = ['port'] + (['ref'] if have_ref else [])
    if native:
        targets += [os.path.join(HERE, '_build', 'liborb_oracle_native.so')]
        if have_ref:
            targets += [os.path.join(HERE, '_ref', 'liborb_ref_native.so')]
    subprocess.run(['make', '-s', '-C', HERE] + targets, check=True)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """kind: 'ref' | 'port'; native=True loads the -O3 -march=native timing build."""

    def __init__(self, kind='port', native=False):
        self.kind = kind
        suffix = '_native' if native else ''
        if kind == 'ref':
            path, self.pre = os.path.join(HERE, '_ref', f'liborb_ref{suffix}.so'), 'ref_'
        else:
            path, self.pre = os.path.join(HERE, '_build', f'liborb_oracle{suffix}.so'), 'orc_'
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = C.CDLL(path)
        f = self._f
        f('extractor_create', C.c_void_p, [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int])
        f('extractor_destroy', None, [C.c_void_p])
        f('extractor_extract', C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int])
        f('extractor_level_size', C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)])
        f('extractor_level_copy', C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t])
        f('extractor_tables', None, [C.c_void_p] + [C.c_void_p] * 4)
        f('feature_quotas', None, [C.c_int, C.c_float, C.c_int, C.c_void_p])
        f('detect_fast', C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int])
        f('quadtree', C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int])
        f('ic_angle', C.c_float, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int])
        f('descriptor', None, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_void_p])
        f('descriptor_distance', C.c_int, [C.c_void_p, C.c_void_p])
        f('stereo_matches', C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Camera),
                                      C.c_void_p, C.c_void_p])
        f('knn2', None, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                         C.c_void_p, C.c_int])
        f('cos_sin_range', None, [C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int])
        f('convert_to_gray', None, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t])
        f('stereo_from_rgbd', None, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(Camera), C.c_void_p, C.c_void_p])
        f('distinctive_index', C.c_int, [C.c_void_p, C.c_int])
        f('undistort_keypoints', None, [C.c_void_p, C.c_int, C.POINTER(Camera), C.c_void_p, C.c_int, C.c_void_p])
        f('grid_create', C.c_void_p, [C.c_void_p, C.c_int, C.POINTER(Bounds), C.c_int])
        f('grid_destroy', None, [C.c_void_p])
        f('grid_query', C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int])
        f('search_local_map', C.c_int, [C.POINTER(FrameView), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float])
        f('search_last_frame', C.c_int, [C.POINTER(FrameView), C.POINTER(Camera), C.POINTER(Pose), C.POINTER(Pose), C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int])
        f('time_search_local_map', C.c_double, [C.POINTER(FrameView), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int])
        f('time_search_last_frame', C.c_double, [C.POINTER(FrameView), C.POINTER(Camera), C.POINTER(Pose), C.POINTER(Pose), C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int])
        f('search_keyframe_projection', C.c_int, [C.POINTER(FrameView), C.POINTER(Camera), C.POINTER(Pose), C.c_float, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int])
        f('search_sim3_projection', C.c_int, [C.POINTER(FrameView), C.POINTER(Camera), C.POINTER(Sim3), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_int])
        f('search_by_bow', C.c_int, [C.POINTER(FrameView), C.POINTER(FeatureVector), C.c_void_p, C.POINTER(FrameView), C.POINTER(FeatureVector),
                                     C.c_void_p, C.c_float, C.c_int, C.c_void_p])
        f('search_for_initialization', C.c_int, [C.POINTER(FrameView), C.POINTER(FrameView), C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int])
        f('fuse', C.c_int, [C.POINTER(FrameView), C.POINTER(Camera), C.POINTER(Pose), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)])
        f('fuse_sim3', C.c_int, [C.POINTER(FrameView), C.POINTER(Camera), C.POINTER(Sim3), C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)])
        f('search_by_sim3', C.c_int, [C.POINTER(FrameView), C.POINTER(Camera), C.POINTER(Pose), C.c_float, C.POINTER(FrameView), C.POINTER(Camera),
                                      C.POINTER(Pose), C.c_float, C.POINTER(Sim3), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p])
        f('search_for_triangulation', C.c_int, [C.POINTER(FrameView), C.POINTER(FeatureVector), C.c_void_p, C.POINTER(FrameView), C.POINTER(FeatureVector),
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p])
        if self.pre == 'ref_':
            f('time_fuse', C.c_double, [C.POINTER(FrameView), C.POINTER(Camera), C.POINTER(Pose), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int])
            f('time_search_by_sim3', C.c_double, [C.POINTER(FrameView), C.POINTER(Camera), C.POINTER(Pose), C.c_float, C.POINTER(FrameView), C.POINTER(Camera),
                                                  C.POINTER(Pose), C.c_float, C.POINTER(Sim3), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int])
            f('time_search_for_triangulation', C.c_double, [C.POINTER(FrameView), C.POINTER(FeatureVector), C.c_void_p, C.POINTER(FrameView),
                                                            C.POINTER(FeatureVector), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int])
        f('voc_load_text', C.c_void_p, [C.c_char_p])
        f('voc_destroy', None, [C.c_void_p])
        f('bow_transform', C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6)
        f('bow_score', C.c_double, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int])
        f('time_bow_transform', C.c_double, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int])
        if self.pre == 'orc_':
            f('voc_create', C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p])
            f('bow_descend', None, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p])
        f('cv_resize', None, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t])
        f('cv_fast', C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int])
        f('cv_gaussian7', None, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t])
        f('cv_remap', None, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t])
        f('cv_fast_atan2', C.c_float, [C.c_float, C.c_float])
        f('cv_round_f', C.c_int, [C.c_float])
        f('cv_round_d', C.c_int, [C.c_double])

    def _f(self, name, res, args):
        fn = getattr(self.lib, self.pre + name)
        fn.restype, fn.argtypes = res, args
        setattr(self, '_' + name, fn)

    # ---- primitives ----
    def resize(self, src, dw, dh):
        src = np.ascontiguousarray(src, np.uint8)
        dst = np.empty((dh, dw), np.uint8)
        self._cv_resize(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dst.strides[0])
        return dst

    def fast(self, img, th, nms=True):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.empty(img.size // 2 + 16, CAND_DTYPE)
        n = self._cv_fast(_p(img), img.shape[1], img.shape[0], img.strides[0], th, int(nms), _p(out), len(out))
        assert n >= 0
        return out[:n].copy()

    def gaussian7(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        dst = np.empty_like(img)
        self._cv_gaussian7(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(dst), dst.strides[0])
        return dst

    def remap(self, img, mapx, mapy):
        """cv::remap(img, mapx, mapy, INTER_LINEAR) for an 8-bit image and two float32 maps of the output size."""
        img = np.ascontiguousarray(img, np.uint8)
        mapx = np.ascontiguousarray(mapx, np.float32); mapy = np.ascontiguousarray(mapy, np.float32)
        dst = np.empty(mapx.shape, np.uint8)
        self._cv_remap(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(mapx), _p(mapy), mapx.strides[0], _p(dst), dst.shape[1], dst.shape[0],
                       dst.strides[0])
        return dst

    def fast_atan2(self, y, x):
        return np.float32(self._cv_fast_atan2(float(y), float(x)))

    # ---- stages ----
    def quotas(self, nfeatures, scale, nlevels):
        out = np.zeros(nlevels, np.int32)
        self._feature_quotas(nfeatures, scale, nlevels, _p(out))
        return out

    def detect_fast(self, img, ini=20, mn=7):
        img = np.ascontiguousarray(img, np.uint8)
        out = np.empty(img.size // 4 + 16, CAND_DTYPE)
        n = self._detect_fast(_p(img), img.shape[1], img.shape[0], img.strides[0], ini, mn, _p(out), len(out))
        assert n >= 0
        return out[:n].copy()

    def quadtree(self, cand, w, h, nfeatures):
        cand = np.ascontiguousarray(cand, CAND_DTYPE)
        out = np.empty(len(cand) + 16, CAND_DTYPE)
        n = self._quadtree(_p(cand), len(cand), w, h, nfeatures, _p(out), len(out))
        assert n >= 0
        return out[:n].copy()

    def ic_angle(self, img, x, y):
        img = np.ascontiguousarray(img, np.uint8)
        return np.float32(self._ic_angle(_p(img), img.shape[1], img.shape[0], img.strides[0], int(x), int(y)))

    def descriptor(self, blurred, x, y, angle):
        blurred = np.ascontiguousarray(blurred, np.uint8)
        d = np.empty(32, np.uint8)
        self._descriptor(_p(blurred), blurred.shape[1], blurred.shape[0], blurred.strides[0], int(x), int(y), float(angle), _p(d))
        return d

    def descriptor_distance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return int(self._descriptor_distance(_p(a), _p(b)))

    # ---- SURVEY §8(f) rows ----
    def convert_to_gray(self, img, rgb=True):
        img = np.ascontiguousarray(img, np.uint8)
        h, w, ch = img.shape
        dst = np.empty((h, w), np.uint8)
        self._convert_to_gray(_p(img), w, h, img.strides[0], ch, int(rgb), _p(dst), dst.strides[0])
        return dst

    def stereo_from_rgbd(self, kps, kps_un, depth_map, cam):
        kps = np.ascontiguousarray(kps, KP_DTYPE); kps_un = np.ascontiguousarray(kps_un, KP_DTYPE)
        depth_map = np.ascontiguousarray(depth_map, np.float32)
        ur = np.empty(len(kps), np.float32); dp = np.empty(len(kps), np.float32)
        c = Camera(*[float(v) for v in cam])
        self._stereo_from_rgbd(_p(kps), _p(kps_un), len(kps), _p(depth_map), depth_map.shape[1], depth_map.shape[0], depth_map.strides[0],
                               C.byref(c), _p(ur), _p(dp))
        return ur, dp

    def undistort_keypoints(self, kps, cam, dist):
        """UndistortKeyPoints (src/System.cc:153-174): cam = (fx, fy, cx, cy, bf, baseline), dist = OpenCV distortion coefficients."""
        kps = np.ascontiguousarray(kps).view(KP_DTYPE)
        d = np.ascontiguousarray(dist, np.float32)
        out = np.empty_like(kps)
        self._undistort_keypoints(_p(kps), len(kps), C.byref(Camera(*[float(c) for c in cam])), _p(d), len(d), _p(out))
        return out

    def distinctive_index(self, desc):
        desc = np.ascontiguousarray(desc, np.uint8)
        return int(self._distinctive_index(_p(desc), len(desc)))

    # ---- extractor ----
    # ---- guided matchers. `frame` is a dict: kps_un, desc, uright (or None), bounds (4 floats), nlevels, scale_factors
    @staticmethod
    def _frame_view(frame):
        keep = {
            'kps': np.ascontiguousarray(frame['kps_un']).view(KP_DTYPE),
            'desc': np.ascontiguousarray(frame['desc'], np.uint8),
            'ur': None if frame.get('uright') is None else np.ascontiguousarray(frame['uright'], np.float32),
            'sf': np.ascontiguousarray(frame['scale_factors'], np.float32),
        }
        v = FrameView(len(keep['kps']), keep['kps'].ctypes.data, keep['desc'].ctypes.data,
                      None if keep['ur'] is None else keep['ur'].ctypes.data, Bounds(*[float(b) for b in frame['bounds']]),
                      int(frame['nlevels']), keep['sf'].ctypes.data)
        return v, keep

    def grid_queries(self, frame, queries):
        """queries: iterable of (x, y, r, min_level, max_level) -> list of index arrays in the reference's output order."""
        kps = np.ascontiguousarray(frame['kps_un']).view(KP_DTYPE)
        g = self._grid_create(_p(kps), len(kps), C.byref(Bounds(*[float(b) for b in frame['bounds']])), int(frame['nlevels']))
        out, buf = [], np.empty(len(kps) + 1, np.int32)
        for (x, y, r, lo, hi) in queries:
            n = self._grid_query(g, float(x), float(y), float(r), int(lo), int(hi), _p(buf), len(buf))
            out.append(buf[:n].copy())
        self._grid_destroy(g)
        return out

    def search_local_map(self, frame, frame_mp, pts, pt_desc, th=3.0, nnratio=0.8):
        v, keep = self._frame_view(frame)
        mp = np.array(frame_mp, np.int32)
        pts = np.ascontiguousarray(pts).view(TRACK_POINT_DTYPE)
        pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
        n = self._search_local_map(C.byref(v), _p(mp), _p(pts), _p(pt_desc), len(pts), th, nnratio)
        return n, mp

    def search_last_frame(self, frame, cam, cur_pose, last_pose, frame_mp, pts, pt_desc, th=15.0, monocular=False, nnratio=0.9,
                          check_orientation=True):
        v, keep = self._frame_view(frame)
        mp = np.array(frame_mp, np.int32)
        pts = np.ascontiguousarray(pts).view(LAST_POINT_DTYPE)
        pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
        poses = []
        for R, t in (cur_pose, last_pose):
            P = Pose()
            P.R[:] = [float(x) for x in np.asarray(R, np.float32).reshape(9)]
            P.t[:] = [float(x) for x in np.asarray(t, np.float32).reshape(3)]
            poses.append(P)
        n = self._search_last_frame(C.byref(v), C.byref(Camera(*[float(c) for c in cam])), C.byref(poses[0]), C.byref(poses[1]), _p(mp),
                                    _p(pts), _p(pt_desc), len(pts), th, int(monocular), nnratio, int(check_orientation))
        return n, mp

    @staticmethod
    def _poses(cur_pose, last_pose):
        poses = []
        for R, t in (cur_pose, last_pose):
            P = Pose()
            P.R[:] = [float(x) for x in np.asarray(R, np.float32).reshape(9)]
            P.t[:] = [float(x) for x in np.asarray(t, np.float32).reshape(3)]
            poses.append(P)
        return poses

    def time_search_local_map(self, frame, frame_mp, pts, pt_desc, th=3.0, nnratio=0.8, reps=20):
        """Mean seconds per SearchByProjection(Frame&, mappoints, th) call on one host core, Frame and grid built outside the timed region."""
        v, keep = self._frame_view(frame)
        mp = np.ascontiguousarray(frame_mp, np.int32)
        pts = np.ascontiguousarray(pts).view(TRACK_POINT_DTYPE)
        pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
        return self._time_search_local_map(C.byref(v), _p(mp), _p(pts), _p(pt_desc), len(pts), th, nnratio, reps)

    def time_search_last_frame(self, frame, cam, cur_pose, last_pose, frame_mp, pts, pt_desc, th=15.0, monocular=False, nnratio=0.9,
                               check_orientation=True, reps=20):
        v, keep = self._frame_view(frame)
        mp = np.ascontiguousarray(frame_mp, np.int32)
        pts = np.ascontiguousarray(pts).view(LAST_POINT_DTYPE)
        pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
        poses = self._poses(cur_pose, last_pose)
        return self._time_search_last_frame(C.byref(v), C.byref(Camera(*[float(c) for c in cam])), C.byref(poses[0]), C.byref(poses[1]),
                                            _p(mp), _p(pts), _p(pt_desc), len(pts), th, int(monocular), nnratio, int(check_orientation), reps)

    @staticmethod
    def _fv(fv):
        """fv: (node_ids uint32 ascending, start int32 [nnodes + 1], indices uint32)."""
        ids = np.ascontiguousarray(fv[0], np.uint32); start = np.ascontiguousarray(fv[1], np.int32); idx = np.ascontiguousarray(fv[2], np.uint32)
        return FeatureVector(len(ids), ids.ctypes.data, start.ctypes.data, idx.ctypes.data), (ids, start, idx)

    def search_by_bow(self, f1, fv1, valid1, f2, fv2, valid2=None, nnratio=0.7, check_orientation=True):
        """SearchByBoW; valid2 None = KeyFrame vs Frame, else KeyFrame vs KeyFrame. Returns (nmatches, match2)."""
        v1, k1 = self._frame_view(f1)
        v2, k2 = self._frame_view(f2)
        c1, keep1 = self._fv(fv1)
        c2, keep2 = self._fv(fv2)
        va1 = np.ascontiguousarray(valid1, np.uint8)
        va2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
        m2 = np.empty(v2.n, np.int32)
        n = self._search_by_bow(C.byref(v1), C.byref(c1), _p(va1), C.byref(v2), C.byref(c2), None if va2 is None else _p(va2), nnratio,
                                int(check_orientation), _p(m2))
        return n, m2

    def search_keyframe_projection(self, frame, cam, pose, log_scale_factor, frame_mp, pts, pt_desc, th=10.0, orb_dist=100,
                                   check_orientation=True):
        """SearchByProjection(Frame&, KeyFrame*, alreadyFound, th, ORBdist) (relocalisation). Returns (nmatches, frame_mp)."""
        v, keep = self._frame_view(frame)
        mp = np.array(frame_mp, np.int32)
        pts = np.ascontiguousarray(pts).view(KF_POINT_DTYPE)
        pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
        P = self._poses(pose, pose)[0]
        n = self._search_keyframe_projection(C.byref(v), C.byref(Camera(*[float(c) for c in cam])), C.byref(P), float(np.float32(log_scale_factor)),
                                             _p(mp), _p(pts), _p(pt_desc), len(pts), th, int(orb_dist), int(check_orientation))
        return n, mp

    def search_sim3_projection(self, keyframe, cam, sim3, log_scale_factor, matched, pts, pt_desc, th=10):
        """SearchByProjection(keyframe, Scw, mappoints, matched, th) (loop closing). sim3 = (R 3x3, t 3, s). Returns (nmatches, matched)."""
        v, keep = self._frame_view(keyframe)
        m = np.array(matched, np.int32)
        pts = np.ascontiguousarray(pts).view(SIM3_POINT_DTYPE)
        pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
        S = Sim3()
        S.R[:] = [float(x) for x in np.asarray(sim3[0], np.float32).reshape(9)]
        S.t[:] = [float(x) for x in np.asarray(sim3[1], np.float32).reshape(3)]
        S.s = float(np.float32(sim3[2]))
        n = self._search_sim3_projection(C.byref(v), C.byref(Camera(*[float(c) for c in cam])), C.byref(S), float(np.float32(log_scale_factor)), _p(m),
                                         _p(pts), _p(pt_desc), len(pts), int(th))
        return n, m

    def search_for_initialization(self, f1, f2, prev_matched, window=100, nnratio=0.9, check_orientation=True):
        v1, k1 = self._frame_view(f1)
        v2, k2 = self._frame_view(f2)
        prev = np.array(prev_matched, np.float32).reshape(-1, 2)
        m12 = np.empty(v1.n, np.int32)
        n = self._search_for_initialization(C.byref(v1), C.byref(v2), _p(prev), _p(m12), int(window), nnratio, int(check_orientation))
        return n, m12, prev

    @staticmethod
    def _sim3(S):
        out = Sim3()
        out.R[:] = [float(x) for x in np.asarray(S[0], np.float32).reshape(9)]
        out.t[:] = [float(x) for x in np.asarray(S[1], np.float32).reshape(3)]
        out.s = float(np.float32(S[2]))
        return out

    @staticmethod
    def _pose1(p):
        out = Pose()
        out.R[:] = [float(x) for x in np.asarray(p[0], np.float32).reshape(9)]
        out.t[:] = [float(x) for x in np.asarray(p[1], np.float32).reshape(3)]
        return out

    def fuse(self, kf, cam, pose, log_scale_factor, inv_sigma_sq, pts, pt_desc, th, state):
        """Fuse(keyframe, mappoints, th) on the one-key-frame map model; state = dict(kf_mp, nobs, bad, in_kf) (copied).
        Returns (nfused, kf_mp, nobs, bad, in_kf, log)."""
        v, keep = self._frame_view(kf)
        pts = np.ascontiguousarray(pts); pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
        sig = np.ascontiguousarray(inv_sigma_sq, np.float32)
        kf_mp = state['kf_mp'].astype(np.int32).copy(); nobs = state['nobs'].astype(np.int32).copy()
        bad = state['bad'].astype(np.uint8).copy(); in_kf = state['in_kf'].astype(np.uint8).copy()
        log = np.zeros(3 * len(pts) + 3, np.int32); nlog = C.c_int(0)
        n = self._fuse(C.byref(v), C.byref(Camera(*[float(c) for c in cam])), C.byref(self._pose1(pose)), float(log_scale_factor), _p(sig), _p(pts),
                       _p(pt_desc), len(pts), float(th), _p(kf_mp), _p(nobs), _p(bad), _p(in_kf), _p(log), len(log), C.byref(nlog))
        return n, kf_mp, nobs, bad, in_kf, log[:nlog.value].copy()

    def fuse_sim3(self, kf, cam, sim3, log_scale_factor, pts, pt_desc, th, state):
        """Fuse(keyframe, Scw, mappoints, th, replacePoints). Returns (nfused, kf_mp, nobs, bad, replace, log)."""
        v, keep = self._frame_view(kf)
        pts = np.ascontiguousarray(pts); pt_desc = np.ascontiguousarray(pt_desc, np.uint8)
        kf_mp = state['kf_mp'].astype(np.int32).copy(); nobs = state['nobs'].astype(np.int32).copy(); bad = state['bad'].astype(np.uint8).copy()
        rep = np.full(max(len(pts), 1), -1, np.int32)
        log = np.zeros(3 * len(pts) + 3, np.int32); nlog = C.c_int(0)
        n = self._fuse_sim3(C.byref(v), C.byref(Camera(*[float(c) for c in cam])), C.byref(self._sim3(sim3)), float(log_scale_factor), _p(pts),
                            _p(pt_desc), len(pts), float(th), _p(kf_mp), _p(nobs), _p(bad), _p(rep), _p(log), len(log), C.byref(nlog))
        return n, kf_mp, nobs, bad, rep[:len(pts)], log[:nlog.value].copy()

    def search_by_sim3(self, c, th):
        """SearchBySim3 on a synth.sim3_pair() scene. Returns (nfound, matches12)."""
        v1, k1 = self._frame_view(c['f1']); v2, k2 = self._frame_view(c['f2'])
        cam = Camera(*[float(x) for x in c['cam']])
        p1 = np.ascontiguousarray(c['pts1']); p2 = np.ascontiguousarray(c['pts2'])
        d1 = np.ascontiguousarray(c['desc1'], np.uint8); d2 = np.ascontiguousarray(c['desc2'], np.uint8)
        m12 = np.empty(max(v1.n, 1), np.int32)
        lsf = float(c['lsf'])
        n = self._search_by_sim3(C.byref(v1), C.byref(cam), C.byref(self._pose1(c['pose1'])), lsf, C.byref(v2), C.byref(cam),
                                 C.byref(self._pose1(c['pose2'])), lsf, C.byref(self._sim3(c['S12'])), float(th), _p(p1), _p(d1), _p(p2), _p(d2), _p(m12))
        return n, m12[:v1.n]

    def search_for_triangulation(self, c, only_stereo, check_orientation):
        v1, k1 = self._frame_view(c['f1']); v2, k2 = self._frame_view(c['f2'])
        c1, keep1 = self._fv(c['fv1']); c2, keep2 = self._fv(c['fv2'])
        h1 = np.ascontiguousarray(c['has1'], np.uint8); h2 = np.ascontiguousarray(c['has2'], np.uint8)
        F = np.ascontiguousarray(c['F12'], np.float32).reshape(9); ep = np.ascontiguousarray(c['ep2'], np.float32); sg = np.ascontiguousarray(c['sigma_sq2'], np.float32)
        m12 = np.empty(max(v1.n, 1), np.int32)
        n = self._search_for_triangulation(C.byref(v1), C.byref(c1), _p(h1), C.byref(v2), C.byref(c2), _p(h2), _p(F), _p(ep), _p(sg), int(only_stereo),
                                           int(check_orientation), _p(m12))
        return n, m12[:v1.n]

    def time_fuse(self, kf, cam, pose, log_scale_factor, inv_sigma_sq, pts, pt_desc, th, state, reps=20):
        v, keep = self._frame_view(kf)
        pts = np.ascontiguousarray(pts); pt_desc = np.ascontiguousarray(pt_desc, np.uint8); sig = np.ascontiguousarray(inv_sigma_sq, np.float32)
        a = [np.ascontiguousarray(state[k], t) for k, t in (('kf_mp', np.int32), ('nobs', np.int32), ('bad', np.uint8), ('in_kf', np.uint8))]
        return self._time_fuse(C.byref(v), C.byref(Camera(*[float(c) for c in cam])), C.byref(self._pose1(pose)), float(log_scale_factor), _p(sig), _p(pts),
                               _p(pt_desc), len(pts), float(th), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), reps)

    def time_search_by_sim3(self, c, th, reps=20):
        v1, k1 = self._frame_view(c['f1']); v2, k2 = self._frame_view(c['f2'])
        cam = Camera(*[float(x) for x in c['cam']])
        p1 = np.ascontiguousarray(c['pts1']); p2 = np.ascontiguousarray(c['pts2'])
        d1 = np.ascontiguousarray(c['desc1'], np.uint8); d2 = np.ascontiguousarray(c['desc2'], np.uint8)
        lsf = float(c['lsf'])
        return self._time_search_by_sim3(C.byref(v1), C.byref(cam), C.byref(self._pose1(c['pose1'])), lsf, C.byref(v2), C.byref(cam),
                                         C.byref(self._pose1(c['pose2'])), lsf, C.byref(self._sim3(c['S12'])), float(th), _p(p1), _p(d1), _p(p2), _p(d2), reps)

    def time_search_for_triangulation(self, c, only_stereo, check_orientation, reps=20):
        v1, k1 = self._frame_view(c['f1']); v2, k2 = self._frame_view(c['f2'])
        c1, keep1 = self._fv(c['fv1']); c2, keep2 = self._fv(c['fv2'])
        h1 = np.ascontiguousarray(c['has1'], np.uint8); h2 = np.ascontiguousarray(c['has2'], np.uint8)
        F = np.ascontiguousarray(c['F12'], np.float32).reshape(9); ep = np.ascontiguousarray(c['ep2'], np.float32); sg = np.ascontiguousarray(c['sigma_sq2'], np.float32)
        return self._time_search_for_triangulation(C.byref(v1), C.byref(c1), _p(h1), C.byref(v2), C.byref(c2), _p(h2), _p(F), _p(ep), _p(sg),
                                                   int(only_stereo), int(check_orientation), reps)

    def vocabulary(self, path=None, arrays=None):
        """path: a vocabulary text file (loadFromTextFile); arrays (port only): dict k, L, scoring, weighting, parent, is_leaf, desc, weights."""
        return OracleVocabulary(self, path, arrays)

    def extractor(self, nfeatures=2000, scale=1.2, nlevels=8, ini=20, mn=7):
        return _Extractor(self, nfeatures, scale, nlevels, ini, mn)

    def stereo(self, kpL, descL, pyrL, kpR, descR, pyrR, scale, inv_scale, cam):
        n = len(pyrL)
        pyrL = [np.ascontiguousarray(p, np.uint8) for p in pyrL]
        pyrR = [np.ascontiguousarray(p, np.uint8) for p in pyrR]
        pl = (C.c_void_p * n)(*[p.ctypes.data for p in pyrL])
        pr = (C.c_void_p * n)(*[p.ctypes.data for p in pyrR])
        lw = np.array([p.shape[1] for p in pyrL], np.int32)
        lh = np.array([p.shape[0] for p in pyrL], np.int32)
        lp = np.array([p.strides[0] for p in pyrL], np.uint64)
        kpL = np.ascontiguousarray(kpL, KP_DTYPE); kpR = np.ascontiguousarray(kpR, KP_DTYPE)
        descL = np.ascontiguousarray(descL, np.uint8); descR = np.ascontiguousarray(descR, np.uint8)
        scale = np.ascontiguousarray(scale, np.float32); inv_scale = np.ascontiguousarray(inv_scale, np.float32)
        ur = np.empty(len(kpL), np.float32); dp = np.empty(len(kpL), np.float32)
        c = Camera(*[float(v) for v in cam])
        rc = self._stereo_matches(_p(kpL), len(kpL), _p(descL), pl, _p(kpR), len(kpR), _p(descR), pr, _p(lw), _p(lh), _p(lp), n,
                                  _p(scale), _p(inv_scale), C.byref(c), _p(ur), _p(dp))
        return rc, ur, dp

    def cos_sin_range(self, first_bits, n, threads=1):
        """(cos, sin) float32 arrays for the n float angles (degrees) whose bit patterns start at first_bits (src/ORBextractor.cc:105-107)."""
        c = np.empty(n, np.float32); s = np.empty(n, np.float32)
        self._cos_sin_range(first_bits, n, _p(c), _p(s), threads)
        return c, s

    def knn2(self, query, train, th_low=50, nnratio=0.6, threads=1):
        query = np.ascontiguousarray(query, np.uint8); train = np.ascontiguousarray(train, np.uint8)
        nq = len(query)
        idx = np.empty(nq, np.int32); best = np.empty(nq, np.uint16); second = np.empty(nq, np.uint16)
        match = np.empty(nq, np.int32)
        self._knn2(_p(query), nq, _p(train), len(train), th_low, nnratio, _p(idx), _p(best), _p(second), _p(match), threads)
        return idx, best, second, match


class _Extractor:
    def __init__(self, o, nfeatures, scale, nlevels, ini, mn):
        self.o, self.nlevels, self.nfeatures = o, nlevels, nfeatures
        self.h = o._extractor_create(nfeatures, scale, nlevels, ini, mn)

    def __del__(self):
        if getattr(self, 'h', None):
            self.o._extractor_destroy(self.h)
            self.h = None

    def extract(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures + 64 * self.nlevels
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = self.o._extractor_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap)
        if n < 0:
            raise RuntimeError(f'oracle extract failed: {n}')
        return kps[:n].copy(), desc[:n].copy()

    def pyramid(self):
        out = []
        for s in range(self.nlevels):
            w, h = C.c_int(), C.c_int()
            assert self.o._extractor_level_size(self.h, s, C.byref(w), C.byref(h)) == 0
            a = np.empty((h.value, w.value), np.uint8)
            self.o._extractor_level_copy(self.h, s, _p(a), a.strides[0])
            out.append(a)
        return out

    def tables(self):
        t = [np.empty(self.nlevels, np.float32) for _ in range(4)]
        self.o._extractor_tables(self.h, *[_p(a) for a in t])
        return t


class OracleVocabulary:
    """ORBVocabulary (DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) of one oracle library."""

    def __init__(self, o, path=None, arrays=None):
        self.o = o
        if path is not None:
            self.h = o._voc_load_text(os.fsencode(path))
        else:
            a = arrays
            parent = np.ascontiguousarray(a['parent'], np.int32); leaf = np.ascontiguousarray(a['is_leaf'], np.uint8)
            desc = np.ascontiguousarray(a['desc'], np.uint8); w = np.ascontiguousarray(a['weights'], np.float64)
            self.h = o._voc_create(a['k'], a['L'], a['scoring'], a['weighting'], len(parent), _p(parent), _p(leaf), _p(desc), _p(w))
        if not self.h:
            raise ValueError('vocabulary rejected')

    def __del__(self):
        if getattr(self, 'h', None):
            self.o._voc_destroy(self.h)
            self.h = None

    def transform(self, desc, levelsup=4):
        """Returns (word_ids, word_vals, (fv_nodes, fv_start, fv_items))."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        wi = np.empty(n + 1, np.int32); wv = np.empty(n + 1, np.float64)
        fn = np.empty(n + 1, np.uint32); fs = np.empty(n + 2, np.int32); fi = np.empty(n + 1, np.uint32)
        nfv = C.c_int32(0)
        nw = self.o._bow_transform(self.h, _p(desc), n, levelsup, _p(wi), _p(wv), _p(fn), _p(fs), _p(fi), C.byref(nfv))
        m = nfv.value
        return wi[:nw].copy(), wv[:nw].copy(), (fn[:m].copy(), fs[:m + 1].copy(), fi[:fs[m]].copy())

    def descend(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        w = np.empty(len(desc), np.int32); nd = np.empty(len(desc), np.int32)
        self.o._bow_descend(self.h, _p(desc), len(desc), levelsup, _p(w), _p(nd))
        return w, nd

    def score(self, a, b):
        ia = np.ascontiguousarray(a[0], np.int32); va = np.ascontiguousarray(a[1], np.float64)
        ib = np.ascontiguousarray(b[0], np.int32); vb = np.ascontiguousarray(b[1], np.float64)
        return self.o._bow_score(self.h, _p(ia), _p(va), len(ia), _p(ib), _p(vb), len(ib))

    def time_transform(self, desc, levelsup=4, reps=5):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        return self.o._time_bow_transform(self.h, _p(desc), len(desc), levelsup, reps)
