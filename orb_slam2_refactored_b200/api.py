"""Python host-side mirror of the reference's interface for the ORB front-end hot path, on top of the C ABI
(include/orbx.h, liborbx_b200.so). Names and argument meaning follow the reference:

    ORB_SLAM2::ORBextractor            include/ORBextractor.h:34-80     -> ORBextractor
    ORB_SLAM2::ComputeStereoMatches    include/ORBmatcher.h:41-45       -> ComputeStereoMatches / StereoMatcher
    ORBmatcher::DescriptorDistance     include/ORBmatcher.h:54          -> ORBmatcher.DescriptorDistance
    best/second ratio-test inner loop  src/ORBmatcher.cc:477-507        -> ORBmatcher.knn2 / knn2_sharded

Everything computes on the GPU through the shared library; there is no CPU path. If the library is missing or
no sm_100 device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

KP_DTYPE = np.dtype([('x', '<f4'), ('y', '<f4'), ('size', '<f4'), ('angle', '<f4'), ('response', '<f4'),
                     ('octave', '<i4'), ('class_id', '<i4')])
assert KP_DTYPE.itemsize == 28

ORBX_OK, ORBX_ERR_INVALID, ORBX_ERR_CUDA, ORBX_ERR_CAPACITY, ORBX_ERR_STATE = range(5)

TH_HIGH = 100   # src/ORBmatcher.cc:41
TH_LOW = 50     # src/ORBmatcher.cc:42


class OrbxError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f'orbx status {status}: {msg}')
        self.status = status


class _Params(C.Structure):
    _fields_ = [('nfeatures', C.c_int32), ('scale_factor', C.c_float), ('nlevels', C.c_int32),
                ('ini_th_fast', C.c_int32), ('min_th_fast', C.c_int32)]


class _Camera(C.Structure):
    _fields_ = [(n, C.c_float) for n in ('fx', 'fy', 'cx', 'cy', 'bf', 'baseline')]


class _Bounds(C.Structure):
    _fields_ = [(n, C.c_float) for n in ('minx', 'maxx', 'miny', 'maxy')]


class _FrameView(C.Structure):
    _fields_ = [('n', C.c_int32), ('kps_un', C.c_void_p), ('desc', C.c_void_p), ('uright', C.c_void_p), ('bounds', _Bounds),
                ('nlevels', C.c_int32), ('scale_factors', C.c_void_p)]


class _FeatureVector(C.Structure):
    _fields_ = [('nnodes', C.c_int32), ('node_ids', C.c_void_p), ('start', C.c_void_p), ('indices', C.c_void_p)]


class _VocabularyDesc(C.Structure):
    _fields_ = [('k', C.c_int32), ('L', C.c_int32), ('scoring', C.c_int32), ('weighting', C.c_int32), ('nnodes', C.c_int64),
                ('parent', C.c_void_p), ('is_leaf', C.c_void_p), ('descriptors', C.c_void_p), ('weights', C.c_void_p)]


class _Sim3(C.Structure):
    _fields_ = [('R', C.c_float * 9), ('t', C.c_float * 3), ('s', C.c_float)]


class _Pose(C.Structure):
    _fields_ = [('R', C.c_float * 9), ('t', C.c_float * 3)]


TRACK_POINT_DTYPE = np.dtype([('proj_x', '<f4'), ('proj_y', '<f4'), ('proj_xr', '<f4'), ('view_cos', '<f4'), ('scale_level', '<i4'),
                              ('flags', '<i4')])
LAST_POINT_DTYPE = np.dtype([('xw', '<f4', (3,)), ('octave', '<i4'), ('angle', '<f4'), ('flags', '<i4')])
SIM3_POINT_DTYPE = np.dtype([('xw', '<f4', (3,)), ('normal', '<f4', (3,)), ('min_distance', '<f4'), ('max_distance', '<f4'), ('flags', '<i4')])
KF_POINT_DTYPE = np.dtype([('xw', '<f4', (3,)), ('min_distance', '<f4'), ('max_distance', '<f4'), ('angle', '<f4'), ('flags', '<i4')])
assert TRACK_POINT_DTYPE.itemsize == 24 and LAST_POINT_DTYPE.itemsize == 24 and KF_POINT_DTYPE.itemsize == 28
GRID_COLS, GRID_ROWS = 64, 48   # include/Frame.h:72-73

_lib = None

_SIGNATURES = {
    'orbx_last_error': (C.c_char_p, []),
    'orbx_device_count': (C.c_int, []),
    'orbx_create': (C.c_int, [C.POINTER(_Params), C.c_int, C.POINTER(C.c_void_p)]),
    'orbx_destroy': (C.c_int, [C.c_void_p]),
    'orbx_get_params': (C.c_int, [C.c_void_p, C.POINTER(_Params)]),
    'orbx_scale_tables': (C.c_int, [C.c_void_p] + [C.c_void_p] * 4),
    'orbx_feature_quotas': (C.c_int, [C.c_void_p, C.c_void_p]),
    'orbx_extract': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    'orbx_extract_batch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_void_p]),
    'orbx_extract_batch_device': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_void_p]),
    'orbx_debug_cos_sin': (C.c_int, [C.c_int, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p]),
    'orbx_max_keypoints': (C.c_int, [C.c_void_p]),
    'orbx_plan': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    'orbx_last_result_shape': (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'orbx_synchronize': (C.c_int, [C.c_void_p]),
    'orbx_stream': (C.c_void_p, [C.c_void_p]),
    'orbx_enable_stage_timing': (C.c_int, [C.c_void_p, C.c_int]),
    'orbx_stage_times': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    'orbx_level_size': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'orbx_pyramid_level': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    'orbx_pyramid_level_device': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    'orbx_debug_candidates': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    'orbx_debug_selected': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    'orbx_debug_blurred_level': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    'orbx_descriptor_distance': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'orbx_stereo_match': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(_Camera), C.c_void_p, C.c_void_p]),
    'orbx_stereo_match_device': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(_Camera), C.c_void_p, C.c_void_p]),
    'orbx_stereo_match_host': (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.POINTER(_Camera), C.c_void_p, C.c_void_p]),
    'orbx_knn2': (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p]),
    'orbx_knn2_device': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    'orbx_knn2_partial_device': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    'orbx_knn2_merge_device': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    'orbx_knn2_sharded': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'orbx_knn2_work_parts': (C.c_int, [C.c_int64, C.c_int64]),
    'orbx_convert_to_gray': (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    'orbx_extract_batch_color': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'orbx_stereo_from_rgbd': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(_Camera),
                                        C.c_void_p, C.c_void_p]),
    'orbx_distinctive_descriptors': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'orbx_measure_popc_peak': (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    'orbx_undistort_keypoints': (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.POINTER(_Camera), C.c_void_p, C.c_int, C.c_void_p]),
    'orbx_remap': (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int,
                             C.c_size_t]),
    'orbx_set_rectification': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int]),
    'orbx_rectify_batch_device': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t]),
    'orbx_extract_batch_rectified': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_void_p]),
    'orbx_frame_create': (C.c_int, [C.POINTER(_FrameView), C.c_int, C.POINTER(C.c_void_p)]),
    'orbx_frame_destroy': (C.c_int, [C.c_void_p]),
    'orbx_frame_grid': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    'orbx_frame_features_in_area': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    'orbx_search_by_projection_local_map': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                                      C.POINTER(C.c_int)]),
    'orbx_search_by_projection_last_frame': (C.c_int, [C.c_void_p, C.POINTER(_Camera), C.POINTER(_Pose), C.POINTER(_Pose), C.c_void_p,
                                                       C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    'orbx_search_for_initialization': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int,
                                                 C.POINTER(C.c_int)]),
    'orbx_frame_last_stats': (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_void_p]),
    'orbx_search_by_projection_keyframe': (C.c_int, [C.c_void_p, C.POINTER(_Camera), C.POINTER(_Pose), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_int, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    'orbx_search_by_projection_sim3': (C.c_int, [C.c_void_p, C.POINTER(_Camera), C.POINTER(_Sim3), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.c_int, C.POINTER(C.c_int)]),
    'orbx_search_windows': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    'orbx_search_by_bow': (C.c_int, [C.c_void_p, C.POINTER(_FeatureVector), C.c_void_p, C.c_void_p, C.POINTER(_FeatureVector), C.c_void_p, C.c_float,
                                     C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    'orbx_frame_assign_device': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_Bounds), C.c_int, C.c_void_p]),
    'orbx_undistort_keypoints_device': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_Camera), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'orbx_frame_assign': (C.c_int, [C.c_void_p, C.POINTER(_FrameView)]),
    'orbx_search_best_in_windows': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'orbx_fuse': (C.c_int, [C.c_void_p, C.POINTER(_Camera), C.POINTER(_Pose), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                            C.c_void_p, C.c_void_p]),
    'orbx_fuse_sim3': (C.c_int, [C.c_void_p, C.POINTER(_Camera), C.POINTER(_Sim3), C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                 C.c_void_p, C.c_void_p]),
    'orbx_search_by_sim3': (C.c_int, [C.c_void_p, C.POINTER(_Camera), C.POINTER(_Pose), C.c_float, C.c_void_p, C.POINTER(_Camera), C.POINTER(_Pose),
                                      C.c_float, C.POINTER(_Sim3), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_int)]),
    'orbx_search_for_triangulation': (C.c_int, [C.c_void_p, C.POINTER(_FeatureVector), C.c_void_p, C.c_void_p, C.POINTER(_FeatureVector), C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    'orbx_vocabulary_create': (C.c_int, [C.POINTER(_VocabularyDesc), C.c_int, C.POINTER(C.c_void_p)]),
    'orbx_vocabulary_load_text': (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    'orbx_vocabulary_info': (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_int64)] * 2),
    'orbx_vocabulary_destroy': (C.c_int, [C.c_void_p]),
    'orbx_bow_transform': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]),
    'orbx_bow_transform_batch_device': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 9),
    'orbx_bow_score_l1': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
}


def exported_symbols():
    """Every entry point include/orbx.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def library_path():
    return _build.LIB


def lib():
    """Loads liborbx_b200.so; raises if it has not been built (there is no fallback implementation)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ImportError(f'{path} is missing: run `python -m orb_slam2_refactored_b200.build` (needs nvcc). '
                              'The ORB front-end has no CPU fallback.')
        l = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def _check(status):
    if status != ORBX_OK:
        raise OrbxError(status, lib().orbx_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _addr(a):
    """Same as _p for a numpy array the caller keeps alive, without building the .ctypes helper object (the one-frame call path)."""
    return C.c_void_p(a.__array_interface__['data'][0])


def device_count():
    return lib().orbx_device_count()


class ORBextractor:
    """ORB_SLAM2::ORBextractor (include/ORBextractor.h:34-80). Same parameters, getters and Extract semantics; plus
    batch and device-resident entry points that the reference does not have."""

    class Parameters:   # include/ORBextractor.h:38-47
        def __init__(self, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7):
            self.nfeatures, self.scaleFactor, self.nlevels = nfeatures, scaleFactor, nlevels
            self.iniThFAST, self.minThFAST = iniThFAST, minThFAST

    def __init__(self, param=None, device=0, **kw):
        if param is None:
            param = ORBextractor.Parameters(**kw)
        self.param_ = param
        self.device = device
        self._h = C.c_void_p()
        p = _Params(param.nfeatures, param.scaleFactor, param.nlevels, param.iniThFAST, param.minThFAST)
        _check(lib().orbx_create(C.byref(p), device, C.byref(self._h)))
        n = param.nlevels
        t = [np.empty(n, np.float32) for _ in range(4)]
        _check(lib().orbx_scale_tables(self._h, *[_p(a) for a in t]))
        self._scale, self._inv_scale, self._sigma_sq, self._inv_sigma_sq = t
        self._last_frames = 0

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value and _lib is not None:
            _lib.orbx_destroy(h)
            self._h = C.c_void_p()

    # ---- getters, include/ORBextractor.h:57-63
    def GetLevels(self): return self.param_.nlevels
    def GetScaleFactor(self): return np.float32(self.param_.scaleFactor)
    def GetScaleFactors(self): return self._scale
    def GetInverseScaleFactors(self): return self._inv_scale
    def GetScaleSigmaSquares(self): return self._sigma_sq
    def GetInverseScaleSigmaSquares(self): return self._inv_sigma_sq

    def GetFeatureQuotas(self):
        q = np.empty(self.param_.nlevels, np.int32)
        _check(lib().orbx_feature_quotas(self._h, _p(q)))
        return q

    def max_keypoints(self):
        return lib().orbx_max_keypoints(self._h)

    def GetImagePyramid(self, frame=0):
        """Levels of the last Extract, downloaded on request (they live on the device)."""
        out = []
        for s in range(self.param_.nlevels):
            w, h = C.c_int(), C.c_int()
            _check(lib().orbx_level_size(self._h, s, C.byref(w), C.byref(h)))
            a = np.empty((h.value, w.value), np.uint8)
            _check(lib().orbx_pyramid_level(self._h, frame, s, _p(a), a.strides[0]))
            out.append(a)
        return out

    # ---- Extract, src/ORBextractor.cc:743-820
    def Extract(self, image):
        """image: (H, W) uint8. Returns (keypoints[KP_DTYPE], descriptors[N,32] uint8)."""
        k, d = self.ExtractBatch(np.asarray(image)[None])
        return k[0], d[0]

    def ExtractBatch(self, images):
        """images: (F, H, W) uint8 host array. Returns two lists of per-frame arrays."""
        images = np.ascontiguousarray(images, np.uint8)
        if images.ndim != 3:
            raise OrbxError(ORBX_ERR_INVALID, 'CV_Assert(image.type() == CV_8U): expected (F, H, W) uint8')   # :457
        F, H, W = images.shape
        cap = lib().orbx_max_keypoints(self._h)
        # rows past n[f] are never returned, so the arrays need no zero fill; a call owns its arrays, so the results are views of them
        kps = np.empty((F, cap), KP_DTYPE)
        desc = np.empty((F, cap, 32), np.uint8)
        n = np.zeros(F, np.int32)
        _check(lib().orbx_extract_batch(self._h, _addr(images), F, W, H, images.strides[1], images.strides[0], _addr(kps), _addr(desc), cap, _addr(n)))
        self._last_frames = F
        return [kps[f, :n[f]] for f in range(F)], [desc[f, :n[f]] for f in range(F)]

    def ExtractBatchColor(self, images, RGB=True):
        """images: (F, H, W, 3|4) uint8. ConvertToGray (src/System.cc:122-137) fused into the upload, then Extract."""
        images = np.ascontiguousarray(images, np.uint8)
        F, H, W, ch = images.shape
        cap = lib().orbx_max_keypoints(self._h)
        kps = np.zeros((F, cap), KP_DTYPE); desc = np.zeros((F, cap, 32), np.uint8); n = np.zeros(F, np.int32)
        _check(lib().orbx_extract_batch_color(self._h, _p(images), F, W, H, images.strides[1], images.strides[0], ch, int(RGB), _p(kps),
                                              _p(desc), cap, _p(n)))
        self._last_frames = F
        return [kps[f, :n[f]].copy() for f in range(F)], [desc[f, :n[f]].copy() for f in range(F)]

    def SetRectification(self, map1, map2, src_shape):
        """Keeps the maps of cv::initUndistortRectifyMap (Examples/Stereo/stereo_euroc.cc:88-89) on the device; src_shape = (h, w) of the raw
        frames. ExtractBatchRectified then remaps and extracts in one pass."""
        map1 = np.ascontiguousarray(map1, np.float32); map2 = np.ascontiguousarray(map2, np.float32)
        _check(lib().orbx_set_rectification(self._h, _p(map1), _p(map2), map1.strides[0], map1.shape[1], map1.shape[0], int(src_shape[1]),
                                            int(src_shape[0])))

    def ExtractBatchRectified(self, images):
        """images: (F, h, w) raw (unrectified) uint8 frames. cv::remap (Examples/Stereo/stereo_euroc.cc:100-101) fused into the upload,
        then Extract on the rectified image. Returns two lists of per-frame arrays like ExtractBatch."""
        images = np.ascontiguousarray(images, np.uint8)
        F, H, W = images.shape
        cap = lib().orbx_max_keypoints(self._h)
        kps = np.zeros((F, cap), KP_DTYPE); desc = np.zeros((F, cap, 32), np.uint8); n = np.zeros(F, np.int32)
        _check(lib().orbx_extract_batch_rectified(self._h, _p(images), F, W, H, images.strides[1], images.strides[0], _p(kps), _p(desc), cap,
                                                  _p(n)))
        self._last_frames = F
        return [kps[f, :n[f]].copy() for f in range(F)], [desc[f, :n[f]].copy() for f in range(F)]

    def extract_batch_device(self, d_images, d_kps=None, d_desc=None, d_n=None):
        """d_images: CUDA torch.uint8 tensor (F, H, W) (row stride free). Outputs are torch tensors on the same device:
        kps (F, cap, 7) float32 viewable as KP_DTYPE, desc (F, cap, 32) uint8, n (F,) int32. Asynchronous on the
        handle's stream; call synchronize() before reading."""
        import torch
        F, H, W = d_images.shape
        assert d_images.dtype == torch.uint8 and d_images.is_cuda and d_images.stride(2) == 1
        # the library's stream is not torch's: order it behind whatever produced d_images (and the output tensors' last use)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(d_images.device))
        torch.cuda.ExternalStream(self.stream(), device=d_images.device).wait_event(ev)
        if d_kps is None:
            # plan first so that cap is the exact per-size bound
            _check(lib().orbx_plan(self._h, W, H, F))
            cap = max(self.max_keypoints(), 1)
            d_kps = torch.empty((F, cap, 7), dtype=torch.float32, device=d_images.device)
            d_desc = torch.empty((F, cap, 32), dtype=torch.uint8, device=d_images.device)
            d_n = torch.empty((F,), dtype=torch.int32, device=d_images.device)
        cap = d_kps.shape[1]
        _check(lib().orbx_extract_batch_device(self._h, C.c_void_p(d_images.data_ptr()), F, W, H, d_images.stride(1), d_images.stride(0),
                                               C.c_void_p(d_kps.data_ptr()), C.c_void_p(d_desc.data_ptr()), cap, C.c_void_p(d_n.data_ptr())))
        self._last_frames = F
        return d_kps, d_desc, d_n

    STAGES = ('pyramid', 'fast', 'quadtree', 'blur', 'describe')

    def enable_stage_timing(self, on=True):
        _check(lib().orbx_enable_stage_timing(self._h, int(on)))

    def stage_times(self):
        """(dict stage -> summed ms, number of extract calls) since the last query; waits for the stream."""
        ms = np.zeros(5, np.float32); calls = C.c_int()
        _check(lib().orbx_stage_times(self._h, _p(ms), C.byref(calls)))
        return dict(zip(self.STAGES, [float(v) for v in ms])), calls.value

    def level_sizes(self):
        out = []
        for s in range(self.param_.nlevels):
            w, h = C.c_int(), C.c_int()
            _check(lib().orbx_level_size(self._h, s, C.byref(w), C.byref(h)))
            out.append((w.value, h.value))
        return out

    def synchronize(self):
        _check(lib().orbx_synchronize(self._h))

    def stream(self):
        return lib().orbx_stream(self._h)

    # ---- stage probes for parity tests
    def debug_candidates(self, frame, level):
        cap = 1 << 16
        while True:
            out = np.empty((cap, 3), np.int32); n = C.c_int()
            st = lib().orbx_debug_candidates(self._h, frame, level, _p(out), cap, C.byref(n))
            if st == ORBX_ERR_CAPACITY:
                cap = n.value + 16
                continue
            _check(st)
            return out[:n.value].copy()

    def debug_selected(self, frame, level):
        cap = self.max_keypoints() + 64
        out = np.empty((cap, 3), np.int32); n = C.c_int()
        _check(lib().orbx_debug_selected(self._h, frame, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def debug_blurred(self, frame, level):
        w, h = C.c_int(), C.c_int()
        _check(lib().orbx_level_size(self._h, level, C.byref(w), C.byref(h)))
        a = np.empty((h.value, w.value), np.uint8)
        _check(lib().orbx_debug_blurred_level(self._h, frame, level, _p(a), a.strides[0]))
        return a


def ComputeStereoMatches(keypointsL, descriptorsL, pyramidL, keypointsR, descriptorsR, pyramidR, scaleFactors, invScaleFactors,
                         camera, device=0):
    """ORB_SLAM2::ComputeStereoMatches (include/ORBmatcher.h:41-45, src/ORBmatcher.cc:72-247), the reference's
    argument list with host data. camera = (fx, fy, cx, cy, bf, baseline). Returns (uright, depth)."""
    n = len(pyramidL)
    pyramidL = [np.ascontiguousarray(p, np.uint8) for p in pyramidL]
    pyramidR = [np.ascontiguousarray(p, np.uint8) for p in pyramidR]
    pl = (C.c_void_p * n)(*[p.ctypes.data for p in pyramidL])
    pr = (C.c_void_p * n)(*[p.ctypes.data for p in pyramidR])
    lw = np.array([p.shape[1] for p in pyramidL], np.int32)
    lh = np.array([p.shape[0] for p in pyramidL], np.int32)
    lp = np.array([p.strides[0] for p in pyramidL], np.uint64)
    kl = np.ascontiguousarray(keypointsL, KP_DTYPE); kr = np.ascontiguousarray(keypointsR, KP_DTYPE)
    dl = np.ascontiguousarray(descriptorsL, np.uint8); dr = np.ascontiguousarray(descriptorsR, np.uint8)
    sc = np.ascontiguousarray(scaleFactors, np.float32); isc = np.ascontiguousarray(invScaleFactors, np.float32)
    ur = np.full(len(kl), -1, np.float32); dp = np.full(len(kl), -1, np.float32)
    cam = _Camera(*[float(v) for v in camera])
    _check(lib().orbx_stereo_match_host(device, _p(kl), len(kl), _p(dl), pl, _p(kr), len(kr), _p(dr), pr, _p(lw), _p(lh), _p(lp), n,
                                        _p(sc), _p(isc), C.byref(cam), _p(ur), _p(dp)))
    return ur, dp


def ComputeStereoMatchesResident(extractorL, extractorR, camera):
    """The same on the device-resident results of the last ExtractBatch of two extractors (what SystemImpl::TrackStereo
    does at src/System.cc:449-461, without moving keypoints, descriptors or pyramids off the GPU).
    Returns (uright, depth) as (F, cap) arrays; entries past a frame's keypoint count are undefined."""
    f, c = C.c_int(), C.c_int()
    _check(lib().orbx_last_result_shape(extractorL._h, C.byref(f), C.byref(c)))
    F, cap = f.value, c.value         # the arrays are laid out like the keypoints of that extract
    ur = np.full((F, cap), -1, np.float32); dp = np.full((F, cap), -1, np.float32)
    cam = _Camera(*[float(v) for v in camera])
    _check(lib().orbx_stereo_match(extractorL._h, extractorR._h, C.byref(cam), _p(ur), _p(dp)))
    return ur, dp


class Frame:
    """What the guided matchers need of ORB_SLAM2::Frame (include/Frame.h:83-168), resident on one GPU: keypointsUn, descriptors,
    uright, imageBounds, the pyramid's scale factors and the FeaturesGrid (built on the device at construction, src/Frame.cc:70-100).
    `mappoints` mirrors frame.mappoints as codes: -1 null, >= 0 index into the point list of the last search, -2 / -3 some other map
    point with / without observations."""

    def __init__(self, keypointsUn, descriptors, scaleFactors, imageBounds, uright=None, device=0):
        self._h = None
        view = self._view(keypointsUn, descriptors, scaleFactors, imageBounds, uright)
        h = C.c_void_p()
        _check(lib().orbx_frame_create(C.byref(view), device, C.byref(h)))
        self._h = h

    def _view(self, keypointsUn, descriptors, scaleFactors, imageBounds, uright):
        self.keypointsUn = np.ascontiguousarray(keypointsUn).view(KP_DTYPE)
        self.descriptors = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        self.uright = None if uright is None else np.ascontiguousarray(uright, np.float32)
        self.scaleFactors = np.ascontiguousarray(scaleFactors, np.float32)
        self.imageBounds = tuple(float(np.float32(b)) for b in imageBounds)     # (minx, maxx, miny, maxy)
        self.N = len(self.keypointsUn)
        self.mappoints = np.full(self.N, -1, np.int32)
        return _FrameView(self.N, self.keypointsUn.ctypes.data, self.descriptors.ctypes.data,
                          None if self.uright is None else self.uright.ctypes.data, _Bounds(*self.imageBounds), len(self.scaleFactors),
                          self.scaleFactors.ctypes.data)

    def assign(self, keypointsUn, descriptors, scaleFactors, imageBounds, uright=None):
        """The next image's Frame in the same device buffers (Tracking constructs one Frame per image)."""
        view = self._view(keypointsUn, descriptors, scaleFactors, imageBounds, uright)
        _check(lib().orbx_frame_assign(self._h, C.byref(view)))

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and _lib is not None:
            _lib.orbx_frame_destroy(h)
            self._h = None

    def assign_device(self, d_kps_un, d_desc, n, scaleFactors, imageBounds, d_uright=None):
        """The next Frame from device-resident keypoints / descriptors (torch tensors or anything with data_ptr()), e.g. the outputs of
        ORBextractor.extract_batch_device. `mappoints` is reset to all -1; keypointsUn / descriptors are not mirrored on the host."""
        sf = np.ascontiguousarray(scaleFactors, np.float32)
        b = _Bounds(*[float(np.float32(v)) for v in imageBounds])
        _check(lib().orbx_frame_assign_device(self._h, C.c_void_p(d_kps_un.data_ptr()), C.c_void_p(d_desc.data_ptr()),
                                              None if d_uright is None else C.c_void_p(d_uright.data_ptr()), int(n), C.byref(b), len(sf), _p(sf)))
        self.N = int(n)
        self.scaleFactors = sf
        self.imageBounds = tuple(float(np.float32(v)) for v in imageBounds)
        self.mappoints = np.full(self.N, -1, np.int32)
        self.keypointsUn = None; self.descriptors = None; self.uright = None

    def grid(self):
        """(cell_start[64*48+1], items): grid_[cx][cy] is items[cell_start[cx*48+cy] : cell_start[cx*48+cy+1]]."""
        start = np.empty(GRID_COLS * GRID_ROWS + 1, np.int32)
        items = np.empty(max(self.N, 1), np.int32)
        n = C.c_int()
        _check(lib().orbx_frame_grid(self._h, _p(start), _p(items), len(items), C.byref(n)))
        return start, items[:n.value].copy()

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        return self.GetFeaturesInAreaBatch([(x, y, r, minLevel, maxLevel)])[0]

    def GetFeaturesInAreaBatch(self, queries):
        """queries: (x, y, r, minLevel, maxLevel) rows -> list of index arrays in the reference's output order (src/Frame.cc:102-145)."""
        q = np.asarray(queries, np.float64).reshape(-1, 5)
        xyr = np.ascontiguousarray(q[:, :3], np.float32)
        lv = np.ascontiguousarray(q[:, 3:], np.int32)
        off = np.empty(len(q) + 1, np.int32)
        idx = np.empty(max(self.N * 4, 1024), np.int32)
        st = lib().orbx_frame_features_in_area(self._h, _p(xyr), _p(lv), len(q), _p(off), _p(idx), len(idx))
        if st == ORBX_ERR_CAPACITY:
            idx = np.empty(int(off[-1]), np.int32)
            st = lib().orbx_frame_features_in_area(self._h, _p(xyr), _p(lv), len(q), _p(off), _p(idx), len(idx))
        _check(st)
        return [idx[off[i]:off[i + 1]].copy() for i in range(len(q))]

    def last_rounds(self):
        return self.last_stats()[0]

    def last_stats(self):
        """(rounds, kernel milliseconds) of the last search on this frame."""
        n, ms = C.c_int(), C.c_float()
        _check(lib().orbx_frame_last_stats(self._h, C.byref(n), C.byref(ms), None))
        return n.value, ms.value

    def last_phase_us(self):
        """Microseconds of the last search's phases: windows, count + scan, fill, distances, rounds, finalisation."""
        ph = np.zeros(7, np.float32)
        _check(lib().orbx_frame_last_stats(self._h, None, None, _p(ph)))
        return ph


def _pose(p):
    R, t = p
    P = _Pose()
    P.R[:] = [float(x) for x in np.asarray(R, np.float32).reshape(9)]
    P.t[:] = [float(x) for x in np.asarray(t, np.float32).reshape(3)]
    return P


class ORBmatcher:
    """The pieces of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:47-103) that are on the hot path: the Hamming kernels and the guided
    window searches of Tracking."""

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.fNNRatio_ = nnratio
        self.checkOrientation_ = checkOri
        self.device = device

    def SearchByProjection(self, frame, mappoints, descriptors, th=3.0):
        """SearchByProjection(Frame&, const std::vector<MapPoint*>&, float th) — src/ORBmatcher.cc:315-382. mappoints: TRACK_POINT_DTYPE
        records (the track* members of each MapPoint), descriptors: their GetDescriptor() rows. Updates frame.mappoints; returns nmatches."""
        pts = np.ascontiguousarray(mappoints).view(TRACK_POINT_DTYPE)
        desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        n = C.c_int()
        _check(lib().orbx_search_by_projection_local_map(frame._h, _p(frame.mappoints), _p(pts), _p(desc), len(pts), th, self.fNNRatio_,
                                                         C.byref(n)))
        return n.value

    def SearchByProjectionLastFrame(self, currFrame, camera, currPose, lastPose, lastPoints, descriptors, th, monocular):
        """SearchByProjection(Frame& currFrame, const Frame& lastFrame, float th, bool monocular) — src/ORBmatcher.cc:1279-1362.
        lastPoints: LAST_POINT_DTYPE records, one per keypoint of the last frame; poses are (R 3x3, t 3) of CameraPose."""
        pts = np.ascontiguousarray(lastPoints).view(LAST_POINT_DTYPE)
        desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        n = C.c_int()
        cam = _Camera(*[float(c) for c in camera])
        cp, lp = _pose(currPose), _pose(lastPose)
        _check(lib().orbx_search_by_projection_last_frame(currFrame._h, C.byref(cam), C.byref(cp), C.byref(lp), _p(currFrame.mappoints),
                                                          _p(pts), _p(desc), len(pts), th, int(bool(monocular)),
                                                          int(bool(self.checkOrientation_)), C.byref(n)))
        return n.value

    def SearchByProjectionKeyFrame(self, frame, camera, pose, logScaleFactor, keyframePoints, descriptors, th, ORBdist):
        """SearchByProjection(Frame&, KeyFrame*, alreadyFound, th, ORBdist) — src/ORBmatcher.cc:1364-1447 (relocalisation).
        keyframePoints: KF_POINT_DTYPE records, one per keypoint of the key frame. Updates frame.mappoints; returns nmatches."""
        pts = np.ascontiguousarray(keyframePoints).view(KF_POINT_DTYPE)
        desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        n = C.c_int()
        cam = _Camera(*[float(c) for c in camera])
        P = _pose(pose)
        _check(lib().orbx_search_by_projection_keyframe(frame._h, C.byref(cam), C.byref(P), float(np.float32(logScaleFactor)), _p(frame.mappoints),
                                                        _p(pts), _p(desc), len(pts), th, int(ORBdist), int(bool(self.checkOrientation_)),
                                                        C.byref(n)))
        return n.value

    def SearchByProjectionSim3(self, keyframe, camera, Scw, logScaleFactor, mappoints, descriptors, th):
        """SearchByProjection(keyframe, Scw, mappoints, matched, th) — src/ORBmatcher.cc:518-612 (loop closing). Scw = (R 3x3, t 3, s);
        mappoints: SIM3_POINT_DTYPE records. keyframe.mappoints plays `matched` (in/out). Returns nmatches."""
        pts = np.ascontiguousarray(mappoints).view(SIM3_POINT_DTYPE)
        desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        S = _Sim3()
        S.R[:] = [float(x) for x in np.asarray(Scw[0], np.float32).reshape(9)]
        S.t[:] = [float(x) for x in np.asarray(Scw[1], np.float32).reshape(3)]
        S.s = float(np.float32(Scw[2]))
        n = C.c_int()
        cam = _Camera(*[float(c) for c in camera])
        _check(lib().orbx_search_by_projection_sim3(keyframe._h, C.byref(cam), C.byref(S), float(np.float32(logScaleFactor)), _p(keyframe.mappoints),
                                                    _p(pts), _p(desc), len(pts), int(th), C.byref(n)))
        return n.value

    def SearchByBoW(self, keyframe, featureVector1, valid1, frame, featureVector2, valid2=None):
        """SearchByBoW(KeyFrame*, Frame&, matches) — src/ORBmatcher.cc:452-516 — when valid2 is None, SearchByBoW(KeyFrame*, KeyFrame*,
        matches12) — :696-766 — otherwise. Feature vectors are (node_ids, start, indices) CSR triples of DBoW2::FeatureVector; valid*:
        per keypoint, map point present and not bad. Returns (nmatches, match2) with match2[idx2] = matched keypoint of `keyframe` or -1."""
        def fv(t):
            ids = np.ascontiguousarray(t[0], np.uint32); start = np.ascontiguousarray(t[1], np.int32); idx = np.ascontiguousarray(t[2], np.uint32)
            return _FeatureVector(len(ids), ids.ctypes.data, start.ctypes.data, idx.ctypes.data), (ids, start, idx)
        c1, k1 = fv(featureVector1)
        c2, k2 = fv(featureVector2)
        v1 = np.ascontiguousarray(valid1, np.uint8)
        v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
        m2 = np.empty(max(frame.N, 1), np.int32)
        n = C.c_int()
        _check(lib().orbx_search_by_bow(keyframe._h, C.byref(c1), _p(v1), frame._h, C.byref(c2), None if v2 is None else _p(v2), self.fNNRatio_,
                                        int(bool(self.checkOrientation_)), _p(m2), C.byref(n)))
        return n.value, m2[:frame.N]

    @staticmethod
    def _sim3(S):
        out = _Sim3()
        out.R[:] = [float(x) for x in np.asarray(S[0], np.float32).reshape(9)]
        out.t[:] = [float(x) for x in np.asarray(S[1], np.float32).reshape(3)]
        out.s = float(np.float32(S[2]))
        return out

    def FuseSearch(self, keyframe, camera, pose, logScaleFactor, invSigmaSq, mappoints, descriptors, th=3.0):
        """The search half of Fuse(keyframe, mappoints, th) — src/ORBmatcher.cc:868-954. mappoints: SIM3_POINT_DTYPE records (flags bit 0 =
        non-null and not bad). Returns (bestIdx, bestDist) per point; the caller replays :876 and :956-976 in order (see FuseApply)."""
        pts = np.ascontiguousarray(mappoints).view(SIM3_POINT_DTYPE)
        desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        sig = np.ascontiguousarray(invSigmaSq, np.float32)
        bi = np.empty(max(len(pts), 1), np.int32); bd = np.empty(max(len(pts), 1), np.int32)
        cam = _Camera(*[float(c) for c in camera]); P = _pose(pose)
        _check(lib().orbx_fuse(keyframe._h, C.byref(cam), C.byref(P), float(np.float32(logScaleFactor)), _p(sig), _p(pts), _p(desc), len(pts),
                               float(th), _p(bi), _p(bd)))
        return bi[:len(pts)], bd[:len(pts)]

    def FuseSim3Search(self, keyframe, camera, Scw, logScaleFactor, mappoints, descriptors, th=4.0):
        """The search half of Fuse(keyframe, Scw, mappoints, th, replacePoints) — src/ORBmatcher.cc:982-1067."""
        pts = np.ascontiguousarray(mappoints).view(SIM3_POINT_DTYPE)
        desc = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        bi = np.empty(max(len(pts), 1), np.int32); bd = np.empty(max(len(pts), 1), np.int32)
        cam = _Camera(*[float(c) for c in camera]); S = self._sim3(Scw)
        _check(lib().orbx_fuse_sim3(keyframe._h, C.byref(cam), C.byref(S), float(np.float32(logScaleFactor)), _p(pts), _p(desc), len(pts), float(th),
                                    _p(bi), _p(bd)))
        return bi[:len(pts)], bd[:len(pts)]

    def SearchBySim3(self, keyframe1, camera1, pose1, logScaleFactor1, keyframe2, camera2, pose2, logScaleFactor2, S12, th, mappoints1,
                     descriptors1, mappoints2, descriptors2):
        """SearchBySim3(kf1, kf2, matches12, S12, th) — src/ORBmatcher.cc:1090-1277. mappoints1/2: KF_POINT_DTYPE records, one per keypoint
        (flags bit 0 = present, not already matched, not bad). Returns (nfound, matches12, match1, match2): matches12[i1] = keypoint of kf2."""
        p1 = np.ascontiguousarray(mappoints1).view(KF_POINT_DTYPE); p2 = np.ascontiguousarray(mappoints2).view(KF_POINT_DTYPE)
        d1 = np.ascontiguousarray(descriptors1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(descriptors2, np.uint8).reshape(-1, 32)
        assert len(p1) == keyframe1.N and len(p2) == keyframe2.N
        m1 = np.empty(max(keyframe1.N, 1), np.int32); m2 = np.empty(max(keyframe2.N, 1), np.int32); m12 = np.empty(max(keyframe1.N, 1), np.int32)
        n = C.c_int()
        c1 = _Camera(*[float(c) for c in camera1]); c2 = _Camera(*[float(c) for c in camera2])
        P1 = _pose(pose1); P2 = _pose(pose2); S = self._sim3(S12)
        _check(lib().orbx_search_by_sim3(keyframe1._h, C.byref(c1), C.byref(P1), float(np.float32(logScaleFactor1)), keyframe2._h, C.byref(c2),
                                         C.byref(P2), float(np.float32(logScaleFactor2)), C.byref(S), float(th), _p(p1), _p(d1), _p(p2), _p(d2),
                                         _p(m1), _p(m2), _p(m12), C.byref(n)))
        return n.value, m12[:keyframe1.N], m1[:keyframe1.N], m2[:keyframe2.N]

    def SearchForTriangulation(self, keyframe1, featureVector1, hasMapPoint1, keyframe2, featureVector2, hasMapPoint2, F12, epipole2, sigmaSq2,
                               onlyStereo=False):
        """SearchForTriangulation(kf1, kf2, F12, matchIds, onlyStereo) — src/ORBmatcher.cc:768-866. Returns (nmatches, matches12); matchIds of
        the reference = [(idx1, matches12[idx1]) for idx1 in ascending order if matches12[idx1] >= 0]."""
        def fv(t):
            ids = np.ascontiguousarray(t[0], np.uint32); start = np.ascontiguousarray(t[1], np.int32); idx = np.ascontiguousarray(t[2], np.uint32)
            return _FeatureVector(len(ids), ids.ctypes.data, start.ctypes.data, idx.ctypes.data), (ids, start, idx)
        c1, k1 = fv(featureVector1)
        c2, k2 = fv(featureVector2)
        h1 = np.ascontiguousarray(hasMapPoint1, np.uint8); h2 = np.ascontiguousarray(hasMapPoint2, np.uint8)
        F = np.ascontiguousarray(F12, np.float32).reshape(9); ep = np.ascontiguousarray(epipole2, np.float32).reshape(2)
        sg = np.ascontiguousarray(sigmaSq2, np.float32)
        m12 = np.empty(max(keyframe1.N, 1), np.int32)
        n = C.c_int()
        _check(lib().orbx_search_for_triangulation(keyframe1._h, C.byref(c1), _p(h1), keyframe2._h, C.byref(c2), _p(h2), _p(F), _p(ep), _p(sg),
                                                   int(bool(onlyStereo)), int(bool(self.checkOrientation_)), _p(m12), C.byref(n)))
        return n.value, m12[:keyframe1.N]

    def SearchForInitialization(self, frame1, frame2, prevMatched, windowSize=10):
        """src/ORBmatcher.cc:614-694. prevMatched: (N1, 2) float32, updated in place; returns (nmatches, matches12)."""
        prev = np.ascontiguousarray(prevMatched, np.float32).reshape(-1, 2)
        m12 = np.empty(max(frame1.N, 1), np.int32)
        n = C.c_int()
        _check(lib().orbx_search_for_initialization(frame1._h, frame2._h, _p(prev), _p(m12), int(windowSize), self.fNNRatio_,
                                                    int(bool(self.checkOrientation_)), C.byref(n)))
        if prev is not prevMatched:
            np.copyto(np.asarray(prevMatched).reshape(-1, 2), prev, casting='unsafe')
        return n.value, m12[:frame1.N]

    @staticmethod
    def DescriptorDistance(a, b, device=0):
        """src/ORBmatcher.cc:1449-1457. a, b: 32-byte descriptors, or (n, 32) arrays for n pairs."""
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        single = a.ndim == 1
        a2 = a.reshape(-1, 32); b2 = b.reshape(-1, 32)
        out = np.empty(len(a2), np.int32)
        _check(lib().orbx_descriptor_distance(device, _p(a2), _p(b2), len(a2), _p(out)))
        return int(out[0]) if single else out

    def knn2(self, query, train, th_low=TH_LOW):
        """Best/second scan + ratio test (src/ORBmatcher.cc:477-507) of every query row against all train rows.
        Returns (idx, best, second, match)."""
        query = np.ascontiguousarray(query, np.uint8); train = np.ascontiguousarray(train, np.uint8)
        nq = len(query)
        idx = np.empty(nq, np.int32); best = np.empty(nq, np.uint16); second = np.empty(nq, np.uint16); match = np.empty(nq, np.int32)
        _check(lib().orbx_knn2(self.device, _p(query), nq, _p(train), len(train), th_low, self.fNNRatio_, _p(idx), _p(best), _p(second),
                               _p(match)))
        return idx, best, second, match

    def knn2_device(self, d_query, d_train, th_low=TH_LOW, stream=None):
        """torch CUDA uint8 tensors (nq, 32), (nt, 32) -> torch tensors (idx i32, best, second as int16 bit patterns, match i32)."""
        import torch
        nq = d_query.shape[0]
        dev = d_query.device
        idx = torch.empty(nq, dtype=torch.int32, device=dev); best = torch.empty(nq, dtype=torch.int16, device=dev)
        second = torch.empty(nq, dtype=torch.int16, device=dev); match = torch.empty(nq, dtype=torch.int32, device=dev)
        st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream)
        _check(lib().orbx_knn2_device(C.c_void_p(d_query.data_ptr()), nq, C.c_void_p(d_train.data_ptr()), d_train.shape[0], th_low,
                                      self.fNNRatio_, C.c_void_p(idx.data_ptr()), C.c_void_p(best.data_ptr()),
                                      C.c_void_p(second.data_ptr()), C.c_void_p(match.data_ptr()), st))
        return idx, best, second, match


def knn2_partial_device(d_query, d_train_shard, index_base, d_partial=None, stream=None):
    """One rank's share of the train-sharded scan: packed (best << 48 | second << 32 | global index) per query."""
    import torch
    nq = d_query.shape[0]
    if d_partial is None:
        d_partial = torch.empty(nq, dtype=torch.int64, device=d_query.device)
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(d_query.device).cuda_stream)
    _check(lib().orbx_knn2_partial_device(C.c_void_p(d_query.data_ptr()), nq, C.c_void_p(d_train_shard.data_ptr()),
                                          d_train_shard.shape[0], index_base, C.c_void_p(d_partial.data_ptr()), st))
    return d_partial


def knn2_merge_device(d_gathered, ranks, nq, th_low=TH_LOW, nnratio=0.6, stream=None):
    import torch
    dev = d_gathered.device
    idx = torch.empty(nq, dtype=torch.int32, device=dev); best = torch.empty(nq, dtype=torch.int16, device=dev)
    second = torch.empty(nq, dtype=torch.int16, device=dev); match = torch.empty(nq, dtype=torch.int32, device=dev)
    st = C.c_void_p(stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream)
    _check(lib().orbx_knn2_merge_device(C.c_void_p(d_gathered.data_ptr()), ranks, nq, th_low, nnratio, C.c_void_p(idx.data_ptr()),
                                        C.c_void_p(best.data_ptr()), C.c_void_p(second.data_ptr()), C.c_void_p(match.data_ptr()), st))
    return idx, best, second, match


def ConvertToGray(src, RGB=True, device=0):
    """static ConvertToGray of src/System.cc:122-137 for a (H, W, 3|4) uint8 image; 1-channel input is returned as is (:129-133)."""
    src = np.asarray(src)
    if src.ndim == 2:
        return src
    src = np.ascontiguousarray(src, np.uint8)
    h, w, ch = src.shape
    dst = np.empty((h, w), np.uint8)
    _check(lib().orbx_convert_to_gray(device, _p(src), w, h, src.strides[0], ch, int(RGB), _p(dst), dst.strides[0]))
    return dst


def UndistortKeyPoints(keypoints, camera, distCoeffs, device=0):
    """src/System.cc:153-174: keypoints with their positions run through cv::undistortPoints(.., K, distCoeffs, noArray(), K)."""
    kps = np.ascontiguousarray(keypoints).view(KP_DTYPE)
    d = np.ascontiguousarray(distCoeffs, np.float32).reshape(-1)
    out = np.empty_like(kps)
    cam = _Camera(*[float(c) for c in camera])
    _check(lib().orbx_undistort_keypoints(device, _p(kps), len(kps), C.byref(cam), _p(d), len(d), _p(out)))
    return out


def Remap(src, map1, map2, device=0):
    """cv::remap(src, dst, map1, map2, cv::INTER_LINEAR) as Examples/Stereo/stereo_euroc.cc:100-101 calls it (8-bit image, float32 maps)."""
    src = np.ascontiguousarray(src, np.uint8)
    map1 = np.ascontiguousarray(map1, np.float32); map2 = np.ascontiguousarray(map2, np.float32)
    dst = np.empty(map1.shape, np.uint8)
    _check(lib().orbx_remap(device, _p(src), src.shape[1], src.shape[0], src.strides[0], _p(map1), _p(map2), map1.strides[0], _p(dst),
                            dst.shape[1], dst.shape[0], dst.strides[0]))
    return dst


def ComputeStereoFromRGBD(keypoints, keypointsUn, depthImage, camera, device=0):
    """src/System.cc:197-219. depthImage: (H, W) float32. Returns (uright, depth)."""
    k = np.ascontiguousarray(keypoints, KP_DTYPE); ku = np.ascontiguousarray(keypointsUn, KP_DTYPE)
    dm = np.ascontiguousarray(depthImage, np.float32)
    ur = np.full(len(k), -1, np.float32); dp = np.full(len(k), -1, np.float32)
    cam = _Camera(*[float(v) for v in camera])
    _check(lib().orbx_stereo_from_rgbd(device, _p(k), _p(ku), len(k), _p(dm), dm.shape[1], dm.shape[0], dm.strides[0], C.byref(cam), _p(ur), _p(dp)))
    return ur, dp


def ComputeDistinctiveDescriptors(descriptor_sets, device=0):
    """The distance matrix + least-median selection of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:286-314) for a list
    of (N_i, 32) descriptor arrays; returns the index of the most distinctive descriptor of every set (-1 for an empty set)."""
    sets = [np.ascontiguousarray(d, np.uint8).reshape(-1, 32) for d in descriptor_sets]
    off = np.zeros(len(sets) + 1, np.int64)
    off[1:] = np.cumsum([len(d) for d in sets])
    allrows = np.concatenate(sets) if off[-1] else np.zeros((1, 32), np.uint8)
    best = np.empty(len(sets), np.int32)
    _check(lib().orbx_distinctive_descriptors(device, _p(allrows), _p(off), len(sets), _p(best)))
    return best


class ORBVocabulary:
    """ORB_SLAM2::ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (include/ORBVocabulary.h,
    Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h): the part Frame::ComputeBoW / KeyFrame::ComputeBoW and the key-frame database use —
    loadFromTextFile, transform(features, BowVector, FeatureVector, levelsup) and score. The tree lives on the GPU."""
    L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = range(6)   # DBoW2::ScoringType
    TF_IDF, TF, IDF, BINARY = range(4)                                        # DBoW2::WeightingType

    def __init__(self, device=0):
        self.device = device
        self._h = C.c_void_p()

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value and _lib is not None:
            _lib.orbx_vocabulary_destroy(h)
            self._h = C.c_void_p()

    def loadFromTextFile(self, filename):
        """TemplatedVocabulary.cpp:21-90; False where the reference returns false."""
        self.__del__()
        st = lib().orbx_vocabulary_load_text(os.fsencode(filename), self.device, C.byref(self._h))
        if st == ORBX_ERR_INVALID:
            return False
        _check(st)
        return True

    def create(self, k, L, parent, is_leaf, descriptors, weights, scoring=0, weighting=0):
        """The same tree from arrays: entry i is node id i + 1 in file order."""
        self.__del__()
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        descriptors = np.ascontiguousarray(descriptors, np.uint8); weights = np.ascontiguousarray(weights, np.float64)
        d = _VocabularyDesc(k, L, scoring, weighting, len(parent), parent.ctypes.data, is_leaf.ctypes.data, descriptors.ctypes.data, weights.ctypes.data)
        _check(lib().orbx_vocabulary_create(C.byref(d), self.device, C.byref(self._h)))
        return self

    def info(self):
        v = [C.c_int() for _ in range(4)] + [C.c_int64(), C.c_int64()]
        _check(lib().orbx_vocabulary_info(self._h, *[C.byref(x) for x in v]))
        return dict(zip(('k', 'L', 'scoring', 'weighting', 'nodes', 'words'), [x.value for x in v]))

    def empty(self):
        return self.info()['words'] == 0

    def transform(self, features, levelsup=4, per_feature=False):
        """transform(features, v, fv, levelsup), TemplatedVocabulary.h:1129-1197. features: N x 32 descriptor rows. Returns
        (word_ids, word_values) = the BowVector in ascending word id and (node_ids, start, indices) = the FeatureVector as CSR
        (what ORBmatcher.SearchByBoW takes); with per_feature also every feature's (word, node)."""
        features = np.ascontiguousarray(features, np.uint8).reshape(-1, 32)
        n = len(features)
        wi = np.empty(n + 1, np.int32); wv = np.empty(n + 1, np.float64)
        fn = np.empty(n + 1, np.uint32); fs = np.empty(n + 2, np.int32); fi = np.empty(n + 1, np.uint32)
        fw = np.empty(n + 1, np.int32); fnode = np.empty(n + 1, np.int32)
        nw, nf = C.c_int32(0), C.c_int32(0)
        _check(lib().orbx_bow_transform(self._h, _p(features), n, levelsup, _p(wi), _p(wv), C.byref(nw), _p(fn), _p(fs), _p(fi), C.byref(nf),
                                        _p(fw) if per_feature else None, _p(fnode) if per_feature else None))
        m = nf.value
        out = (wi[:nw.value].copy(), wv[:nw.value].copy()), (fn[:m].copy(), fs[:m + 1].copy(), fi[:fs[m]].copy())
        return out + ((fw[:n].copy(), fnode[:n].copy()),) if per_feature else out

    def transform_batch_device(self, d_desc, d_n, frames, cap, levelsup=4, stream=None):
        """Batch of frames resident on the device (torch tensors: d_desc [frames, cap, 32] u8, d_n [frames] i32), e.g. what
        ORBextractor.extract_batch_device leaves there. Returns torch tensors (word_ids, word_vals, fv_nodes, fv_start, fv_items, counts)."""
        import torch
        dev = d_desc.device
        word_ids = torch.empty((frames, cap), dtype=torch.int32, device=dev); word_vals = torch.empty((frames, cap), dtype=torch.float64, device=dev)
        fv_nodes = torch.empty((frames, cap), dtype=torch.int32, device=dev); fv_start = torch.empty((frames, cap + 1), dtype=torch.int32, device=dev)
        fv_items = torch.empty((frames, cap), dtype=torch.int32, device=dev); counts = torch.empty((frames, 2), dtype=torch.int32, device=dev)
        _check(lib().orbx_bow_transform_batch_device(self._h, d_desc.data_ptr(), d_n.data_ptr(), frames, cap, levelsup, word_ids.data_ptr(),
                                                     word_vals.data_ptr(), fv_nodes.data_ptr(), fv_start.data_ptr(), fv_items.data_ptr(),
                                                     counts.data_ptr(), None, None, stream))
        return word_ids, word_vals, fv_nodes, fv_start, fv_items, counts

    def score(self, a, b):
        """TemplatedVocabulary::score for the reference's L1 vocabulary (ScoringObject.cpp:24-58); a, b = (word_ids, word_values)."""
        return float(self.score_pairs([a, b], [(0, 1)])[0])

    def score_pairs(self, vectors, pairs):
        ids = np.concatenate([np.asarray(v[0], np.int32) for v in vectors]) if vectors else np.empty(0, np.int32)
        vals = np.concatenate([np.asarray(v[1], np.float64) for v in vectors]) if vectors else np.empty(0, np.float64)
        off = np.zeros(len(vectors) + 1, np.int32); off[1:] = np.cumsum([len(v[0]) for v in vectors])
        pa = np.ascontiguousarray([p[0] for p in pairs], np.int32); pb = np.ascontiguousarray([p[1] for p in pairs], np.int32)
        out = np.empty(len(pairs), np.float64)
        ids = np.ascontiguousarray(ids); vals = np.ascontiguousarray(vals)
        _check(lib().orbx_bow_score_l1(self._h, _p(ids), _p(vals), _p(off), _p(pa), _p(pb), len(pairs), _p(out)))
        return out


def measure_popc_peak(device=0):
    v = C.c_double()
    _check(lib().orbx_measure_popc_peak(device, C.byref(v)))
    return v.value
