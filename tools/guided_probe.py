"""Times the guided matchers through the C ABI (host buffers in and out, one search per call) against the reference text on one host
core. Usage: python tools/guided_probe.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))

import guided_cases as gc  # noqa: E402
from oracle import bindings  # noqa: E402
from orb_slam2_refactored_b200 import api, synth  # noqa: E402


def bench(fn, reps):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps * 1e6


def main():
    bindings.build(native=True)
    try:
        ref = bindings.Oracle('ref', native=True)
    except FileNotFoundError:
        ref = bindings.Oracle('port', native=True)
    sizes = {'C1': (1000, 640, 480, 1000), 'C2': (2000, 1241, 376, 2000), 'C4': (8000, 3840, 2160, 6000)}
    if '--quick' in sys.argv:      # profiling runs: the small case only
        sizes = {'C1': sizes['C1']}
    for tag, (n, w, h, npts) in sizes.items():
        fr = synth.frame(1, n=n, w=w, h=h)
        f = gc.make_frame(api, fr)
        mp0 = np.full(n, -1, np.int32)
        pts, desc = synth.local_map_points(1, fr, npts=npts)
        cp, lp, lpts, ldesc = synth.last_frame_points(1, fr, synth.KITTI_CAMERA, npts=npts)
        m = api.ORBmatcher(0.8, True)

        def g_local():
            f.mappoints[:] = -1
            return m.SearchByProjection(f, pts, desc, 3.0)

        def g_last():
            f.mappoints[:] = -1
            return m.SearchByProjectionLastFrame(f, synth.KITTI_CAMERA, cp, lp, lpts, ldesc, 7.0, False)

        t_create = bench(lambda: gc.make_frame(api, fr), 10)
        t_assign = bench(lambda: f.assign(fr['kps_un'], fr['desc'], fr['scale_factors'], fr['bounds'], fr['uright']), 50)
        r = dict(frame_create=t_create, frame_assign=t_assign, gpu_local=bench(g_local, 50), kernel_local=f.last_stats()[1] * 1e3,
                 rounds_local=f.last_rounds(), gpu_last=bench(g_last, 50), kernel_last=f.last_stats()[1] * 1e3, rounds_last=f.last_rounds(),
                 cpu_local=bench(lambda: ref.search_local_map(fr, mp0, pts, desc, 3.0, 0.8), 10),
                 cpu_last=bench(lambda: ref.search_last_frame(fr, synth.KITTI_CAMERA, cp, lp, mp0, lpts, ldesc, 7.0, False, 0.9, True), 10))
        g_local(); print(tag, 'local phases us', np.round(f.last_phase_us(), 1)); g_last(); print(tag, 'last phases us', np.round(f.last_phase_us(), 1))
        print(tag, {k: round(v, 1) for k, v in r.items()}, 'us (cpu = %s incl. building its Frame/grid and the ctypes marshalling)' % ref.kind)


if __name__ == '__main__':
    main()
