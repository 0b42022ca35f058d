"""ComputeOrbDescriptor's rotation terms, exhaustively: a = (float)cos((double)(angle * factorPI)), b = (float)sin(...)
(src/ORBextractor.cc:105-107) for EVERY float angle in [0, 360) — 1 135 869 953 bit patterns — on the device (CUDA's FP64 cos / sin,
<= 2 ulp) against the host's libm through the oracle (glibc, the reference's own expression text in oracle/_ref where present).
A differing float needs the double to sit within 2 ulp of a float rounding boundary; this test turns "not observed" into a count."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LAST = int(np.float32(360.0).view(np.uint32))       # 0x43B40000: angles are in [0, 360), fastAtan2 never returns 360


def test_every_float_angle(orbx, oracle_final):
    import ctypes as C
    lib = orbx.lib()
    chunk = 1 << 25
    threads = max(1, min(32, os.cpu_count() or 1))
    bad = []
    gc = np.empty(chunk, np.float32); gs = np.empty(chunk, np.float32)
    total = 0
    for first in range(0, LAST + 1, chunk):
        n = min(chunk, LAST + 1 - first)
        orbx._check(lib.orbx_debug_cos_sin(0, first, n, gc.ctypes.data_as(C.c_void_p), gs.ctypes.data_as(C.c_void_p)))
        hc, hs = oracle_final.cos_sin_range(first, n, threads)
        d = np.nonzero((gc[:n].view(np.uint32) != hc.view(np.uint32)) | (gs[:n].view(np.uint32) != hs.view(np.uint32)))[0]
        bad += [(first + int(i), float(gc[i]), float(hc[i]), float(gs[i]), float(hs[i])) for i in d[:8]]
        total += n
    assert total == LAST + 1
    assert not bad, f'{len(bad)}+ float angles where the device and the host differ (bits, cos dev/host, sin dev/host): {bad[:8]}'
