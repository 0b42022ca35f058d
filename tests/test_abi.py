"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/orbx.h declares, and — with
no GPU — fails loudly instead of computing anything on the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def api():
    from orb_slam2_refactored_b200 import api, build
    build.build()
    api.lib()
    return api


def test_header_symbols_are_exported_and_bound(api):
    header = open(os.path.join(ROOT, 'include', 'orbx.h')).read()
    declared = sorted(set(re.findall(r'\b(orbx_[a-z0-9_]+)\s*\(', header)))
    assert len(declared) >= 25
    lib = C.CDLL(api.library_path())
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/orbx.h but not exported by liborbx_b200.so'
    assert sorted(api.exported_symbols()) == declared, 'the Python mirror must bind exactly the declared ABI'


def test_every_declaration_cites_the_reference():
    header = open(os.path.join(ROOT, 'include', 'orbx.h')).read()
    assert len(re.findall(r'(src|include)/[A-Za-z]+\.(cc|h):\d+', header)) >= 15


def test_keypoint_layout_mirrors_cv_keypoint(api):
    assert api.KP_DTYPE.itemsize == 28
    assert api.KP_DTYPE.names == ('x', 'y', 'size', 'angle', 'response', 'octave', 'class_id')


def test_no_cpu_fallback(api):
    if api.device_count() > 0:
        pytest.skip('a B200 is visible; this test is for GPU-less machines')
    with pytest.raises(api.OrbxError) as e:
        api.ORBextractor(nfeatures=1000)
    assert e.value.status == api.ORBX_ERR_CUDA
    a = np.zeros((4, 32), np.uint8)
    with pytest.raises(api.OrbxError) as e:
        api.ORBmatcher.DescriptorDistance(a, a)
    assert e.value.status == api.ORBX_ERR_CUDA
    with pytest.raises(api.OrbxError) as e:
        api.ORBmatcher().knn2(a, a)
    assert e.value.status == api.ORBX_ERR_CUDA


def test_product_does_not_use_the_oracle():
    """The product path must not import, include, link or execute anything under oracle/ (comments may cite it)."""
    import re
    pkg = os.path.join(ROOT, 'orb_slam2_refactored_b200')
    bad = re.compile(r'^\s*(from\s+oracle|import\s+oracle)|#\s*include\s*[<"][^>"]*oracle/|liborb_oracle|liborb_ref|oracle\.bindings', re.M)
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cc')):
                text = open(os.path.join(d, f), errors='ignore').read()
                assert not bad.search(text), f'{f} uses the oracle'
    for f in ('include/orbx.h', 'include/orbx/ORBextractor.h', 'include/orbx/ORBmatcher.h'):
        assert not bad.search(open(os.path.join(ROOT, f)).read()), f
