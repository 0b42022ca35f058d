import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(1, HERE)      # tests/guided_cases.py is shared with tests/golden/make_golden.py


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def oracle_port():
    """Stand-alone CPU restatement (oracle/orb_oracle.cc). Test infrastructure only."""
    from oracle import bindings
    bindings.build()
    return bindings.Oracle('port')


@pytest.fixture(scope='session')
def oracle_ref():
    """The reference's own TUs (oracle/_ref). Present where /root/reference exists or where the prebuilt .so travelled."""
    from oracle import bindings
    try:
        bindings.build()
        return bindings.Oracle('ref')
    except FileNotFoundError:
        pytest.skip('oracle/_ref/liborb_ref.so not available (no /root/reference and no prebuilt copy)')


@pytest.fixture(scope='session')
def oracle_final():
    """The strongest checker present for whole-pipeline outputs: the reference's own TUs (oracle/_ref) where the library exists
    (it travels to the GPU box with the snapshot), else the restatement. port == ref is held by tests/test_oracle_vs_ref.py."""
    from oracle import bindings
    bindings.build()
    try:
        return bindings.Oracle('ref')
    except (FileNotFoundError, OSError):
        return bindings.Oracle('port')


@pytest.fixture(scope='session')
def orbx():
    """The product: liborbx_b200.so through its Python mirror. Fails loudly if the library is missing."""
    from orb_slam2_refactored_b200 import api
    api.lib()
    if api.device_count() < 1:
        pytest.fail('no sm_100 device visible: the gpu tests must run on a B200')
    return api
