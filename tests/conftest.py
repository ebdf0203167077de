import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle import hsoracle
    hsoracle.build()
    return hsoracle.Port()


@pytest.fixture(scope="session")
def refs():
    """The reference's own code (oracle/_ref). Present here (built from /root/reference) and on
    the GPU box (prebuilt .so travels); skipped only if neither exists."""
    from oracle import hsoracle
    if not all(hsoracle.ref_available(i) for i in hsoracle.IMPLS):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return {i: hsoracle.Ref(i) for i in hsoracle.IMPLS}
