import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle_lib import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference: make -C oracle ref)")
    return Ref()


@pytest.fixture(scope="session")
def ref_schar():
    from oracle_lib import Ref, ref_available
    if not ref_available(schar=True):
        pytest.skip("oracle/_ref (signed-char variant) not built")
    return Ref(schar=True)
