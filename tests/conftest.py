import os
import sys

import pytest

# parameter values for the reference's cslam/config.h when it is compiled against the cv::FileStorage stand-in (oracle/ref_stub): read
# during static initialisation of oracle/_ref/liboptimizer_shim.so (tests/test_shim_optimizer.py)
os.environ.setdefault("CCM_REF_STUB_CONF", "Opt.EssGraphMinFeats=100,Timing.LockSleep=1000")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle
