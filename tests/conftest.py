import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libref.so)."""
    from oracle import cpu
    if not cpu.have("reference"):
        if os.path.exists("/root/reference/infMagSim_c.c"):
            cpu.build()
        else:
            pytest.skip("oracle/_ref/libref.so not built and /root/reference absent")
    return cpu.load("reference")


@pytest.fixture(scope="session")
def port():
    """Our C restatement (oracle/libespic_oracle.so)."""
    from oracle import cpu
    if not cpu.have("port"):
        cpu.build()
    return cpu.load("port")
