import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def gpu():
    """Initialised B200 library (session scope).  Only requested by @pytest.mark.gpu tests."""
    import __graft_entry__ as ge
    ge.build()
    from mima_b200 import rrtmg
    rrtmg.set_device(0)
    rrtmg.rrtmg_lw_ini(allow_synthetic_lw=True)
    rrtmg.rrtmg_sw_ini()
    return rrtmg
