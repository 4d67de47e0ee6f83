import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")
    # torch.func (used by the torch stand-in evaluator of tests/test_ipsolver_cpu.py) trips a deprecation notice inside torch
    config.addinivalue_line("filterwarnings", "ignore:.*torch.jit.script.*:DeprecationWarning")


@pytest.fixture(scope="session")
def model():
    from hippopt_b200.robot_model import synthetic_ergocub

    return synthetic_ergocub()


@pytest.fixture(scope="session")
def built_library():
    """The C-ABI shared library, built in-tree (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as entry

    entry.build()
    from hippopt_b200 import _capi

    return _capi.lib()
