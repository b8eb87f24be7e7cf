import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    # the oracle's C restatement is test infrastructure; build it once (seconds)
    from oracle import capi
    capi.lib()
    yield


@pytest.fixture(scope="session")
def prfdd():
    import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr
    pr.lib()
    return pr


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
