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
    from oracle import oracle_binding
    return oracle_binding.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libbmf_ref.so); built in the authoring container only."""
    from oracle import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref/libbmf_ref.so not built (needs /root/reference: make -C oracle ref)")
    return ref_binding.RefLib()


@pytest.fixture(scope="session")
def gpu():
    """A bmf_ctx on cuda:0 through the C ABI.  Fails (never skips) if the library or the device is missing."""
    from binarymeshfitting_b200 import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
