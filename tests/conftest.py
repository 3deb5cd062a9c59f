import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build libsmgpu.so / the oracle if they are missing (nvcc cross-compiles without a GPU)."""
    import smoothmesh_b200 as sm
    if not (os.path.exists(sm.LIB_PATH) and os.path.exists(os.path.join(ROOT, "oracle", "_build", "liboracle.so"))):
        import subprocess
        subprocess.check_call(["make", "-s", "-j8", "all"], cwd=ROOT)
    yield
