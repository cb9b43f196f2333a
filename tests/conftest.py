import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests", ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def gpu_ctx():
    import gaussianvi_b200 as gv
    ctx = gv.Context(0)  # raises if the CUDA library or a device is missing: no silent fallback
    yield ctx
    ctx.close()
