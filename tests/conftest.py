import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the in-tree libraries if they are missing (no-op on the GPU box: they travel)."""
    from firefoam_dev_b200 import _lib
    from oracle import oracle as orc
    if not (os.path.exists(_lib.PCG_SO) and os.path.exists(_lib.MESH_SO)):
        _lib.build()
    if not os.path.exists(orc.SO):
        orc.build()


def has_gpu():
    try:
        from firefoam_dev_b200 import _lib
        return _lib.load_pcg().b200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session", params=["multi-kernel", "cluster-kernel"])
def ctx(request):
    """Every parity test runs twice: through the multi-kernel PCG loop (the path large systems
    take; forced here with B200PCG_SMALL_N=0) and through the single-launch thread-block-cluster
    kernel that systems of up to 65 536 cells take by default."""
    from firefoam_dev_b200 import Context
    old = os.environ.get("B200PCG_SMALL_N")
    os.environ["B200PCG_SMALL_N"] = "0" if request.param == "multi-kernel" else "65536"
    try:
        c = Context()
    finally:
        if old is None:
            os.environ.pop("B200PCG_SMALL_N", None)
        else:
            os.environ["B200PCG_SMALL_N"] = old
    yield c
    c.close()
