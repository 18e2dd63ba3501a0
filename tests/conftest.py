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


@pytest.fixture(scope="session", params=["multi-kernel", "cluster-on-chip", "cluster-l2"])
def ctx(request):
    """Every parity test runs three times: through the multi-kernel PCG loop (the path large systems
    take; forced here with B200PCG_SMALL_N=0), through the single-launch thread-block-cluster kernel
    with the system in shared memory / DSMEM (what systems of up to ~32 K cells take by default), and
    through the cluster kernel with the system in L2 (what systems of up to 150 K cells take)."""
    from firefoam_dev_b200 import Context
    env = {"multi-kernel": {"B200PCG_SMALL_N": "0"},
           "cluster-on-chip": {"B200PCG_SMALL_N": "150000", "B200PCG_SMALL_FAST": "1"},
           "cluster-l2": {"B200PCG_SMALL_N": "150000", "B200PCG_SMALL_FAST": "0"}}[request.param]
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        c = Context()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    yield c
    c.close()
