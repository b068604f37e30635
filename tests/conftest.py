import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref built from /root/reference")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracles and the product library must exist; build them when they do not."""
    from oracle import oracle as O
    if not os.path.exists(O.PORT_LIB) or (os.path.isdir("/root/reference/source") and not O.have_ref()):
        O.build()
    lib = os.path.join(ROOT, "ray-tracing-engine_b200", "lib", "librt_b200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "ray-tracing-engine_b200"), "lib/librt_b200.so"],
                       check=True)


@pytest.fixture(scope="session")
def gold():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLD, name))
    return load


def scene_path(name):
    return os.path.join(GOLD, "scenes", f"{name}.rtscene")
