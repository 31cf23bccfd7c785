import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build the native pieces when they are missing (the driver normally runs build() first)."""
    lib = os.path.join(ROOT, "grayshift_b200", "lib")
    need = [os.path.join(lib, "libgrayshift_cuda.so"), os.path.join(lib, "libgrayshift_host.so"),
            os.path.join(ROOT, "oracle", "_build", "liboracle.so"), os.path.join(ROOT, "oracle", "_build", "liboracle_counters.so"),
            os.path.join(ROOT, "assets", "earthmap.ppm")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__

        __graft_entry__.build()
    return True


def tonemap_np(x):
    """write_color's transfer (src/color.rs:8-26) without the quantisation."""
    import numpy as np

    return np.clip(np.sqrt(np.maximum(x, 0.0)), 0.0, 0.999)
