import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libc8p_ref.so (the compiled unmodified reference)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = {}
    for n in ("frames_siso", "frames_bench", "ref_vectors", "frames_mimo", "frames_mu"):
        g[n] = np.load(os.path.join(HERE, "golden", n + ".npz"))
    return g
