"""pytest configuration: registers the `gpu` marker and puts the repo root on
sys.path so `oracle`, `millipyde_b200`, `bench` import from a plain checkout."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout)")
    # The native pieces are build artefacts (git-ignored): build them if a fresh checkout has none.
    # nvcc cross-compiles sm_100a without a GPU; on the GPU box the snapshot already carries them.
    from millipyde_b200 import build as _build
    if not (os.path.exists(_build.LIB) and os.path.exists(_build.EXT)):
        _build.build_all(force=False, verbose=False)


@pytest.fixture(scope="session")
def charlie_small():
    """The one real fixture carried over from the reference's tests/images/
    (500 x 667 RGBA8, alpha == 255); charlie.png itself is a missing blob."""
    from PIL import Image
    import numpy as np
    return np.asarray(Image.open(os.path.join(GOLDEN, "charlie_small.png")))
