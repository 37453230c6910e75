import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_err(a, b):
    """print_error metric of the reference, DH/Utils.h:315-319."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(min(np.linalg.norm(a), np.linalg.norm(b)), 1e-4))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_scene_from_blob(ibuf, dbuf):
    from tests.blob_scene import scene_from_blob
    return scene_from_blob(ibuf, dbuf)
