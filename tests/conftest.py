import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_sessionstart(session):
    """The C-ABI library is a build artefact (git-ignored): build it once if a fresh checkout has none (nvcc cross-compiles
    sm_100a without a GPU; a few minutes).  The product itself never builds or falls back at import time."""
    lib = os.path.join(ROOT, "jax_cosmo_b200", "libjc_b200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["bash", os.path.join(ROOT, "jax_cosmo_b200", "csrc", "build.sh")], check=True, cwd=ROOT)


def golden_cl_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "cl_*.npz")))


def load_golden(path):
    g = np.load(path)
    scn = json.loads(str(g["spec"]))
    return scn, g


def relerr(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


@pytest.fixture(scope="session")
def jc():
    import jax_cosmo_b200
    return jax_cosmo_b200
