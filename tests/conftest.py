import os
import sys

# several logical shards in one process need their streams on distinct hardware queues
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def split_prefixed(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def assert_close(a, b, rtol, atol, name=""):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return
    bad_nan = np.isnan(a) != np.isnan(b)
    assert not bad_nan.any(), f"{name}: NaN mismatch at {np.argwhere(bad_nan)[:5].tolist()}"
    m = ~np.isnan(b)
    err = np.abs(a[m] - b[m])
    tol = atol + rtol * np.abs(b[m])
    worst = np.argmax(err - tol) if err.size else 0
    assert (err <= tol).all(), (f"{name}: max |diff| {err.max():.3e} (tol {tol.flat[worst]:.3e}) "
                                f"got {a[m].flat[worst]!r} want {b[m].flat[worst]!r}")


def assert_state_close(got, want, rtol=1e-5, atol=1e-6, atol_z=1e-4, name=""):
    for k in want:
        at = atol_z if k.endswith("_z") else atol
        if k == "bias":
            assert_close(got[k], want[k], rtol, atol_z, f"{name}:{k}")
        else:
            assert_close(got[k], want[k], rtol, at, f"{name}:{k}")
