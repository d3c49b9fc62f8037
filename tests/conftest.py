import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _engine_options_from_env():
    """B200_TEST_OPTS="name=value,..." runs the suite with engine options set (e.g. tc_pair=1: the CTA-pair
    tensor-core kernel behind every tile test).  Unset: nothing happens."""
    import os
    opts = os.environ.get("B200_TEST_OPTS", "")
    if opts and _has_gpu():
        from matrix_multiplication_b200 import capi
        for kv in opts.split(","):
            k, v = kv.split("=")
            capi.set_option(k.strip(), int(v))
    yield


def load_golden(name):
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def random_csr(rng, M, K, density, heavy_rows=(), empty_rows=(), dtype=np.float32, signed=True):
    """Seeded CSR test matrix: Bernoulli mask, optional full ('heavy') rows and forced-empty rows."""
    mask = rng.random((M, K)) < density
    for r in heavy_rows:
        mask[r, :] = True
    for r in empty_rows:
        mask[r, :] = False
    vals = rng.standard_normal((M, K)) if signed else rng.random((M, K))
    dense = (vals * mask).astype(dtype)
    dense[mask & (dense == 0)] = 1.0
    rp = np.zeros(M + 1, dtype=np.int32)
    rp[1:] = np.cumsum(mask.sum(axis=1))
    r, c = np.nonzero(mask)
    return rp, c.astype(np.int32), dense[r, c], dense
