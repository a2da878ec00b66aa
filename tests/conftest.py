import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Reference golden object (tests/golden/make_golden.py)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "linnarsson.npz"))
    g = {k: d[k] for k in d.files}
    x = g["counts"].astype(np.float64)
    g["x"] = x
    g["sf"] = g["size_factor"]
    g["y"] = x / g["sf"]
    g["xest"] = np.ascontiguousarray(np.log((x + 1.0) / g["sf"]))  # (p, n): R/saver.R:156-157
    lm, sd = g["lambda_max"], g["sd_cv"]
    g["null_genes"] = np.where(lm == 0)[0]
    g["fixed_genes"] = np.where(np.isnan(sd))[0]
    g["cv_genes"] = np.where(np.isfinite(sd) & (lm > 0))[0]
    return g


@pytest.fixture(scope="session")
def handle():
    import saver_b200
    return saver_b200.default_handle(0)
