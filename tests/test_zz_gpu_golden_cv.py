"""GPU: cross-validation against the reference's STORED results (not only against the oracle).

tests/golden/linnarsson_cv_seed_scan.npz holds, for every cross-validated gene of the reference's
stored linnarsson run, the seed j whose R folds (set.seed(j); sample(rep(seq(5), length = 200)))
make the CPU oracle reproduce the stored lambda.min and sd.cv (test_oracle_golden.py, T13).  Here
the CUDA solver gets the same R folds for the genes where that identification is sharpest and
must arrive at the values R stored.  (Last file of the suite on purpose: it is the one GPU test
whose thresholds were set from CPU-side evidence only.)"""
import os

import numpy as np
import pytest

from saver_b200 import ops, rrng

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


@pytest.mark.parametrize("gram", [0, 1], ids=["nspace", "gram"])
def test_cv_with_r_folds_reproduces_the_stored_lambda_min_and_sd_cv(golden, handle, gram):
    scan = np.load(os.path.join(ROOT, "tests", "golden", "linnarsson_cv_seed_scan.npz"))
    genes = scan["genes"]
    gl, gs = golden["lambda_min"][genes], golden["sd_cv"][genes]
    same = np.abs(scan["lambda_min"] - gl[:, None]) < 1e-9 * gl[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(same & (gs[:, None] > 0), np.abs(scan["sd_cv"] / gs[:, None] - 1.0), np.inf)
    best = rel.min(1)
    pick = np.argsort(best)[:24]
    assert best[pick].max() < 1e-3          # the oracle reproduces these to better than 0.1 %
    sel = genes[pick]
    seeds = scan["j"][np.argmin(rel, 1)][pick]
    n = golden["x"].shape[1]
    fold = np.stack([rrng.cv_folds(int(j), n, 5, "rounding") for j in seeds]).astype(np.int32)
    handle.set_option("lasso_gram", gram)
    try:
        ops.set_design(golden["xest"], handle=handle)
        got = ops.predict_cv(golden["y"][sel], fold, self_col=sel, want_stats=True, handle=handle)
    finally:
        handle.set_option("lasso_gram", 1)
    assert (got["status"] == 0).all()
    assert _rel(got["lambda_max"], golden["lambda_max"][sel]).max() < 1e-12
    ok_l = _rel(got["lambda_min"], golden["lambda_min"][sel]) < 1e-9
    ok_s = _rel(got["sd_cv"], golden["sd_cv"][sel]) < (5e-3 if gram == 0 else 2e-2)
    # n-space kernel = fishnet1's iterate sequence, i.e. the oracle's; the Gram kernel reaches the
    # same optimum by other iterates, so sd.cv carries its own convergence error on top
    assert ok_l.mean() >= (0.8 if gram == 0 else 0.6), ok_l.mean()
    assert (ok_l & ok_s).mean() >= (0.7 if gram == 0 else 0.4), (ok_l & ok_s).mean()
