"""Generate tests/golden/linnarsson.npz from the reference's bundled data.

Run ONCE in the build container (it reads /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

Inputs : /root/reference/data/linnarsson.rda        (int 3529 x 200 counts)
         /root/reference/data/linnarsson_saver.rda  (golden saver object,
         produced by the reference's own ``saver(linnarsson, ncores = 12)``,
         /root/reference/R/saver.R:75)
Output : tests/golden/linnarsson.npz, holding
         counts       int16  (3529, 200)
         estimate_m   int32  (3529, 200)   golden estimate * 1000 (exact: the
         se_m         int32  (3529, 200)   reference rounds up to 3 decimals,
                                           /root/reference/R/calc_posterior.R:66-67)
         size_factor, maxcor, lambda_max, lambda_min, sd_cv, pred_time,
         var_time     float64 (per cell / per gene)
         cutoff, lambda_coefs, total_time_mins
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import rdx2  # noqa: E402

REF = "/root/reference/data"


def main():
    lin = rdx2.load_rda(os.path.join(REF, "linnarsson.rda"))["linnarsson"]
    counts, (genes, cells) = rdx2.as_matrix(lin)
    assert counts.shape == (3529, 200) and counts.min() >= 0 and counts.max() < 32768
    sav = rdx2.as_list(rdx2.load_rda(os.path.join(REF, "linnarsson_saver.rda"))["linnarsson_saver"])
    est, _ = rdx2.as_matrix(sav["estimate"])
    se, _ = rdx2.as_matrix(sav["se"])
    info = rdx2.as_list(sav["info"])
    est_m = np.rint(est * 1000.0)
    se_m = np.rint(se * 1000.0)
    assert np.abs(est_m / 1000.0 - est).max() < 1e-9 and np.abs(se_m / 1000.0 - se).max() < 1e-9
    out = dict(
        counts=counts.astype(np.int16),
        estimate_m=est_m.astype(np.int32),
        se_m=se_m.astype(np.int32),
        size_factor=info["size.factor"]["value"],
        maxcor=info["maxcor"]["value"],
        lambda_max=info["lambda.max"]["value"],
        lambda_min=info["lambda.min"]["value"],
        sd_cv=info["sd.cv"]["value"],
        pred_time=info["pred.time"]["value"],
        var_time=info["var.time"]["value"],
        cutoff=info["cutoff"]["value"],
        lambda_coefs=info["lambda.coefs"]["value"],
        total_time_mins=info["total.time"]["value"],
        gene_names=np.array(genes),
        cell_names=np.array(cells),
    )
    path = os.path.join(HERE, "linnarsson.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
