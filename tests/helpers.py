"""Test-only backend: calc.estimate through the CPU oracle (same interface as host.CudaBackend).

Lets the host scheduler (saver.fit stages, regressions, sharding, gathers) be tested on a box
without a GPU, and gives the GPU tests an independent end-to-end reference."""
import numpy as np

import oracle
from saver_b200 import host


class OracleBackend:
    block = 64

    def __init__(self, threads=None):
        self.threads = threads
        self.xest = None
        self.calls = []

    def set_design(self, xest):
        self.xest = np.ascontiguousarray(np.asarray(xest), dtype=np.float64)

    def calc_estimate(self, x, sf, scale_sf, mode, cutoff=0.0, coefs=None, self_col=None,
                      is_pred=None, pred_mask=None, foldid=None, nfolds=5, calc_maxcor=False,
                      estimates_only=False, want_mu=False, dfmax=300, nlambda=100, thresh=1e-7):
        x = np.asarray(x, dtype=np.float64)
        B, n = x.shape
        self.calls.append((mode, B, bool(calc_maxcor), float(cutoff)))
        y = x / sf
        rows = np.arange(n) if pred_mask is None else np.where(np.asarray(pred_mask) != 0)[0]
        sc = np.full(B, -1) if self_col is None else np.asarray(self_col)
        ip = np.ones(B, dtype=bool) if is_pred is None else np.asarray(is_pred).astype(bool)
        maxcor = oracle.maxcor(self.xest, y, self_col=sc, threads=self.threads) if calc_maxcor \
            else np.full(B, 2.0)
        mu = np.repeat(y[:, rows].mean(1)[:, None], n, 1)
        lmax, lmin, sdcv = np.zeros(B), np.zeros(B), np.zeros(B)
        status = np.zeros(B, dtype=np.int32)
        for g in range(B):
            if mode == 0 or not (maxcor[g] > cutoff and ip[g]):
                continue
            if mode == 1:
                r = oracle.predict_cv(self.xest, y[g], foldid[g][rows], rows=rows, excl=int(sc[g]),
                                      nfolds=nfolds, dfmax=dfmax, thresh=thresh)
                lmax[g], lmin[g], sdcv[g] = r["lambda_max"], r["lambda_min"], r["sd_cv"]
            else:
                lam = host.est_lambda(y[g], maxcor[g], coefs)
                lmax[g], lmin[g] = lam
                r = oracle.predict_path(self.xest, y[g], lam[0], lam[1], rows=rows, excl=int(sc[g]),
                                        dfmax=dfmax, thresh=thresh)
                sdcv[g] = np.nan if r["rc"] == 0 else 0.0
            mu[g] = r["mu"]
            status[g] = r["rc"]
        post = oracle.calc_post_batch(x, mu, sf, scale_sf, threads=self.threads)
        return dict(est=post["est"], se=None if estimates_only else post["se"], maxcor=maxcor,
                    lambda_max=lmax, lambda_min=lmin, sd_cv=sdcv, ct=np.zeros(B), vt=np.zeros(B),
                    status=status, mu=mu if want_mu else None)
