"""Where does the Gram kernel's distance to the oracle come from?  cfg2-shaped genes, the oracle at
glmnet's thresh (1e-7) and at 1e-10 (the optimum, for practical purposes), and solver variants."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import saver_b200  # noqa: E402
from saver_b200 import host, ops, synth  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 24
x0 = synth.nb_counts(5000, 2000, 20262)
xc, _ = host.clean_data(x0)
sf, scale_sf = host.calc_size_factor(xc, None, xc.shape[1])
good = np.where(xc.mean(1) >= 0.1)[0]
x_est = np.log((xc[good] + 1.0) / sf)
self_col = np.full(xc.shape[0], -1, dtype=np.int32)
self_col[good] = np.arange(good.size, dtype=np.int32)
pred = np.where(xc.sum(1) > 0)[0]
sel = pred[np.linspace(0, pred.size - 1, G).astype(int)]
fold = host.draw_folds(np.arange(1, G + 1), 2000, 5, fold_seed=5)
y = xc[sel] / sf
t = time.time()
ref = oracle.predict_cv_batch(x_est, y, fold, self_col=self_col[sel])
t1 = time.time() - t
t = time.time()
opt = oracle.predict_cv_batch(x_est, y, fold, self_col=self_col[sel], thresh=1e-10)
print("oracle thr 1e-7: %.1f s, thr 1e-10: %.1f s" % (t1, time.time() - t), flush=True)


def cmp(a, b, name):
    same = a["idx"] == b["idx"]
    rel = np.abs(a["mu"] - b["mu"]) / np.maximum(b["mu"], 1e-300)
    pg = np.median(rel, 1)
    print("  %-34s idx equal %2d/%d | rel mu: median(all) %.2e  median(same idx) %.2e  p90 per-gene median %.2e" % (
        name, same.sum(), same.size, np.median(rel), np.median(rel[same]) if same.any() else np.nan,
        np.quantile(pg, 0.9)), flush=True)


cmp(ref, opt, "oracle(1e-7) vs oracle(1e-10)")
h = saver_b200.default_handle(0)
ops.set_design(x_est, handle=h)
for name, opts in [("gram default", {}), ("gram inner_scale 0.01", {"inner_scale": 0.01}),
                   ("gram rebuild every IRLS (dbg 64)", {"lasso_dbg": 64}),
                   ("gram rebuild + inner 0.01", {"lasso_dbg": 64, "inner_scale": 0.01}),
                   ("gram inner_scale 1e-4", {"inner_scale": 1e-4}),
                   ("n-space kernel", {"lasso_gram": 0})]:
    h.set_option("lasso_gram", 1)
    h.set_option("lasso_dbg", 0)
    h.set_option("inner_scale", 1.0)
    for k, v in opts.items():
        h.set_option(k, v)
    ops.predict_cv(y, fold, self_col=self_col[sel], handle=h)
    t = time.time()
    got = ops.predict_cv(y, fold, self_col=self_col[sel], want_stats=True, handle=h)
    dt = time.time() - t
    print("%s: %.2f s, sweeps/gene %.0f" % (name, dt, got["stats"][:, 2].mean()), flush=True)
    got["idx"] = got["idx"].astype(int)
    cmp(got, ref, "vs oracle(1e-7)")
    cmp(got, opt, "vs oracle(1e-10)")
