"""Which genes make the fixed-lambda stage of do.fast slow?  Runs saver(do.fast=TRUE), then replays
the fixed-lambda genes through predict_path with per-gene statistics (path length, glmnet jerr,
sweeps); saves the inputs of the heaviest genes for a CPU replay through the oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saver_b200
from saver_b200 import host, ops, synth

G, n = int(sys.argv[1]), int(sys.argv[2])
x = synth.nb_counts(G, n, 20264).astype(np.float64)
t = time.time()
res = saver_b200.saver(x, do_fast=True, fold_seed=1, perm=2)
print("saver do.fast %.1f s" % (time.time() - t), flush=True)
info = res.info
fixed = np.where(np.isnan(info["sd.cv"]) & (info["lambda.max"] > 0))[0]
xc, sf = res.x, res.sf
good = np.where(xc.mean(1) >= 0.1)[0]
x_est = np.log((xc[good] + 1.0) / sf)
self_col = np.full(G, -1, dtype=np.int32); self_col[good] = np.arange(good.size, dtype=np.int32)
h = saver_b200.default_handle(0)
ops.set_design(x_est, handle=h)
y = xc[fixed] / sf
t = time.time()
r = ops.predict_path(y, info["lambda.max"][fixed], info["lambda.min"][fixed], self_col=self_col[fixed],
                     want_stats=True, handle=h)
print("predict_path on %d genes %.1f s" % (fixed.size, time.time() - t), flush=True)
st = np.asarray(r["stats"])
order = np.argsort(-st[:, 2])
print("sweeps: total %d, top 10 genes hold %.0f%%; jerr != 0 on %d genes" % (
    st[:, 2].sum(), 100.0 * st[order[:10], 2].sum() / st[:, 2].sum(), (st[:, 1] != 0).sum()))
ratio = info["lambda.min"][fixed] / info["lambda.max"][fixed]
for i in order[:12]:
    print("gene %5d  lmu %3d jerr %6d sweeps %7d rounds %3d  lambda.max %.4g lambda.min/max %.4g maxcor %.3f" % (
        fixed[i], st[i, 0], st[i, 1], st[i, 2], st[i, 3], info["lambda.max"][fixed[i]], ratio[i], info["maxcor"][fixed[i]]))
print("lambda.min/max quantiles", np.quantile(ratio, [0, .01, .1, .5, .9, 1]))
print("sweeps quantiles", np.quantile(st[:, 2], [0, .5, .9, .99, 1]))
os.makedirs("gpurun_out", exist_ok=True)
top = fixed[order[:6]]
np.savez("gpurun_out/fastpath_heavy.npz", genes=top, lmax=info["lambda.max"][top], lmin=info["lambda.min"][top],
         G=G, n=n, stats=st[order[:6]])
