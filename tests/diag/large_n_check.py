"""Sanity of the large-n code paths (cfg3 shapes): n = 10,000 cells, a few genes, CV vs oracle."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import oracle, saver_b200
from saver_b200 import ops, synth, host
G, n = int(sys.argv[1]), int(sys.argv[2])
x = synth.nb_counts(G, n, 20263).astype(np.float64)
sf = x.sum(0) / x.sum(0).mean()
good = np.where(x.mean(1) >= 0.1)[0]
xest = np.log((x[good] + 1) / sf)
print('design', xest.shape, flush=True)
h = saver_b200.default_handle(0)
ops.set_design(xest, handle=h)
sel = good[:: max(1, good.size // 12)][:12]
self_col = np.array([int(np.where(good == g)[0][0]) for g in sel], dtype=np.int32)
fold = host.draw_folds(np.arange(1, sel.size + 1), n, 5, 5)
t = time.time()
r = ops.predict_cv(x[sel] / sf, fold, self_col=self_col, want_stats=True, handle=h)
print('gpu s', time.time() - t, 'status', r['status'], 'lmu', r['stats'][:, 0], flush=True)
t = time.time()
ref = oracle.predict_cv_batch(xest, x[sel[:4]] / sf, fold[:4], self_col=self_col[:4])
print('oracle s (4 genes)', time.time() - t, 'idx gpu', r['idx'][:4], 'ref', ref['idx'])
rel = np.abs(r['mu'][:4] - ref['mu']) / ref['mu']
print('mu rel median', np.median(rel, 1), 'max', rel.max(1))
pr = ops.calc_post(x[sel], r['mu'], sf, 1.0, handle=h)
print('post finite', np.isfinite(pr['est']).all())
