"""GPU vs oracle on CV genes of linnarsson (same folds): index agreement and mu error."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import oracle, saver_b200
from saver_b200 import ops, synth
d = np.load('/root/repo/tests/golden/linnarsson.npz')
x = d['counts'].astype(np.float64); sf = d['size_factor']; y = x / sf
xest = np.ascontiguousarray(np.log((x + 1) / sf))
lm, sd = d['lambda_max'], d['sd_cv']
cv = np.where(np.isfinite(sd) & (lm > 0))[0][::3]
h = saver_b200.default_handle(0)
ops.set_design(xest, handle=h)
fold = synth.fold_ids(cv.size, 200, 5, 3)
got = ops.predict_cv(y[cv], fold, self_col=cv, want_stats=True, want_cv=True, handle=h)
ref = oracle.predict_cv_batch(xest, y[cv], fold, self_col=cv)
same = got['idx'].astype(int) == ref['idx']
rel = np.abs(got['mu'] - ref['mu']) / ref['mu']
print('genes', cv.size, 'same idx', same.mean(), 'same lmu', (got['stats'][:, 0] == ref['lmu']).mean())
print('mu rel median(all)', np.median(rel), 'median over same-idx genes of max', np.median(rel[same].max(1)), 'max', rel[same].max())
print('idx diff', (got['idx'].astype(int) - ref['idx'])[~same])
