import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle, saver_b200
from saver_b200 import ops
d = np.load('/root/repo/tests/golden/linnarsson.npz')
x = d['counts'].astype(np.float64); sf = d['size_factor']; y = x / sf
xest = np.ascontiguousarray(np.log((x + 1) / sf))
lm, sd = d['lambda_max'], d['sd_cv']
fixed = np.where(np.isnan(sd))[0][::8]
h = saver_b200.default_handle(0)
ops.set_design(xest, handle=h)
got = ops.predict_path(y[fixed], lm[fixed], d['lambda_min'][fixed], self_col=fixed, want_stats=True, handle=h)
ref = oracle.predict_path_batch(xest, y[fixed], lm[fixed], d['lambda_min'][fixed], self_col=fixed)
rel = (np.abs(got['mu'] - ref['mu']) / ref['mu']).max(1)
same_nlp = got['stats'][:, 2] == ref['nlp']
print('genes', fixed.size, 'same nlp', same_nlp.mean(), 'rel<1e-8', (rel < 1e-8).mean())
print('rel<1e-8 & same nlp', ((rel < 1e-8) & same_nlp).sum(), 'rel>=1e-8 & same nlp', ((rel >= 1e-8) & same_nlp).sum(),
      'rel<1e-8 & diff nlp', ((rel < 1e-8) & ~same_nlp).sum(), 'rel>=1e-8 & diff', ((rel >= 1e-8) & ~same_nlp).sum())
bad = np.where(rel >= 1e-8)[0][:12]
for b in bad:
    print(fixed[b], 'rel %.2e' % rel[b], 'gpu nlp/ngrad', got['stats'][b, 2], got['stats'][b, 3], 'ref', ref['nlp'][b], ref['ngrad'][b], 'lmu', got['stats'][b,0], ref['lmu'][b])
