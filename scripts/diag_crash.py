import sys, numpy as np
sys.path.insert(0, '/root/repo')
import saver_b200
from saver_b200 import ops
d = np.load('/root/repo/tests/golden/linnarsson.npz')
x = d['counts'].astype(np.float64); sf = d['size_factor']; y = x / sf
xest = np.ascontiguousarray(np.log((x + 1) / sf))
h = saver_b200.default_handle(0)
ops.set_design(xest, handle=h)
sd = d['sd_cv']
fixed = np.where(np.isnan(sd))[0][::8]
lo, hi = int(sys.argv[1]), int(sys.argv[2])
g = fixed[lo:hi]
got = ops.predict_path(y[g], d['lambda_max'][g], d['lambda_min'][g], self_col=g, want_stats=True, handle=h)
print('ok', lo, hi, got['stats'][:, 2].sum(), np.isfinite(got['mu']).all())
