import sys, numpy as np
sys.path.insert(0, '/root/repo')
import saver_b200
from saver_b200 import host, ops
d = np.load('/root/repo/tests/golden/linnarsson.npz')
x = d['counts'].astype(np.float64)
h = saver_b200.default_handle(0)
res = saver_b200.saver(x, do_fast=True, perm=1, fold_seed=2, backend=host.CudaBackend(handle=h), return_mu=True)
bad = np.where(~np.isfinite(res.estimate).all(1))[0]
print('bad genes', bad[:20], 'count', bad.size)
info = res.info
for g in bad[:6]:
    print(g, 'maxcor', info['maxcor'][g], 'lmax', info['lambda.max'][g], 'lmin', info['lambda.min'][g], 'sdcv', info['sd.cv'][g], 'status', res.status[g], 'mu finite', np.isfinite(res.mu[g]).all(), 'mu min/max', np.nanmin(res.mu[g]), np.nanmax(res.mu[g]))
if bad.size:
    g = int(bad[0])
    sf = res.sf; y = x / sf
    xest = np.ascontiguousarray(np.log((x + 1) / sf))
    ops.set_design(xest, handle=h)
    for gram in (1, 0):
        h.set_option('lasso_gram', gram)
        r = ops.predict_path(y[g:g+1], info['lambda.max'][g:g+1], info['lambda.min'][g:g+1], self_col=np.array([g]), want_stats=True, handle=h)
        print('gram', gram, 'path stats', r['stats'], 'finite', np.isfinite(r['mu']).all(), r['status'])
