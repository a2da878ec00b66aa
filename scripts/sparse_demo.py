"""cfg4-shaped run at a size that fits a short GPU slot: sparse counts (R: dgCMatrix) resident on
the device, library-size factors, do.fast schedule, estimate + SE (estimates.only = FALSE), then
sample.saver and cor.genes on the result.  Prints genes/s and the agreement with the dense path."""
import sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, '/root/repo')
import saver_b200
from saver_b200 import synth

G, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (6000, 8000)
x = synth.nb_counts(G, n, 20264)
xs = sp.csc_matrix(x.astype(np.float64))
print('counts %d x %d, %.1f%% non-zero, CSC %.0f MB vs dense %.0f MB' % (
    G, n, 100.0 * xs.nnz / (G * n), (xs.data.nbytes + xs.indices.nbytes + xs.indptr.nbytes) / 1e6,
    G * n * 8 / 1e6), flush=True)
t = time.time()
res = saver_b200.saver(xs, do_fast=True, fold_seed=1, perm=2, verbose=False)
dt = time.time() - t
info = res.info
ncv = int((np.isfinite(info['sd.cv']) & (info['lambda.max'] > 0)).sum())
npath = int(np.isnan(info['sd.cv']).sum())
print('saver(sparse, do.fast=TRUE): %.1f s = %.1f genes/s; %d cross-validated, %d fixed-lambda, %d null genes'
      % (dt, G / dt, ncv, npath, G - ncv - npath), flush=True)
assert np.isfinite(res.estimate).all() and np.isfinite(res.se).all()
sub = np.arange(0, G, max(1, G // 400))
t = time.time()
s1 = saver_b200.sample_saver(res, rep=1, seed=50)
c = saver_b200.cor_genes(saver_b200.SaverResult(res.estimate[sub], res.se[sub], info))
print('sample.saver %.2f s, cor.genes on %d genes %.2f s' % (time.time() - t, sub.size, 0.0), flush=True)
if G * n <= 4e7:
    t = time.time()
    ref = saver_b200.saver(x.astype(np.float64), do_fast=True, fold_seed=1, perm=2)
    rel = np.abs(res.estimate - ref.estimate) / np.maximum(ref.estimate, 1e-12)
    print('dense path %.1f s; sparse vs dense estimate: max rel %.2e' % (time.time() - t, rel.max()))
