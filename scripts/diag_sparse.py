import numpy as np, scipy.sparse as sp, sys
sys.path.insert(0, '.')
import saver_b200
from saver_b200 import synth
x = synth.nb_counts(150, 90, seed=2).astype(np.float64); x[:, 11] = 0
a = saver_b200.saver(x, fold_seed=5, do_fast=False)
b = saver_b200.saver(sp.csr_matrix(x), fold_seed=5, do_fast=False)
re = np.abs(a.estimate-b.estimate)/np.maximum(np.abs(a.estimate),1e-300)
rs = np.abs(a.se-b.se)/np.maximum(np.abs(a.se),1e-300)
print('est max rel', re.max(), 'genes differing', np.where(re.max(1)>1e-9)[0])
print('se max rel', rs.max(), 'genes differing >1e-6', np.where(rs.max(1)>1e-6)[0], ' >1e-9', np.where(rs.max(1)>1e-9)[0].size)
for k in ("maxcor", "lambda.max", "lambda.min", "sd.cv"):
    d = np.abs(a.info[k]-b.info[k])/np.maximum(np.abs(a.info[k]),1e-300)
    print(k, np.nanmax(d), np.where(d>1e-9)[0])
