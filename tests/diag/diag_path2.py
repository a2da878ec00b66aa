import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo')
import oracle, saver_b200
from saver_b200 import ops
d = np.load('/root/repo/tests/golden/linnarsson.npz')
x = d['counts'].astype(np.float64); sf = d['size_factor']; y = x / sf
xest = np.ascontiguousarray(np.log((x + 1) / sf))
lm = d['lambda_max']
h = saver_b200.default_handle(0)
ops.set_design(xest, handle=h)
L = oracle.lib()
for g in (9, 174, 0, 379):
    got = ops.predict_path(y[g:g+1], lm[g:g+1], d['lambda_min'][g:g+1], self_col=np.array([g]), want_stats=True, handle=h)
    a0 = np.zeros(128); dev = np.zeros(128); lam = np.zeros(128); nin = np.zeros(128, np.int32); nlp = np.zeros(128, np.int32)
    h.check(h.lib.saver_cuda_debug_path(h.h, 0, ops.ptr(a0), ops.ptr(dev), ops.ptr(nin), ops.ptr(nlp), ops.ptr(lam)))
    ref = oracle.predict_path(xest, y[g], lm[g], d['lambda_min'][g], excl=g)
    onlp = np.zeros(128, np.int32); ongr = np.zeros(128, np.int32); onin = np.zeros(128, np.int32)
    L.orc_debug_last_fit(ops.ptr(onlp), ops.ptr(ongr), ops.ptr(onin))
    n = ref['lmu']
    print('gene', g, 'lmu', n, got['stats'][0])
    print('  gpu nlp', nlp[:n], 'nin', nin[:n])
    print('  ref nlp', onlp[:n], 'nin', onin[:n], 'ngrad', ongr[:n])
    print('  gpu dev', dev[:n])
