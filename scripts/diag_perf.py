"""Round-kernel accounting on a slice of cfg2: slot utilisation, cycles per column pass."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import saver_b200
from saver_b200 import ops, synth, host
G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = synth.nb_counts(5000, 2000, 20262).astype(np.float64)
sf = x.sum(0) / x.sum(0).mean()
good = np.where(x.mean(1) >= 0.1)[0]
xest = np.log((x[good] + 1) / sf)
h = saver_b200.default_handle(0)
ops.set_design(xest, handle=h)
sel = np.arange(G)
self_col = np.full(5000, -1, np.int32); self_col[good] = np.arange(good.size)
fold = host.draw_folds(sel + 1, 2000, 5, 5)
for it in range(2):
    c0 = np.zeros(8); e0 = np.zeros(4)
    h.lib.saver_cuda_counters(h.h, ops.ptr(c0)); h.lib.saver_cuda_counters_ext(h.h, ops.ptr(e0))
    t = time.time()
    r = ops.predict_cv(x[sel] / sf, fold, self_col=self_col[sel], want_stats=True, handle=h)
    dt = time.time() - t
    c1 = np.zeros(8); e1 = np.zeros(4)
    h.lib.saver_cuda_counters(h.h, ops.ptr(c1)); h.lib.saver_cuda_counters_ext(h.h, ops.ptr(e1))
    d = c1 - c0
    cyc = e1[0] - e0[0]
    print('genes %d wall %.2fs round %.1fms gemm %.1fms setup %.1fms rounds %d cols %.3g gemmcols %d' % (G, dt, d[0], d[1], d[2], d[6], d[4], d[5]))
    ph = np.zeros(16); h.lib.saver_cuda_phase_cycles(h.h, ops.ptr(ph))
    names = ['phaseA', 'vectors', 'gramP', 'sweeps', 'material', 'inactive', 'relin']
    if it == 1:
        d_ph = ph - ph_prev
        tot = d_ph[:7].sum()
        print('  phases:', ' '.join('%s %.1f%%' % (n_, 100 * v / tot) for n_, v in zip(names, d_ph[:7])), 'IRLS/gene %.0f' % (d_ph[7] / G), 'cyc/IRLS %.0f' % (tot / d_ph[7]))
        print('  steps/IRLS %.0f sweeps/IRLS %.1f cyc/step %.0f | inactive cols/IRLS %.0f checks/IRLS %.2f cyc/col %.0f | mean m %.0f' % (d_ph[8]/d_ph[7], d_ph[9]/d_ph[7], d_ph[3]/max(d_ph[8],1), d_ph[10]/d_ph[7], d_ph[11]/d_ph[7], d_ph[5]/max(d_ph[10],1), d_ph[12]/d_ph[7]))
        print('  activations/IRLS %.2f cyc/activation %.0f (%.1f%% of total)' % (d_ph[14]/d_ph[7], d_ph[13]/max(d_ph[14],1), 100*d_ph[13]/(tot+d_ph[13])), 'first part cyc %.0f' % (d_ph[15]/max(d_ph[14],1)))
    ph_prev = ph
    hist = np.zeros(20); h.lib.saver_cuda_sweep_hist(h.h, ops.ptr(hist))
    if hist.sum() > 0: print('  sweep steps by m/32: ' + ' '.join('%.1f' % (100 * v / hist.sum()) for v in hist))
    print('  gram TFLOP %.3g ;' % (2 * (e1[2] - e0[2]) / 1e12), end='')
    print('  cycles/pass (CTA lifetime) %.0f ; slot utilisation (2 CTA/SM) %.2f ; passes/gene %.3g ; sweeps/gene %.0f lmu mean %.1f' % (
        cyc / d[4], cyc / (d[0] * 1.965e6 * 296), d[4] / G, r['stats'][:, 2].mean(), r['stats'][:, 0].mean()))
