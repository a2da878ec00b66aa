"""Per-stage wall time of saver(do.fast = TRUE): which calc.estimate blocks the schedule spends its time in."""
import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import saver_b200
from saver_b200 import synth, host

G, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4000, 5000)
x = synth.nb_counts(G, n, 20264).astype(np.float64)


from saver_b200 import ops


class Timed(host.CudaBackend):
    def counters(self):
        h = self.handle
        c = np.zeros(8); ph = np.zeros(16)
        h.lib.saver_cuda_counters(h.h, ops.ptr(c)); h.lib.saver_cuda_phase_cycles(h.h, ops.ptr(ph))
        return c, ph

    def calc_estimate(self, xb, *a, **k):
        c0, p0 = self.counters()
        t = time.time()
        r = super().calc_estimate(xb, *a, **k)
        self.handle.synchronize()
        c1, p1 = self.counters()
        d, ph = c1 - c0, p1 - p0
        mode = a[2] if len(a) > 2 else k.get('mode')
        print('  calc.estimate mode %s genes %5d maxcor %s: %.2f s | round kernels %.0f ms gemm %.0f ms rounds %d gemm cols %d'
              % (mode, xb.shape[0], k.get('calc_maxcor'), time.time() - t, d[0], d[1], d[6], d[5]), flush=True)
        if ph[7] > 0:
            names = ['phaseA', 'vectors', 'gramP', 'sweeps', 'material', 'inactive', 'relin']
            tot = ph[:7].sum()
            print('     ' + ' '.join('%s %.0f%%' % (n_, 100 * v / tot) for n_, v in zip(names, ph[:7])),
                  '| IRLS/gene %.0f steps/IRLS %.0f sweeps/IRLS %.1f checks/IRLS %.2f activations/IRLS %.2f (%.0f%% of time) mean m %.0f'
                  % (ph[7] / xb.shape[0], ph[8] / ph[7], ph[9] / ph[7], ph[11] / ph[7], ph[14] / ph[7],
                     100 * ph[13] / (tot + ph[13]), ph[12] / ph[7]), flush=True)
        return r


t = time.time()
res = saver_b200.saver(x, do_fast=True, fold_seed=1, perm=2, backend=Timed(), verbose=True)
print('total %.1f s' % (time.time() - t))
