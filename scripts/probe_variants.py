"""A/B of solver launch-plan variants on one cfg3-shaped gene block (run on the GPU box).

    python scripts/probe_variants.py GENES "K1=V1,K2=V2" "K1=V3" ...

Every variant = one fresh handle created under the given SAVER_* environment (the tuning is read once
per handle), the same genes and folds; prints round / GEMM time, phase shares and a checksum of the
results (variants that only change scheduling must agree bitwise).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saver_b200  # noqa: E402
from saver_b200 import host, ops, synth  # noqa: E402

SHAPE = os.environ.get("PROBE_SHAPE", "cfg3")


def main():
    G = int(sys.argv[1])
    variants = sys.argv[2:] or [""]
    Gt, n, seed = (20000, 10000, 20263) if SHAPE == "cfg3" else (5000, 2000, 20262)
    x0 = synth.nb_counts(Gt, n, seed)
    xc, _ = host.clean_data(x0)
    sf, scale_sf = host.calc_size_factor(xc, None, n)
    good = np.where(xc.mean(1) >= 0.1)[0]
    x_est = np.log((xc[good] + 1.0) / sf)
    self_col = np.full(xc.shape[0], -1, dtype=np.int32)
    self_col[good] = np.arange(good.size, dtype=np.int32)
    pred = np.where(xc.sum(1) > 0)[0]
    sel = pred[np.linspace(0, pred.size - 1, G).astype(int)]
    fold = host.draw_folds(np.arange(1, G + 1), n, 5, fold_seed=5)
    y = xc[sel] / sf
    base = None
    for v in variants:
        env = dict(kv.split("=") for kv in v.split(",") if kv)
        for k in list(os.environ):
            if k.startswith("SAVER_"):
                del os.environ[k]
        os.environ.update(env)
        h = saver_b200.Handle(0)
        ops.set_design(x_est, handle=h)
        h.synchronize()
        c0 = np.zeros(8)
        h.lib.saver_cuda_counters(h.h, ops.ptr(c0))
        t0 = time.time()
        r = ops.predict_cv(y, fold, self_col=self_col[sel], want_stats=True, handle=h)
        dt = time.time() - t0
        c = np.zeros(8)
        h.lib.saver_cuda_counters(h.h, ops.ptr(c))
        d = c - c0
        mu = np.asarray(r["mu"])
        chk = (float(mu.sum()), float(np.asarray(r["lambda_min"]).sum()))
        if base is None:
            base = mu
        rel = np.abs(mu - base) / np.maximum(np.abs(base), 1e-300)
        msg = "[%s] %d genes: wall %.2f s round %.0f ms gemm %.0f ms rounds %d gemmcols %d | sum(mu) %.12g " \
              "sum(lmin) %.12g | vs first variant: max rel %.3g median rel %.3g" % (
                  v, G, dt, d[0], d[1], d[6], d[5], chk[0], chk[1], rel.max(), np.median(rel))
        print(msg, flush=True)
        if int(env.get("SAVER_LASSO_DBG", "0")) & 128:
            ph = np.zeros(16)
            h.lib.saver_cuda_phase_cycles(h.h, ops.ptr(ph))
            names = ["phaseA", "vectors", "gramP", "sweeps", "material", "inactive", "relin"]
            tot = ph[:7].sum() + ph[13]
            print("   phases: " + " ".join("%s %.1f%%" % (k, 100 * q / tot) for k, q in zip(names, ph[:7]))
                  + " activate %.1f%% | IRLS/gene %.0f mean m %.0f inactive cols/IRLS %.0f act/IRLS %.2f fp64 rechecks/IRLS %.2f" % (
                      100 * ph[13] / tot, ph[7] / G, ph[12] / max(ph[7], 1), ph[10] / max(ph[7], 1),
                      ph[14] / max(ph[7], 1), ph[15] / max(ph[7], 1)), flush=True)
        h.close()
        del h


if __name__ == "__main__":
    main()
