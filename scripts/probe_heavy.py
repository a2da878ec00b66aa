"""Replay the genes on which the fixed-lambda path hit maxit (scripts/data/fastpath_heavy.npz, from
scripts/probe_fastpath.py 1200 2000) under solver variants."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saver_b200
from saver_b200 import host, ops, synth

d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "fastpath_heavy.npz"))
G, n = int(d["G"]), int(d["n"])
x = synth.nb_counts(G, n, 20264).astype(np.float64)
xc, _ = host.clean_data(x)
sf, scale = host.calc_size_factor(xc, None, n)
good = np.where(xc.mean(1) >= 0.1)[0]
x_est = np.log((xc[good] + 1.0) / sf)
self_col = np.full(G, -1, dtype=np.int32); self_col[good] = np.arange(good.size, dtype=np.int32)
genes = d["genes"]
y = xc[genes] / sf
base = None
for v in sys.argv[1:] or [""]:
    env = dict(kv.split("=") for kv in v.split(",") if kv)
    for k in list(os.environ):
        if k.startswith("SAVER_"):
            del os.environ[k]
    opts = {k[4:]: float(val) for k, val in env.items() if k.startswith("OPT_")}
    os.environ.update({k: val for k, val in env.items() if not k.startswith("OPT_")})
    h = saver_b200.Handle(0)
    for k, val in opts.items():
        h.set_option(k, val)
    ops.set_design(x_est, handle=h)
    t = time.time()
    r = ops.predict_path(y, d["lmax"], d["lmin"], self_col=self_col[genes], want_stats=True, handle=h)
    dt = time.time() - t
    st = np.asarray(r["stats"])
    mu = np.asarray(r["mu"])
    if base is None:
        base = mu
    print("[%s] %.2f s  lmu %s jerr %s sweeps %s  | max rel vs first %.3g" % (
        v, dt, st[:, 0].tolist(), st[:, 1].tolist(), st[:, 2].tolist(),
        (np.abs(mu - base) / np.maximum(base, 1e-300)).max()), flush=True)
    h.close()
