"""ctypes binding of the CPU oracle (``oracle/saver_oracle.c``).

TEST INFRASTRUCTURE ONLY.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  The product (``saver_b200/``) never imports it.

All matrices are passed gene-major the way the C-ABI of the product takes
them: a panel ``Y`` of B genes over n cells is a C-contiguous ``(B, n)``
array (= the n x B column-major panel of the C side).  The design ``x`` is a
C-contiguous ``(p, n)`` array (= n x p column-major ``x.est`` of
/root/reference/R/saver.R:157).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile liboracle.so in-tree (gcc, a few seconds)."""
    src = os.path.join(_HERE, "saver_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_loglik.restype = C.c_double
        _lib.orc_maxcor.restype = C.c_double
        _lib.orc_loglik.argtypes = [C.c_int, C.c_double, _dp, _dp, _dp, C.c_int]
        _lib.orc_maxcor.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int]
    return _lib


def set_threads(outer=1, inner=1):
    """Host threads INSIDE one gene (see saver_oracle.c: orc_set_threads): `outer` fold fits at a
    time, each splitting its column loops over `inner` threads.  Results do not change."""
    lib().orc_set_threads(int(outer), int(inner))


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _expand(v, n):
    v = np.atleast_1d(np.asarray(v, dtype=np.float64))
    return _f64(np.full(n, v[0]) if v.size == 1 else v)


def loglik(model, param, y, mu, sf):
    """calc.loglik.a/b/k (model 0/1/2)."""
    y = _f64(y)
    n = y.size
    mu, sf = _expand(mu, n), _expand(sf, n)
    return lib().orc_loglik(int(model), float(param), _d(y), _d(mu), _d(sf), n)


def calc_var(model, y, mu, sf, want_grid=False):
    """calc.a/b/k -> dict(param, nll, rc, flags, nevals[, grid_x, grid_ll])."""
    y = _f64(y)
    n = y.size
    mu, sf = _expand(mu, n), _expand(sf, n)
    out = np.zeros(2)
    flags, nev = C.c_int(0), C.c_int(0)
    gx = np.full(100, np.nan) if want_grid else None
    gl = np.full(100, np.nan) if want_grid else None
    rc = lib().orc_calc_var(int(model), _d(y), _d(mu), _d(sf), n, _d(out), C.byref(flags),
                            C.byref(nev), _d(gx), _d(gl))
    res = dict(param=out[0], nll=out[1], rc=rc, flags=flags.value, nevals=nev.value)
    if want_grid:
        res.update(grid_x=gx, grid_ll=gl)
    return res


def calc_post(y, mu, sf, scale_sf=1.0):
    """calc.post for one gene -> dict(est, se, est_raw, se_raw, info, rc)."""
    y = _f64(y)
    n = y.size
    mu, sf = _expand(mu, n), _expand(sf, n)
    est, se, er, sr = (np.zeros(n) for _ in range(4))
    info = np.zeros(6)
    rc = lib().orc_calc_post(_d(y), _d(mu), _d(sf), n, C.c_double(scale_sf), _d(est), _d(se),
                             _d(er), _d(sr), _d(info))
    return dict(est=est, se=se, est_raw=er, se_raw=sr, info=info, rc=rc)


def _pool_map(fn, count, threads):
    """ctypes drops the GIL during foreign calls: plain threads scale."""
    from concurrent.futures import ThreadPoolExecutor
    threads = threads or os.cpu_count() or 1
    if threads <= 1 or count <= 1:
        return [fn(g) for g in range(count)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(fn, range(count)))


def calc_post_batch(Y, MU, sf, scale_sf=1.0, threads=None):
    """Y, MU: (B, n).  Returns dict of (B, n) est/se/est_raw/se_raw, info (B, 6), rc (B,)."""
    Y, MU = _f64(Y), _f64(MU)
    B, n = Y.shape
    sf = _expand(sf, n)
    est, se, er, sr = (np.zeros((B, n)) for _ in range(4))
    info = np.zeros((B, 6))
    rc = np.zeros(B, dtype=np.int32)
    L = lib()

    def one(g):
        rc[g] = L.orc_calc_post(_d(Y[g]), _d(MU[g]), _d(sf), n, C.c_double(scale_sf), _d(est[g]),
                                _d(se[g]), _d(er[g]), _d(sr[g]), _d(info[g]))

    _pool_map(one, B, threads)
    return dict(est=est, se=se, est_raw=er, se_raw=sr, info=info, rc=rc)


def maxcor(x, Y, self_col=None, threads=None):
    """calc.maxcor: x (p, n) design, Y (B, n) normalised genes, self_col (B,) or None."""
    x, Y = _f64(x), _f64(np.atleast_2d(Y))
    p, n = x.shape
    B = Y.shape[0]
    sc = np.full(B, -1, dtype=np.int32) if self_col is None else _i32(self_col)
    out = np.zeros(B)
    L = lib()

    def one(g):
        out[g] = L.orc_maxcor(_d(x), n, p, _d(Y[g]), int(sc[g]))

    _pool_map(one, B, threads)
    return out


def est_lambda(y, r, coefs):
    y = _f64(y)
    coefs = _f64(coefs)
    out = np.zeros(2)
    lib().orc_est_lambda(_d(y), y.size, C.c_double(r), _d(coefs), _d(out))
    return out


def glmnet_poisson(x, y, rows=None, excl=-1, lam=None, dfmax=300, thresh=1e-7):
    """One glmnet(x[rows,], y[rows], family='poisson', dfmax[, lambda]) fit."""
    x, y = _f64(x), _f64(y)
    p, n = x.shape
    rows = _i32(np.arange(n) if rows is None else rows)
    nl = 100 if lam is None else len(lam)
    lamv = None if lam is None else _f64(lam)
    lo, a0, dev = np.zeros(nl), np.zeros(nl), np.zeros(nl)
    beta = np.zeros((nl, p))
    st = np.zeros(4, dtype=np.int32)
    lib().orc_glmnet_poisson(_d(x), n, p, _i(rows), rows.size, _d(y), int(excl), _d(lamv), nl,
                             int(dfmax), C.c_double(thresh), _d(lo), _d(a0), _d(beta), _d(dev),
                             _i(st))
    L = int(st[0])
    return dict(lam=lo[:L], a0=a0[:L], beta=beta[:L], dev=dev[:L], lmu=L, jerr=int(st[1]),
                nlp=int(st[2]), ngrad=int(st[3]))


def predict_cv(x, y, foldid, rows=None, excl=-1, nfolds=5, dfmax=300, thresh=1e-7):
    """expr.predict CV variant for one gene (foldid aligned with rows, values 1..nfolds)."""
    x, y = _f64(x), _f64(y)
    p, n = x.shape
    rows = _i32(np.arange(n) if rows is None else rows)
    foldid = _i32(foldid)
    mu, out4, cv = np.zeros(n), np.zeros(4), np.zeros(300)
    st = np.zeros(4, dtype=np.int32)
    rc = lib().orc_predict_cv(_d(x), n, p, _d(y), _i(rows), rows.size, _i(foldid), int(nfolds),
                              int(excl), int(dfmax), C.c_double(thresh), _d(mu), _d(out4), _d(cv),
                              _i(st))
    return dict(mu=mu, lambda_max=out4[0], lambda_min=out4[1], sd_cv=out4[2], idx=int(out4[3]),
                lam=cv[:100], cvm=cv[100:200], cvsd=cv[200:], rc=rc, lmu=int(st[0]),
                jerr=int(st[1]), nlp=int(st[2]), ngrad=int(st[3]))


def predict_path(x, y, lmax, lmin, rows=None, excl=-1, dfmax=300, thresh=1e-7):
    """expr.predict fixed-lambda variant for one gene."""
    x, y = _f64(x), _f64(y)
    p, n = x.shape
    rows = _i32(np.arange(n) if rows is None else rows)
    mu = np.zeros(n)
    st = np.zeros(4, dtype=np.int32)
    rc = lib().orc_predict_path(_d(x), n, p, _d(y), _i(rows), rows.size, int(excl),
                                C.c_double(lmax), C.c_double(lmin), int(dfmax),
                                C.c_double(thresh), _d(mu), _i(st))
    return dict(mu=mu, rc=rc, lmu=int(st[0]), jerr=int(st[1]), nlp=int(st[2]), ngrad=int(st[3]))


def predict_cv_batch(x, Y, foldid, self_col=None, rows=None, nfolds=5, dfmax=300, thresh=1e-7,
                     threads=None):
    """Batched CV prediction.  Y (B, n); foldid (B, m)."""
    Y = _f64(np.atleast_2d(Y))
    B = Y.shape[0]
    sc = np.full(B, -1, dtype=np.int32) if self_col is None else _i32(self_col)
    res = _pool_map(lambda g: predict_cv(x, Y[g], foldid[g], rows, int(sc[g]), nfolds, dfmax,
                                         thresh), B, threads)
    keys = ("lambda_max", "lambda_min", "sd_cv", "idx", "rc", "lmu", "jerr", "nlp", "ngrad")
    out = {k: np.array([r[k] for r in res]) for k in keys}
    out["mu"] = np.stack([r["mu"] for r in res])
    return out


def predict_path_batch(x, Y, lmax, lmin, self_col=None, rows=None, dfmax=300, thresh=1e-7,
                       threads=None):
    Y = _f64(np.atleast_2d(Y))
    B = Y.shape[0]
    sc = np.full(B, -1, dtype=np.int32) if self_col is None else _i32(self_col)
    res = _pool_map(lambda g: predict_path(x, Y[g], lmax[g], lmin[g], rows, int(sc[g]), dfmax,
                                           thresh), B, threads)
    out = {k: np.array([r[k] for r in res]) for k in ("rc", "lmu", "jerr", "nlp", "ngrad")}
    out["mu"] = np.stack([r["mu"] for r in res])
    return out
