"""Thin Python calls over the C ABI (one function per export of include/saver_cuda.h).

Panels are passed gene-major: a C-contiguous numpy array of shape (B, n) is exactly the
n x B column-major panel of the C side.  torch CUDA tensors of the same shape are accepted
wherever the ABI has ``device_ptrs`` (everything stays on the device then).
"""
import ctypes as C

import numpy as np

from ._lib import default_handle, is_device, ptr

POST_INFO = 8
_D = C.c_double


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _vec(v, n):
    v = np.atleast_1d(np.asarray(v, dtype=np.float64))
    if v.size not in (1, n):
        raise ValueError("length must be 1 or %d" % n)
    return np.ascontiguousarray(v)


def loglik(model, param, y, mu, sf, handle=None):
    """calc.loglik.a/b/k (R/calc_loglik.R:26,44,61).  Scalars mu/sf are recycled."""
    h = handle or default_handle()
    y = _f64(y)
    n = y.size
    mu, sf = _vec(mu, n), _vec(sf, n)
    out = _D(0.0)
    h.check(h.lib.saver_cuda_loglik(h.h, int(model), _D(float(param)), ptr(y), ptr(mu), mu.size,
                                    ptr(sf), sf.size, n, C.byref(out)))
    return out.value


def calc_var(model, y, mu, sf, want_grid=False, handle=None):
    """calc.a / calc.b / calc.k (R/optimize_variance.R) -> (param, nll, flags[, grid])."""
    h = handle or default_handle()
    y = _f64(y)
    n = y.size
    mu, sf = _vec(mu, n), _vec(sf, n)
    out = np.zeros(2)
    flags = C.c_int(0)
    grid = np.zeros(200) if want_grid else None
    h.check(h.lib.saver_cuda_calc_var(h.h, int(model), ptr(y), ptr(mu), mu.size, ptr(sf), sf.size,
                                      n, ptr(out), C.byref(flags), ptr(grid)))
    if want_grid:
        return out[0], out[1], flags.value, grid
    return out[0], out[1], flags.value


def calc_post(Y, MU, sf, scale_sf=1.0, want_se=True, want_raw=False, want_info=True,
              want_grid=False, out=None, handle=None):
    """calc.post (R/calc_posterior.R:26) for a panel: Y (B, n) raw counts, MU (B, n) or (B,).

    Returns dict(est, se, est_raw, se_raw, info, grid); absent pieces are None.  With torch
    CUDA tensors as inputs the outputs are torch CUDA tensors and nothing is copied.
    """
    h = handle or default_handle()
    dev = is_device(Y)
    if dev:
        import torch
        B, n = Y.shape
        scalar = MU.dim() == 1
        mk = lambda *s: torch.empty(*s, dtype=torch.float64, device=Y.device)  # noqa: E731
        assert Y.is_contiguous() and MU.is_contiguous() and sf.is_contiguous()
        assert Y.dtype == torch.float64 and MU.dtype == torch.float64 and sf.dtype == torch.float64
    else:
        Y = _f64(np.atleast_2d(Y))
        B, n = Y.shape
        MU = _f64(MU)
        scalar = MU.ndim == 1 and not (B == 1 and MU.size == n and n != 1)
        if not scalar:
            MU = _f64(MU.reshape(B, n))
        sf = np.ascontiguousarray(np.broadcast_to(_vec(sf, n), (n,)))
        mk = lambda *s: np.empty(s, dtype=np.float64)  # noqa: E731
    res = out or {}
    est = res.get("est") if res.get("est") is not None else mk(B, n)
    se = (res.get("se") if res.get("se") is not None else mk(B, n)) if want_se else None
    er = mk(B, n) if want_raw else None
    sr = mk(B, n) if want_raw else None
    info = mk(B, POST_INFO) if want_info else None
    grid = mk(B, 600) if want_grid else None
    h.check(h.lib.saver_cuda_calc_post(h.h, ptr(Y), ptr(MU), int(scalar), ptr(sf), n, B,
                                       _D(float(scale_sf)), ptr(est), ptr(se), ptr(er), ptr(sr),
                                       ptr(info), ptr(grid), int(dev)))
    return dict(est=est, se=se, est_raw=er, se_raw=sr, info=info, grid=grid)


def set_design(xest, handle=None):
    """Install x.est (R/saver.R:156-157).  xest: (p, n) gene-major == n x p column-major."""
    h = handle or default_handle()
    dev = is_device(xest)
    if not dev:
        xest = _f64(xest)
    p, n = xest.shape
    h.check(h.lib.saver_cuda_set_design(h.h, ptr(xest), n, p, int(dev)))
    h.design_shape = (n, p)


def maxcor(Y, self_col=None, handle=None):
    """calc.maxcor(x.est, t(y)) (R/calc_maxcor.R:26).  Y: (B, n); self_col: (B,) or None."""
    h = handle or default_handle()
    dev = is_device(Y)
    if dev:
        import torch
        B, n = Y.shape
        out = torch.empty(B, dtype=torch.float64, device=Y.device)
        sc = self_col
    else:
        Y = _f64(np.atleast_2d(Y))
        B, n = Y.shape
        out = np.empty(B)
        sc = None if self_col is None else np.ascontiguousarray(self_col, dtype=np.int32)
    if n != h.design_shape[0]:
        raise ValueError("Y has %d cells, design has %d" % (n, h.design_shape[0]))
    h.check(h.lib.saver_cuda_maxcor(h.h, ptr(Y), B, ptr(sc), ptr(out), int(dev)))
    return out


NLAM_CAP = 128


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def predict_cv(Y, foldid, self_col=None, nfolds=5, dfmax=300, nlambda=100, thresh=1e-7,
               want_stats=False, want_cv=False, handle=None):
    """expr.predict, CV variant (R/expr_predict.R:39-60) for a panel.

    Y (B, n) normalised expression; foldid (B, n) ints: 0 = not a pred cell, 1..nfolds.
    Returns dict(mu (B, n), lambda_max, lambda_min, sd_cv, idx, status[, stats, cv]).
    """
    h = handle or default_handle()
    dev = is_device(Y)
    if dev:
        import torch
        B, n = Y.shape
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=Y.device)  # noqa: E731
        f64, i32 = torch.float64, torch.int32
        sc = self_col
        assert foldid.dtype == torch.int32 and foldid.is_contiguous() and Y.is_contiguous()
    else:
        Y = _f64(np.atleast_2d(Y))
        B, n = Y.shape
        foldid = _i32(foldid).reshape(B, n)
        sc = None if self_col is None else _i32(self_col)
        mk = lambda shape, dt: np.empty(shape, dtype=dt)  # noqa: E731
        f64, i32 = np.float64, np.int32
    if n != h.design_shape[0]:
        raise ValueError("Y has %d cells, design has %d" % (n, h.design_shape[0]))
    mu = mk((B, n), f64)
    fit4 = mk((B, 4), f64)
    status = mk((B,), i32)
    stats = mk((B, 4), i32) if want_stats else None
    cv = mk((B, 3, NLAM_CAP), f64) if want_cv else None
    h.check(h.lib.saver_cuda_predict_cv(h.h, ptr(Y), B, ptr(sc), ptr(foldid), int(nfolds), int(dfmax),
                                        int(nlambda), _D(float(thresh)), ptr(mu), ptr(fit4),
                                        ptr(status), ptr(stats), ptr(cv), int(dev)))
    return dict(mu=mu, lambda_max=fit4[:, 0], lambda_min=fit4[:, 1], sd_cv=fit4[:, 2],
                idx=fit4[:, 3], status=status, stats=stats, cv=cv)


def predict_path(Y, lambda_max, lambda_min, self_col=None, pred_mask=None, dfmax=300,
                 thresh=1e-7, want_stats=False, handle=None):
    """expr.predict, fixed-lambda variant (R/expr_predict.R:61-82) for a panel."""
    h = handle or default_handle()
    dev = is_device(Y)
    lmax = _f64(np.asarray(lambda_max, dtype=np.float64))
    lmin = _f64(np.asarray(lambda_min, dtype=np.float64))
    if dev:
        import torch
        B, n = Y.shape
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=Y.device)  # noqa: E731
        f64, i32 = torch.float64, torch.int32
        sc, pm = self_col, pred_mask
    else:
        Y = _f64(np.atleast_2d(Y))
        B, n = Y.shape
        sc = None if self_col is None else _i32(self_col)
        pm = None if pred_mask is None else _i32(pred_mask)
        mk = lambda shape, dt: np.empty(shape, dtype=dt)  # noqa: E731
        f64, i32 = np.float64, np.int32
    if n != h.design_shape[0]:
        raise ValueError("Y has %d cells, design has %d" % (n, h.design_shape[0]))
    if lmax.size != B or lmin.size != B:
        raise ValueError("lambda_max / lambda_min need one value per gene")
    mu = mk((B, n), f64)
    status = mk((B,), i32)
    stats = mk((B, 4), i32) if want_stats else None
    h.check(h.lib.saver_cuda_predict_path(h.h, ptr(Y), B, ptr(sc), ptr(pm), ptr(lmax), ptr(lmin),
                                          int(dfmax), _D(float(thresh)), ptr(mu), ptr(status),
                                          ptr(stats), int(dev)))
    return dict(mu=mu, status=status, stats=stats)


def calc_estimate(x, sf, scale_sf, mode, cutoff=0.0, coefs=None, self_col=None, is_pred=None,
                  pred_mask=None, foldid=None, nfolds=5, calc_maxcor=False, want_se=True,
                  want_mu=False, dfmax=300, nlambda=100, thresh=1e-7, handle=None):
    """calc.estimate body (R/calc_estimate.R:75-141) for one block: x (B, n) raw counts.

    mode 0 null.model / 1 cross-validated lasso / 2 est.lambda + fixed-lambda path.
    Returns dict(est, se, maxcor, lambda_max, lambda_min, sd_cv, ct, vt, status[, mu])."""
    h = handle or default_handle()
    dev = is_device(x)
    if dev:
        import torch
        B, n = x.shape
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=x.device)  # noqa: E731
        f64, i32 = torch.float64, torch.int32
    else:
        x = _f64(np.atleast_2d(x))
        B, n = x.shape
        sf = _f64(np.broadcast_to(_vec(sf, n), (n,)))
        self_col = None if self_col is None else _i32(self_col)
        is_pred = None if is_pred is None else _i32(is_pred)
        pred_mask = None if pred_mask is None else _i32(pred_mask)
        foldid = None if foldid is None else _i32(foldid).reshape(B, n)
        mk = lambda shape, dt: np.empty(shape, dtype=dt)  # noqa: E731
        f64, i32 = np.float64, np.int32
    cf = None if coefs is None else _f64(np.asarray(coefs, dtype=np.float64))
    est = mk((B, n), f64)
    se = mk((B, n), f64) if want_se else None
    info = mk((6, B), f64)
    mu = mk((B, n), f64) if want_mu else None
    status = mk((B,), i32)
    h.check(h.lib.saver_cuda_calc_estimate(
        h.h, ptr(x), n, B, ptr(sf), _D(float(scale_sf)), ptr(self_col), ptr(is_pred),
        ptr(pred_mask), ptr(foldid), int(nfolds), int(mode), int(bool(calc_maxcor)),
        _D(float(cutoff)), ptr(cf), int(dfmax), int(nlambda), _D(float(thresh)), ptr(est), ptr(se),
        ptr(info), ptr(mu), ptr(status), int(dev)))
    return dict(est=est, se=se, maxcor=info[0], lambda_max=info[1], lambda_min=info[2],
                sd_cv=info[3], ct=info[4], vt=info[5], status=status, mu=mu)


def sample(est, se, nb_mode=False, seed=0, rep_idx=0, gene0=0, handle=None):
    """sample.saver draws for a panel: est, se (B, n) -> (B, n)."""
    h = handle or default_handle()
    dev = is_device(est)
    if dev:
        import torch
        out = torch.empty_like(est)
    else:
        est, se = _f64(np.atleast_2d(est)), _f64(np.atleast_2d(se))
        out = np.empty_like(est)
    B, n = est.shape
    h.check(h.lib.saver_cuda_sample(h.h, ptr(est), ptr(se), n, B, C.c_longlong(int(gene0)),
                                    int(bool(nb_mode)), C.c_ulonglong(int(seed) & (2**64 - 1)),
                                    int(rep_idx), ptr(out), int(dev)))
    return out


def cor_adjust(est, se, by_cells=False, cor_in=None, col0=0, ncols=None, handle=None):
    """cor.genes / cor.cells: est, se (G, n) -> (ncols, M) rows = the requested columns of the
    symmetric adjusted correlation (M = G, or n when by_cells)."""
    h = handle or default_handle()
    dev = is_device(est)
    G, n = est.shape
    M = n if by_cells else G
    ncols = M - col0 if ncols is None else ncols
    if dev:
        import torch
        out = torch.empty((ncols, M), dtype=torch.float64, device=est.device)
    else:
        est, se = _f64(est), _f64(se)
        cor_in = None if cor_in is None else _f64(cor_in)
        out = np.empty((ncols, M))
    h.check(h.lib.saver_cuda_cor_adjust(h.h, ptr(est), ptr(se), n, G, int(bool(by_cells)),
                                        ptr(cor_in), int(col0), int(ncols), ptr(out), int(dev)))
    return out


def cor_prep(est, se, handle=None):
    """First half of cor.genes for a block of genes: est, se (B, n) -> Z (B, ldz) centred unit-norm
    rows (ldz = n rounded up to 16, zero padded), adj (B).  Device tensors stay on the device."""
    h = handle or default_handle()
    dev = is_device(est)
    B, n = est.shape
    ldz = (n + 15) // 16 * 16
    if dev:
        import torch
        Z = torch.empty((B, ldz), dtype=torch.float64, device=est.device)
        adj = torch.empty((B,), dtype=torch.float64, device=est.device)
    else:
        est, se = _f64(est), _f64(se)
        Z, adj = np.empty((B, ldz)), np.empty(B)
    h.check(h.lib.saver_cuda_cor_prep(h.h, ptr(est), ptr(se), n, B, ptr(Z), ldz, ptr(adj), int(dev)))
    return Z, adj


def cor_block(Z, adj, M, row0, nrows, col0, ncols, out, handle=None):
    """Second half: out[c, row0:row0+nrows] (out: (ncols, >= row0+nrows) device tensor, one row per
    column c of [col0, col0+ncols)) = adjusted correlation block from the gathered panel Z
    ((>= M + 128, ldz) device tensor, rows beyond M zero), adj (M)."""
    h = handle or default_handle()
    if not (is_device(Z) and is_device(adj) and is_device(out)):
        raise ValueError("cor_block works on device tensors")
    if Z.shape[0] < M + 128:
        raise ValueError("Z needs M + 128 rows (zero filled beyond M): the GEMM reads whole tiles")
    h.check(h.lib.saver_cuda_cor_block(h.h, ptr(Z), int(Z.shape[1]), ptr(adj), int(M), int(row0),
                                       int(nrows), int(col0), int(ncols), ptr(out), int(out.shape[1])))
    return out


def set_counts_csc(x, handle=None):
    """Upload a scipy.sparse genes x cells matrix (kept resident; R's dgCMatrix layout)."""
    h = handle or default_handle()
    x = x.tocsc()
    indptr = np.ascontiguousarray(x.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(x.indices, dtype=np.int32)
    data = np.ascontiguousarray(x.data, dtype=np.float64)
    G, n = x.shape
    h.check(h.lib.saver_cuda_set_counts_csc(h.h, ptr(indptr), ptr(indices), ptr(data), G, n))
    h.csc_shape = (G, n)


def csc_sums(handle=None):
    h = handle or default_handle()
    G, n = h.csc_shape
    cs, rs = np.zeros(n), np.zeros(G)
    h.check(h.lib.saver_cuda_csc_sums(h.h, ptr(cs), ptr(rs)))
    return cs, rs


def csc_gather(gene_idx, sf=None, transform=False, device=False, handle=None):
    """Dense (B, n) block of the resident sparse counts; log((x + 1) / sf) when transform."""
    h = handle or default_handle()
    G, n = h.csc_shape
    idx = np.ascontiguousarray(gene_idx, dtype=np.int32)
    B = idx.size
    sfv = None if sf is None else _f64(sf)
    if device:
        import torch
        out = torch.empty((B, n), dtype=torch.float64, device=torch.device("cuda", h.device))
    else:
        out = np.empty((B, n))
    h.check(h.lib.saver_cuda_csc_gather(h.h, ptr(idx), B, ptr(sfv), int(bool(transform)), ptr(out),
                                        int(bool(device))))
    return out


def set_counts_dense(x, handle=None):
    """Upload a dense genes x cells count matrix (float64 or int32, C-contiguous) and keep it resident
    (saver_cuda_set_counts_dense): the whole-matrix passes of saver() then run on the device."""
    h = handle or default_handle()
    if not (isinstance(x, np.ndarray) and x.ndim == 2):
        raise ValueError("x must be a 2-D numpy array")
    if x.dtype == np.int32:
        dtype = 1
    else:
        x = _f64(x)
        dtype = 0
    x = np.ascontiguousarray(x)
    G, n = x.shape
    h.check(h.lib.saver_cuda_set_counts_dense(h.h, ptr(x), dtype, G, n))
    h.dense_shape = (G, n)


def dense_sums(handle=None):
    h = handle or default_handle()
    G, n = h.dense_shape
    cs, rs = np.zeros(n), np.zeros(G)
    h.check(h.lib.saver_cuda_dense_sums(h.h, ptr(cs), ptr(rs)))
    return cs, rs


def dense_gather(gene_idx, sf=None, transform=False, device=False, handle=None):
    """Dense (B, n) block of the resident dense counts; log((x + 1) / sf) when transform."""
    h = handle or default_handle()
    G, n = h.dense_shape
    idx = np.ascontiguousarray(gene_idx, dtype=np.int32)
    B = idx.size
    sfv = None if sf is None else _f64(sf)
    if device:
        import torch
        out = torch.empty((B, n), dtype=torch.float64, device=torch.device("cuda", h.device))
    else:
        out = np.empty((B, n))
    h.check(h.lib.saver_cuda_dense_gather(h.h, ptr(idx), B, ptr(sfv), int(bool(transform)), ptr(out),
                                          int(bool(device))))
    return out


def release_counts(handle=None):
    h = handle or default_handle()
    h.check(h.lib.saver_cuda_release_counts(h.h))
