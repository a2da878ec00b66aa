"""The reference-side binding (integration/src/rcall_shim.c + integration/R/saver_cuda.R), checked
without R: the shim must compile against include/saver_cuda.h (gcc type-checks every ABI call),
register every routine the R wrappers .Call with the right arity, and -- on a GPU -- produce
through the R calling convention (column-major matrices, t(x[block, ]) panels, NULL arguments,
external-pointer handle) exactly what the ctypes path produces."""
import os

import numpy as np
import pytest

import rshim
from saver_b200 import host, ops, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R_SRC = os.path.join(ROOT, "integration", "R", "saver_cuda.R")


@pytest.fixture(scope="module")
def R():
    s = rshim.Session()
    yield s
    s.close()


def test_shim_compiles_and_registers_what_the_r_wrappers_call(R):
    rshim.build(force=True)  # -Werror on implicit declarations / pointer and int conversions
    reg = R.routines()
    sites = rshim.r_call_sites(R_SRC)
    assert len(sites) >= 15
    for name, nargs in sites:
        assert name in reg, "%s is .Call-ed from saver_cuda.R but not registered by the shim" % name
        assert reg[name] == nargs, "%s: R passes %d arguments, the shim takes %d" % (name, nargs, reg[name])
    called = {n for n, _ in sites}
    assert called == set(reg), "registered but never called from R: %s" % (set(reg) - called)


def test_dot_call_arity_and_unknown_routine_are_errors(R):
    with pytest.raises(rshim.RError, match="takes 2 arguments"):
        R.call("C_saver_cuda_set_design", None)
    with pytest.raises(rshim.RError, match="no such .Call routine"):
        R.call("C_saver_cuda_nope", None)


def test_no_device_is_an_r_error_not_a_fallback(R):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rshim.RError, match="no usable CUDA device"):
        R.call("C_saver_cuda_create", 0)


# ---------------------------------------------------------------------------------------------
# GPU: the R calling convention end to end
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def rh(R):
    return R.call("C_saver_cuda_create", 0)


def _folds(B, n, seed):
    return host.draw_folds(np.arange(1, B + 1), n, 5, fold_seed=seed)


@pytest.mark.gpu
def test_calc_post_and_variance_through_dot_call(R, rh, handle):
    rng = np.random.default_rng(3)
    n = 300
    sf = rng.gamma(8.0, 1 / 8.0, n)
    mu = rng.gamma(2.0, 1.0, n) + 0.05
    y = rng.poisson(mu * sf).astype(np.float64)
    # calc.post <- function(y, mu, sf, scale.sf) .Call(C_saver_cuda_calc_post, h, y, mu, sf, scale.sf)
    est, se = R.call("C_saver_cuda_calc_post", rh, y, mu, sf, 1.0)
    ref = ops.calc_post(y[None], mu[None], sf, 1.0, handle=handle)
    assert np.array_equal(est, ref["est"][0]) and np.array_equal(se, ref["se"][0])
    # scalar mu is recycled (R/calc_posterior.R:28-33): length(mu) == 1
    est1, _ = R.call("C_saver_cuda_calc_post", rh, y, np.array([1.7]), sf, 1.0)
    ref1 = ops.calc_post(y[None], np.array([1.7]), sf, 1.0, handle=handle)
    assert np.allclose(est1, ref1["est"][0], rtol=1e-12)
    for model in (0, 1, 2):  # calc.a / calc.b / calc.k -> c(param, nll); calc.loglik.*
        pv = R.call("C_saver_cuda_calc_var", rh, model, y, mu, sf)
        par, nll, _ = ops.calc_var(model, y, mu, sf, handle=handle)
        assert pv.shape == (2,) and pv[0] == par and pv[1] == nll
        ll = R.call("C_saver_cuda_loglik", rh, model, 0.37, y, mu, sf)
        assert ll.shape == (1,) and ll[0] == ops.loglik(model, 0.37, y, mu, sf, handle=handle)


@pytest.mark.gpu
def test_calc_estimate_block_through_dot_call(R, rh, handle):
    """The body of calc.estimate in saver_cuda.R for one block, mode 1 (cv lasso with host folds),
    against the ctypes path on the same inputs."""
    x = synth.nb_counts(120, 180, seed=21).astype(np.float64)
    x = x[x.sum(1) > 0]
    G, n = x.shape
    sf = x.sum(0) / x.sum(0).mean()
    good = np.where(x.mean(1) >= 0.1)[0]
    x_est = np.log((x[good] + 1.0) / sf).T                      # n x p, as in R/saver.R:157
    R.call("C_saver_cuda_set_design", rh, x_est)
    ops.set_design(np.ascontiguousarray(x_est.T), handle=handle)
    blk = np.arange(0, 40)
    self_col = np.full(blk.size, -1, np.int32)
    pos = {g: i for i, g in enumerate(good)}
    for i, g in enumerate(blk):
        self_col[i] = pos.get(g, -1)
    fold = _folds(blk.size, n, 9)                                 # (B, n)
    out = R.call("C_saver_cuda_calc_estimate", rh, x[blk].T, sf, 1.0, self_col,
                 np.ones(blk.size, np.int32), np.ones(n, np.int32), fold.T, 1, False, 0.0, None, False)
    est_t, se_t, info = out
    assert est_t.shape == (n, blk.size) and se_t.shape == (n, blk.size) and info.shape == (blk.size, 6)
    ref = ops.calc_estimate(x[blk], sf, 1.0, 1, self_col=self_col, is_pred=np.ones(blk.size, np.int32),
                            pred_mask=np.ones(n, np.int32), foldid=fold, handle=handle)
    rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-12)  # noqa: E731
    assert np.median(rel(est_t.T, ref["est"])) < 1e-9 and np.median(rel(se_t.T, ref["se"])) < 1e-9
    assert np.allclose(info[:, 1], ref["lambda_max"], rtol=1e-12)
    assert np.mean(rel(info[:, 2], ref["lambda_min"]) < 1e-12) >= 0.9
    # estimates.only = TRUE: se comes back as a logical NA (R/calc_estimate.R:156)
    out = R.call("C_saver_cuda_calc_estimate", rh, x[blk].T, sf, 1.0, self_col,
                 np.ones(blk.size, np.int32), np.ones(n, np.int32), None, 0, False, 0.0, None, True)
    assert out[1].shape == (1,) and out[1][0] == rshim.NA_LOGICAL and np.isfinite(out[0]).all()
    # calc.maxcor(x.est, t(y)) with the design installed
    mc = R.call("C_saver_cuda_maxcor", rh, np.log((x[blk] + 1.0) / sf).T, self_col)
    assert np.allclose(mc, ops.maxcor(np.log((x[blk] + 1.0) / sf), self_col=self_col, handle=handle),
                       rtol=1e-12)


@pytest.mark.gpu
def test_sample_cor_and_sparse_through_dot_call(R, rh, handle):
    import scipy.sparse as sp
    rng = np.random.default_rng(8)
    G, n = 30, 50
    est = rng.gamma(2.0, 1.0, (G, n)) + 0.1
    se = rng.gamma(1.0, 0.2, (G, n)) + 0.01
    s = R.call("C_saver_cuda_sample", rh, est.T, se.T, False, 50.0, 0)
    assert np.array_equal(s.T, ops.sample(est, se, seed=50, handle=handle))
    cg = R.call("C_saver_cuda_cor_adjust", rh, est.T, se.T, False, None)
    # the ABI writes column blocks: ops returns them as rows, the R matrix holds them as columns
    assert np.array_equal(cg.T, ops.cor_adjust(est, se, handle=handle))
    assert np.allclose(cg, cg.T, rtol=0, atol=1e-14)
    cc = R.call("C_saver_cuda_cor_adjust", rh, est.T, se.T, True, None)
    assert cc.shape == (n, n) and np.array_equal(cc.T, ops.cor_adjust(est, se, by_cells=True, handle=handle))
    # dgCMatrix slots -> resident counts -> sums and dense / log-normalised row blocks
    x = rng.poisson(0.4, (G, n)).astype(np.float64)
    m = sp.csc_matrix(x)
    R.call("C_saver_cuda_set_counts_csc", rh, m.indptr, m.indices, m.data, np.array([G, n], np.int32))
    cs, rs = R.call("C_saver_cuda_csc_sums", rh, np.array([G, n], np.int32))
    assert np.array_equal(cs, x.sum(0)) and np.array_equal(rs, x.sum(1))
    genes0 = np.array([3, 0, 17], np.int32)
    rows = R.call("C_saver_cuda_csc_gather", rh, genes0, None, n)
    assert np.array_equal(rows.T, x[genes0])
    sf = np.maximum(cs, 1.0) / np.maximum(cs, 1.0).mean()
    lr = R.call("C_saver_cuda_csc_gather", rh, genes0, sf, n)
    assert np.allclose(lr.T, np.log((x[genes0] + 1.0) / sf), rtol=1e-12, atol=1e-13)
