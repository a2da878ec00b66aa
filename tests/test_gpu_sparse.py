"""Sparse input (R: dgCMatrix): the resident CSC matrix and its device-side
densification against scipy, and saver(sparse) against saver(dense) -- bit for bit, since both
feed the same count panels to the same kernels (R/saver.R:156-157, R/calc_estimate.R:69)."""
import numpy as np
import pytest
import scipy.sparse as sp

import saver_b200
from saver_b200 import ops, synth

pytestmark = pytest.mark.gpu


def _counts(G, n, seed, density=0.15):
    rng = np.random.default_rng(seed)
    x = rng.poisson(2.0, (G, n)).astype(np.float64) * (rng.random((G, n)) < density)
    x[3] = 0                      # an all-zero gene
    x[5, ::7] = 0.0004            # below clean.data's 0.001 threshold -> treated as zero
    return x


@pytest.mark.parametrize("G,n", [(64, 33), (257, 1000), (40, 9001)])
def test_csc_sums_and_gather_match_scipy(handle, G, n):
    x = _counts(G, n, G + n)
    ops.set_counts_csc(sp.csc_matrix(x), handle=handle)
    xc = np.where(x < 0.001, 0.0, x)
    cs, rs = ops.csc_sums(handle=handle)
    assert np.array_equal(cs, xc.sum(0)) and np.array_equal(rs, xc.sum(1))
    idx = np.array([G - 1, 3, 5, 0, 17], dtype=np.int32)
    assert np.array_equal(ops.csc_gather(idx, handle=handle), xc[idx])
    sf = np.random.default_rng(1).uniform(0.3, 3.0, n)
    want = np.log((xc[idx] + 1.0) / sf)
    got = ops.csc_gather(idx, sf=sf, transform=True, handle=handle)
    assert np.allclose(got, want, rtol=1e-15, atol=1e-15)
    dev = ops.csc_gather(idx, sf=sf, transform=True, device=True, handle=handle)
    assert np.array_equal(dev.cpu().numpy(), got)
    assert ops.csc_gather(np.zeros(0, dtype=np.int32), handle=handle).shape == (0, n)


def test_csc_argument_errors(handle):
    x = sp.csc_matrix(_counts(20, 12, 0))
    ops.set_counts_csc(x, handle=handle)
    with pytest.raises(saver_b200.SaverCudaError):
        ops.csc_gather(np.array([20], dtype=np.int32), handle=handle)
    with pytest.raises(saver_b200.SaverCudaError):
        ops.csc_gather(np.array([-1], dtype=np.int32), handle=handle)


def test_empty_matrix_has_zero_sums(handle):
    ops.set_counts_csc(sp.csc_matrix((6, 9)), handle=handle)
    cs, rs = ops.csc_sums(handle=handle)
    assert not cs.any() and not rs.any()
    assert not ops.csc_gather(np.arange(6, dtype=np.int32), handle=handle).any()


@pytest.mark.parametrize("kw", [dict(do_fast=False), dict(do_fast=True, perm=3),
                                dict(null_model=True), dict(do_fast=False, size_factor=1),
                                dict(do_fast=False, pred_genes=np.array([2, 9, 30]),
                                     pred_genes_only=True)])
def test_saver_sparse_equals_dense(kw):
    x = synth.nb_counts(150, 90, seed=2).astype(np.float64)
    x[:, 11] = 0                                   # an empty cell: dropped by clean.data
    a = saver_b200.saver(x, fold_seed=5, **kw)
    b = saver_b200.saver(sp.csr_matrix(x), fold_seed=5, **kw)
    # the counts panels are identical; the design differs by the ulp between the device's log and
    # numpy's, which the 3-decimal rounding of the prediction absorbs (R/expr_predict.R:58,80).  The
    # solver's curvature model reads an FP32 copy of the design, where that ulp can become an FP32
    # ulp: the iterates then agree to ~1e-8 instead of 1e-15, which can move a value across one
    # step of the 3-decimal ceiling of estimate / se (R/calc_posterior.R:66-67) in a few cells.
    for u, v in ((a.estimate, b.estimate), (a.se, b.se)):
        d = np.abs(u - v)
        assert d.max() <= 1.0001e-3 and (d > 1e-9).mean() < 1e-3
    assert np.array_equal(a.info["size.factor"], b.info["size.factor"])
    for k in ("maxcor", "lambda.max", "lambda.min", "sd.cv"):
        assert np.allclose(a.info[k], b.info[k], rtol=1e-9, equal_nan=True), k


def test_saver_sparse_with_mu(handle):
    x = synth.nb_counts(60, 50, seed=4).astype(np.float64)
    mu = np.random.default_rng(0).gamma(2.0, 1.0, x.shape)
    a = saver_b200.saver(x, mu=mu)
    b = saver_b200.saver(sp.csc_matrix(x), mu=mu)
    assert np.array_equal(a.estimate, b.estimate) and np.array_equal(a.se, b.se)


@pytest.mark.parametrize("dtype", [np.float64, np.int32])
def test_dense_device_counts_match_numpy(handle, dtype):
    """The dense analogue (saver_cuda_set_counts_dense / dense_sums / dense_gather): clean.data's
    threshold, colSums, rowSums and the log-normalised rows against numpy."""
    x = _counts(130, 777, 5)
    if dtype is np.int32:
        x = np.floor(x)
    xd = x.astype(dtype)
    ops.set_counts_dense(xd, handle=handle)
    xc = np.where(xd < 0.001, 0.0, xd.astype(np.float64))
    cs, rs = ops.dense_sums(handle=handle)
    assert np.allclose(cs, xc.sum(0), rtol=1e-14) and np.allclose(rs, xc.sum(1), rtol=1e-14)
    idx = np.array([129, 3, 5, 0, 17], dtype=np.int32)
    assert np.array_equal(ops.dense_gather(idx, handle=handle), xc[idx])
    sf = np.random.default_rng(1).uniform(0.3, 3.0, 777)
    got = ops.dense_gather(idx, sf=sf, transform=True, device=True, handle=handle).cpu().numpy()
    assert np.allclose(got, np.log((xc[idx] + 1.0) / sf), rtol=1e-15, atol=1e-15)
    ops.release_counts(handle=handle)
    with pytest.raises(saver_b200.SaverCudaError):
        ops.dense_sums(handle=handle)


def test_saver_device_prep_equals_host_prep():
    """saver() on a matrix large enough for the device preparation path (one upload of x, design
    built on the device) against the host preparation: the design differs by the device log's last
    bit at most, the estimates by rounding."""
    x = synth.nb_counts(1100, 1000, seed=31)
    x[:, 17] = np.maximum(x[:, 17], 0)  # keep every cell non-empty
    pg = np.arange(5, 1100, 97)
    kw = dict(do_fast=False, pred_genes=pg, pred_genes_only=True, fold_seed=4, perm="identity")
    a = saver_b200.saver(x, device_prep=True, **kw)
    b = saver_b200.saver(x.astype(np.float64), device_prep=False, **kw)
    assert np.allclose(a.info["size.factor"], b.info["size.factor"], rtol=1e-14)
    assert np.array_equal(a.info["lambda.min"] > 0, b.info["lambda.min"] > 0)
    assert np.allclose(a.info["lambda.max"], b.info["lambda.max"], rtol=1e-9)
    rel = np.abs(a.estimate - b.estimate) / np.maximum(b.estimate, 1e-12)
    assert np.median(rel) < 1e-9 and np.quantile(rel, 0.99) < 2e-3   # one ceiling quantum at most
