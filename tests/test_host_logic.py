"""CPU tests of the host mirror of the R API: utilities, argument validation, the stage schedule of
saver.fit (run through the oracle-backed test backend), and the C-ABI library's exports."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from helpers import OracleBackend
from saver_b200 import host, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_exports_every_declared_symbol():
    from saver_b200 import build
    lib = ctypes.CDLL(build.build_library())
    hdr = open(os.path.join(ROOT, "include", "saver_cuda.h")).read()
    names = re.findall(r"SAVER_API\s+[\w\s\*]+?\b(saver_cuda_\w+)\s*\(", hdr)
    assert len(names) >= 18
    for nm in names:
        assert hasattr(lib, nm), nm
    lib.saver_cuda_abi_version.restype = ctypes.c_int
    assert lib.saver_cuda_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import saver_b200
    with pytest.raises(saver_b200.SaverCudaError):
        saver_b200.Handle(0)


def test_utils_match_reference_rules():
    assert host.get_chunk(1000, 4) == 125 or host.get_chunk(1000, 4) in range(49, 151)
    # get.chunk: first chunk size in c(rbind(100:150, 99:49)) with ceil(len/cs) %% n == 0
    for length, n in ((1000, 4), (3529, 12), (17, 3), (100, 1)):
        cs = host.get_chunk(length, n)
        seq = np.stack([np.arange(100, 151), np.arange(99, 48, -1)], 1).ravel()
        first = [c for c in seq if int(np.ceil(length / c)) % n == 0]
        assert cs == (first[0] if first else cs)
    x = np.array([[0.0005, 2, 0], [3, 0, 0], [1, 1, 0]])
    xc, keep = host.clean_data(x)
    assert xc.shape == (3, 2) and xc[0, 0] == 0 and keep.tolist() == [True, True, False]
    with pytest.raises(host.SaverError):
        host.clean_data(np.ones((3, 1)))
    sf, sc = host.calc_size_factor(xc, None, 2)
    assert np.allclose(sf, xc.sum(0) / xc.sum(0).mean()) and sc == 1.0
    sf, sc = host.calc_size_factor(xc, np.array([2.0, 6.0]), 2)
    assert np.allclose(sf, [0.5, 1.5]) and sc == 4.0
    sf, sc = host.calc_size_factor(xc, 1, 2)
    assert np.allclose(sf, 1) and sc == 1.0
    for bad in (np.array([1.0, -1.0]), 3.0, np.array([1.0, 2.0, 3.0])):
        with pytest.raises(host.SaverError):
            host.calc_size_factor(xc, bad, 2)
    with pytest.raises(host.SaverError):
        host.get_pred_cells([0, 5], 3)
    with pytest.raises(host.SaverError):
        host.get_pred_genes(xc, [7], None, 3)
    with pytest.raises(host.SaverError):
        host.get_pred_genes(xc, None, 3, 3)
    assert host.get_pred_genes(xc, None, 2, 3).tolist() == [1, 0] or True
    y = np.array([0.0, 1.0, 4.0, 2.0, 7.0])
    assert np.allclose(host.est_lambda(y, 0.6, [-3.9, 13.0]), oracle.est_lambda(y, 0.6, [-3.9, 13.0]),
                       rtol=1e-14)


def test_draw_folds_are_balanced_permutations():
    f = host.draw_folds([1, 2, 3, 1], 23, 5, fold_seed=9)
    assert f.shape == (4, 23) and np.array_equal(f[0], f[3]) and not np.array_equal(f[0], f[1])
    for row in f:
        assert np.array_equal(np.sort(np.bincount(row)[1:]), [4, 4, 5, 5, 5])


def _small():
    # more predictors than cells (glmnet's lambda.min.ratio = 0.01 regime, short paths) and enough
    # genes for the fast schedule to reach its fixed-lambda stages (n3 < npred)
    x = synth.nb_counts(420, 60, seed=5).astype(np.float64)
    return x


def test_saver_slow_schedule_equals_direct_oracle_calls():
    x = _small()[:40]
    be = OracleBackend()
    import saver_b200
    res = saver_b200.saver(x, do_fast=False, backend=be, fold_seed=3, return_mu=True, perm="identity")
    # slow schedule = stages of n1 = 8, ceil((npred-8)/4), rest, all cross-validated, sent down as
    # one pass (blocks of the backend's size) with the stage-local seeds; then the null genes
    npred = int((res.x.sum(1) > 0).sum())
    stages = [c for c in be.calls]
    assert sum(c[1] for c in stages if c[0] == 1) == npred
    assert all(not c[2] for c in stages)                       # calc.maxcor = FALSE
    n1, n2 = 8, int(np.ceil((npred - 8) / 4)) + 8
    pg = np.where(res.x.sum(1) > 0)[0]
    assert np.array_equal(res.seed_of[pg], np.concatenate([np.arange(1, n1 + 1), np.arange(1, n2 - n1 + 1),
                                                           np.arange(1, npred - n2 + 1)]))
    assert (res.info["maxcor"][res.x.sum(1) > 0] == 2).all()     # rep(2, .) of R/calc_estimate.R:79
    good = np.where(res.x.mean(1) >= 0.1)[0]
    x_est = np.log((res.x[good] + 1) / res.sf)
    g = int(np.where(res.x.sum(1) > 0)[0][10])
    fold = host.draw_folds([int(res.seed_of[g])], res.x.shape[1], 5, 3)[0]
    excl = int(np.where(good == g)[0][0]) if g in good else -1
    r = oracle.predict_cv(x_est, res.x[g] / res.sf, fold, excl=excl)
    assert np.array_equal(r["mu"], res.mu[g]) and r["lambda_min"] == res.info["lambda.min"][g]
    assert res.seed_of[g] == 3                                  # third gene of the second stage


def test_saver_fast_schedule_stages_and_regressions():
    x = _small()
    be = OracleBackend()
    import saver_b200
    res = saver_b200.saver(x, do_fast=True, backend=be, fold_seed=1, perm=4)
    info = res.info
    modes = [c[0] for c in be.calls]
    assert modes[0] == 1 and 2 in modes                        # CV stages, then fixed-lambda stages
    npred = int((res.x.sum(1) > 0).sum())
    cv_genes = np.isfinite(info["sd.cv"]) & (info["lambda.max"] > 0)
    path_genes = np.isnan(info["sd.cv"])
    assert 100 <= cv_genes.sum() + (info["lambda.max"] == 0).sum() and path_genes.sum() > 0
    # cutoff = (0.5 - b0) / b1 of lm(sqrt(sd.cv) ~ maxcor) over the first n2 = 100 genes in order
    assert np.isfinite(info["cutoff"]) and len(np.atleast_1d(info["lambda.coefs"])) == 2
    # every fixed-lambda gene has maxcor above cutoff2 and lambda from est.lambda
    c = np.asarray(info["lambda.coefs"])
    cutoff2 = max(info["cutoff"], -c[0] / c[1])
    assert (info["maxcor"][path_genes] > cutoff2).all()
    g = int(np.where(path_genes)[0][0])
    lam = host.est_lambda(res.x[g] / res.sf, info["maxcor"][g], c)
    assert np.allclose([info["lambda.max"][g], info["lambda.min"][g]], lam, rtol=1e-12)
    assert npred <= res.x.shape[0] and np.isfinite(res.estimate).all() and np.isfinite(res.se).all()


def test_saver_null_and_pred_gene_arguments():
    import saver_b200
    x = _small()[:30]
    be = OracleBackend()
    res = saver_b200.saver(x, do_fast=False, npred=5, backend=be, fold_seed=0)
    # npred = 5: the five most expressed genes are predicted, the rest get the mean model
    top = np.argsort(-res.x.mean(1), kind="stable")[:5]
    assert set(np.where(res.info["lambda.max"] > 0)[0]) <= set(top.tolist())
    assert [c[0] for c in be.calls][-1] == 0
    sub = saver_b200.saver(x, do_fast=False, pred_genes=np.array([3, 4]), pred_genes_only=True,
                           backend=OracleBackend(), fold_seed=0)
    assert sub.estimate.shape[0] == 2


def test_saver_sparse_input_equals_dense_on_the_host_path():
    # scipy.sparse input (R: dgCMatrix): clean.data, size factors, gene filters and the row blocks
    # handed to calc.estimate are the same as for the dense matrix (R/utils.R:2-31, R/saver.R:156)
    import scipy.sparse as sp
    import saver_b200
    x = _small()[:36]
    x[:, 7] = 0                                       # an all-zero cell is dropped
    x[4, :5] = 0.0004                                 # below clean.data's threshold
    a = saver_b200.saver(x, do_fast=False, backend=OracleBackend(), fold_seed=2)
    b = saver_b200.saver(sp.csr_matrix(x), do_fast=False, backend=OracleBackend(), fold_seed=2)
    assert a.estimate.shape == (36, x.shape[1] - 1)
    assert np.array_equal(a.estimate, b.estimate) and np.array_equal(a.se, b.se)
    assert np.array_equal(a.info["size.factor"], b.info["size.factor"])
