"""GPU parity: batched Poisson-lasso paths (expr.predict, both variants) vs the CPU oracle's
glmnet restatement, on the reference's bundled data and on synthetic blocks.

Two round kernels solve the same problems (saver_cuda_set_option "lasso_gram"):
  0  the n-space kernel follows fishnet1's iterate sequence (same sweep order, same thresholds),
     so it agrees with the oracle to rounding on most genes;
  1  the Gram-space kernel (default, faster) visits coordinates in another order and reuses the
     Gram block inside a lambda step: same optimum, different iterates, so it differs from the
     oracle by glmnet's own convergence error (thresh = 1e-7 -> ~1e-4 in mu) -- still far inside
     the end-to-end tolerance of BASELINE.md (median relative error <= 1e-3, Spearman >= 0.999).
"""
import os

import numpy as np
import pytest

import oracle
from saver_b200 import ops, synth

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def _folds(B, n, nfolds, seed):
    return synth.fold_ids(B, n, nfolds, seed)


@pytest.fixture(params=[0, 1], ids=["nspace", "gram"])
def solver(request, handle):
    handle.set_option("lasso_gram", request.param)
    yield request.param
    handle.set_option("lasso_gram", 1)


def test_fixed_lambda_path_matches_oracle_and_golden(golden, handle, solver):
    """T5 on the GPU: est.lambda -> glmnet(lambda = seq) -> mu, then calc.post vs golden digits."""
    ops.set_design(golden["xest"], handle=handle)
    sel = golden["fixed_genes"][::8]
    got = ops.predict_path(golden["y"][sel], golden["lambda_max"][sel], golden["lambda_min"][sel],
                           self_col=sel, want_stats=True, handle=handle)
    ref = oracle.predict_path_batch(golden["xest"], golden["y"][sel], golden["lambda_max"][sel],
                                    golden["lambda_min"][sel], self_col=sel)
    assert (got["status"] == 0).all() and (ref["rc"] == 0).all()
    assert (got["stats"][:, 0] == ref["lmu"]).all()
    # Same iterate sequence => most genes agree to rounding; where a `dlx < thresh` decision
    # lands on the other side (fp summation order) the two stop one sweep apart, i.e. they differ
    # by glmnet's own convergence error (thresh = 1e-7 -> up to ~1e-3 in mu).
    rel = _rel(got["mu"], ref["mu"]).max(1)
    if solver == 0:
        assert np.median(rel) < 1e-9 and (rel < 1e-8).mean() > 0.6 and rel.max() < 5e-3, \
            (np.median(rel), (rel < 1e-8).mean(), rel.max())
    else:
        assert np.median(rel) < 1e-3 and rel.max() < 1e-2, (np.median(rel), rel.max())
        assert np.median(_rel(got["mu"], ref["mu"])) < 2e-4
    pr = ops.calc_post(golden["x"][sel], got["mu"], golden["sf"], 1.0, handle=handle)
    info = pr["info"]
    fb = np.array([(int(i[5]) >> (1 if i[0] == 3 else int(i[0]))) & 1 for i in info], dtype=bool)
    exact = np.rint(pr["est"] * 1000) == golden["estimate_m"][sel]
    # lambda_0 == lambda_max is a structural tie: entry of the first variable is a coin flip
    assert exact[~fb].mean() > (0.95 if solver == 0 else 0.7)


def test_cv_matches_oracle_on_golden_genes(golden, handle, solver):
    """cv.glmnet with identical folds: same lambda grid, same cvm curve, same lambda.min, same mu."""
    ops.set_design(golden["xest"], handle=handle)
    sel = golden["cv_genes"][::6]
    n = golden["x"].shape[1]
    fold = _folds(sel.size, n, 5, seed=3)
    got = ops.predict_cv(golden["y"][sel], fold, self_col=sel, want_stats=True, want_cv=True,
                         handle=handle)
    ref = oracle.predict_cv_batch(golden["xest"], golden["y"][sel], fold, self_col=sel)
    assert (got["status"] == 0).all() and (ref["rc"] == 0).all()
    assert _rel(got["lambda_max"], golden["lambda_max"][sel]).max() < 1e-12
    assert (got["stats"][:, 0] == ref["lmu"]).all(), "master path length"
    same = got["idx"].astype(int) == ref["idx"]
    assert same.mean() >= 0.9, "lambda.min index"
    assert _rel(got["lambda_min"][same], ref["lambda_min"][same]).max() < 1e-12
    rel = _rel(got["mu"][same], ref["mu"][same]).max(1)
    d = np.abs(got["sd_cv"][same] - ref["sd_cv"][same])
    if solver == 0:
        assert np.median(rel) < 1e-6 and rel.max() < 5e-3, (np.median(rel), rel.max())
        assert np.median(d) < 1e-6 and d.max() < 1e-2 * max(1.0, np.abs(ref["sd_cv"][same]).max())
    else:
        assert np.median(rel) < 1e-3 and rel.max() < 1e-2, (np.median(rel), rel.max())
        assert np.median(d) < 1e-3 * max(1.0, np.median(np.abs(ref["sd_cv"][same])))
    # acceptance #2 over all genes, flips included
    allrel = _rel(got["mu"], ref["mu"])
    assert np.median(allrel) < 1e-3


def test_cv_synthetic_block_with_pred_cells_subset(handle):
    """pred.cells subset (fold id 0 = not fitted), no self column, n and p off the tile sizes."""
    x = synth.nb_counts(150, 333, seed=21).astype(np.float64)
    sf = x.sum(0) / x.sum(0).mean()
    keep = x.mean(1) >= 0.1
    xest = np.ascontiguousarray(np.log((x[keep] + 1.0) / sf))
    Y = (x / sf)[:40]
    n = x.shape[1]
    rng = np.random.default_rng(0)
    pred = np.sort(rng.choice(n, 280, replace=False))
    fold = np.zeros((Y.shape[0], n), dtype=np.int32)
    fold[:, pred] = _folds(Y.shape[0], pred.size, 5, seed=9)
    keep_idx = np.where(keep)[0]
    self_col = np.array([int(np.where(keep_idx == g)[0][0]) if keep[g] else -1 for g in range(40)],
                        dtype=np.int32)
    ops.set_design(xest, handle=handle)
    got = ops.predict_cv(Y, fold, self_col=self_col, want_stats=True, handle=handle)
    for g in range(0, 40, 3):
        ref = oracle.predict_cv(xest, Y[g], fold[g, pred], rows=pred, excl=int(self_col[g]))
        assert got["status"][g] == ref["rc"]
        if ref["rc"] == 0 and int(got["idx"][g]) == ref["idx"]:
            # the oracle itself moves by up to 2e-2 (max over cells; median up to 2e-3) on these genes
            # between thresh = 1e-7 and 1e-11: glmnet's stopping rule, not the solver, sets the scale
            r = _rel(got["mu"][g], ref["mu"])
            assert r.max() < 3e-2 and np.median(r) < 3e-3, (g, r.max(), np.median(r))
            # the deviance-change stopping rule of the master path is razor thin: a few lambdas of slack
            assert abs(int(got["stats"][g, 0]) - ref["lmu"]) <= 5


def test_predict_edge_cases(handle):
    """sd(y) == 0 shortcut, an all-zero training fold (glmnet error -> mean), constant column."""
    rng = np.random.default_rng(4)
    n, p = 120, 60
    xest = rng.standard_normal((p, n))
    xest[7] = 1.5                                  # constant predictor is never eligible
    Y = np.exp(0.5 * xest[:4] + 0.1 * rng.standard_normal((4, n)))
    Y[1] = 2.0                                     # sd(y) == 0
    Y[2] = 0.0
    Y[2, :3] = 5.0                                 # non-zero only in three cells ...
    fold = _folds(4, n, 5, seed=1)
    fold[2, :3] = 1                                # ... all held out together: the fit that
    fold[2, 3:] = np.arange(n - 3) % 4 + 2         # leaves fold 1 out sees only zeros (jerr 9999)
    ops.set_design(xest, handle=handle)
    got = ops.predict_cv(Y, fold, handle=handle)
    assert got["status"][1] == 1 and np.allclose(got["mu"][1], 2.0)
    for g in (0, 2, 3):
        ref = oracle.predict_cv(xest, Y[g], fold[g])
        assert got["status"][g] == ref["rc"], g
        if ref["rc"] == 0 and int(got["idx"][g]) == ref["idx"]:
            # the oracle itself moves by up to 2e-2 (max over cells; median up to 2e-3) on these genes
            # between thresh = 1e-7 and 1e-11: glmnet's stopping rule, not the solver, sets the scale
            r = _rel(got["mu"][g], ref["mu"])
            assert r.max() < 3e-2 and np.median(r) < 3e-3, (g, r.max(), np.median(r))
        if ref["rc"] != 0:
            assert np.allclose(got["mu"][g], ref["mu"])


@pytest.mark.skipif(os.environ.get("SAVER_TEST_SLICES") != "1",
                    reason="time slices are opt-in: the 15,000 x n default faulted on the device in "
                           "the auto-lambda cv path (DESIGN.md); set SAVER_TEST_SLICES=1 to run")
def test_time_slices_resume_to_the_same_solution(handle, monkeypatch):
    """A round-kernel CTA gives up the SM after its time slice and resumes next round from a
    materialised n-space state.  With slices so short that every problem is cut many times the
    solver must arrive at the same path (same lambda.min, mu within the convergence error)."""
    G, n = 260, 400
    x = synth.nb_counts(G, n, seed=11).astype(np.float64)
    sf = x.sum(0) / x.sum(0).mean()
    good = np.where(x.mean(1) >= 0.1)[0]
    ops.set_design(np.log((x[good] + 1.0) / sf), handle=handle)
    sel = good[:48]
    self_col = np.arange(48, dtype=np.int32)
    fold = _folds(48, n, 5, 3)
    ref = ops.predict_cv(x[sel] / sf, fold, self_col=self_col, want_stats=True, handle=handle)
    monkeypatch.setenv("SAVER_SLICE_CYCLES", "400000")
    cut = ops.predict_cv(x[sel] / sf, fold, self_col=self_col, want_stats=True, handle=handle)
    assert (cut["stats"][:, 3] > ref["stats"][:, 3]).all()        # more rounds: the slices did cut
    assert (cut["status"] == ref["status"]).all()
    assert np.mean(cut["idx"] == ref["idx"]) >= 0.9
    same = cut["idx"] == ref["idx"]
    assert np.median(_rel(cut["mu"][same], ref["mu"][same])) < 1e-4


def test_fixed_lambda_path_converges_on_coarse_lambda_steps(handle):
    """The fixed-lambda variant steps lambda by e^-0.2 (R/expr_predict.R:62-63): on these six genes of
    a 1,200 x 2,000 synthetic matrix (tests/golden/fastpath_heavy.npz, found by
    scripts/probe_fastpath.py) the Gram kernel's kept curvature model used to oscillate up to maxit
    (glmnet jerr = -lambda index, a truncated path) where fishnet1 needs a few hundred sweeps."""
    from saver_b200 import host
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "fastpath_heavy.npz"))
    G, n = int(d["G"]), int(d["n"])
    x = synth.nb_counts(G, n, 20264).astype(np.float64)
    xc, _ = host.clean_data(x)
    sf, _ = host.calc_size_factor(xc, None, n)
    good = np.where(xc.mean(1) >= 0.1)[0]
    x_est = np.log((xc[good] + 1.0) / sf)
    self_col = np.full(G, -1, dtype=np.int32)
    self_col[good] = np.arange(good.size, dtype=np.int32)
    genes = d["genes"]
    y = xc[genes] / sf
    handle.set_option("lasso_gram", 1)
    ops.set_design(x_est, handle=handle)
    got = ops.predict_path(y, d["lmax"], d["lmin"], self_col=self_col[genes], want_stats=True, handle=handle)
    ref = oracle.predict_path_batch(x_est, y, d["lmax"], d["lmin"], self_col=self_col[genes])
    st = np.asarray(got["stats"])
    assert (got["status"] == 0).all() and (st[:, 1] == 0).all(), st
    assert (st[:, 0] == ref["lmu"]).all()
    assert (st[:, 2] < 3 * np.maximum(ref["nlp"], 300)).all(), (st[:, 2], ref["nlp"])
    # on these genes glmnet's own stopping rule leaves ~1-2e-3 (median, relative, in mu) between the
    # thresh = 1e-7 solution and the optimum (oracle at 1e-7 vs 1e-11); the Gram kernel stops by the
    # same rule on other iterates, so it is compared with the optimum on that scale
    tight = oracle.predict_path_batch(x_est, y, d["lmax"], d["lmin"], self_col=self_col[genes], thresh=1e-11)
    own = np.median(_rel(ref["mu"], tight["mu"]), 1)
    err = np.median(_rel(got["mu"], tight["mu"]), 1)
    assert (own < 5e-3).all() and (err < 1e-2).all(), (own, err)
