"""GPU: sample.saver (distributional parity + determinism) and cor.genes / cor.cells."""
import numpy as np
import pytest
from scipy import stats

import saver_b200
from oracle import posthoc
from saver_b200 import ops

pytestmark = pytest.mark.gpu


class _Obj:
    def __init__(self, est, se, sf):
        self.estimate, self.se, self.info = est, se, {"size.factor": sf}


def _golden_obj(golden):
    return _Obj(golden["estimate_m"] / 1000.0, golden["se_m"] / 1000.0, golden["sf"])


def test_cor_genes_and_cells_match_numpy(golden, handle):
    """T12: cor(t(estimate)) * outer(adj, adj) on the reference's stored object
    (man/cor_adjust.Rd:31-36 runs exactly this); the diagonal is adj^2, not 1."""
    obj = _golden_obj(golden)
    est, se = obj.estimate[:700], obj.se[:700]
    o = _Obj(est, se, None)
    got = saver_b200.cor_genes(o, handle=handle)
    ref = posthoc.cor_adjust(est, se)
    assert got.shape == (700, 700) and np.abs(got - ref).max() < 1e-10
    assert np.abs(np.diag(got) - np.diag(ref)).max() < 1e-12 and (np.diag(got) < 1).all()
    gc = saver_b200.cor_cells(o, handle=handle)
    assert gc.shape == (200, 200) and np.abs(gc - posthoc.cor_adjust(est, se, by_cells=True)).max() < 1e-10
    # column block (what one rank of a multi-GPU job computes) == the same columns of the whole
    blk = ops.cor_adjust(est, se, col0=130, ncols=77, handle=handle)
    assert np.array_equal(blk, got[130:207])
    # user-supplied cor.mat is only rescaled
    cm = np.corrcoef(est)
    assert np.abs(saver_b200.cor_genes(o, cor_mat=cm, handle=handle) - ref).max() < 1e-12


def test_sample_saver_distribution_and_determinism(golden, handle):
    """T11: Gamma(est^2/se^2, rate est/se^2) rounded to 3 dp; NB(mu = est, size = est^2/se^2)."""
    obj = _golden_obj(golden)
    sub = _Obj(obj.estimate[:64], obj.se[:64], None)
    a = saver_b200.sample_saver(sub, seed=50, handle=handle)
    b = saver_b200.sample_saver(sub, seed=50, handle=handle)
    c = saver_b200.sample_saver(sub, seed=51, handle=handle)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.abs(a * 1000 - np.rint(a * 1000)).max() < 1e-6 and (a >= 0).all()
    # blocking invariance: the draw depends on (seed, gene, cell) only
    blk = saver_b200.sample_saver(sub, seed=50, handle=handle, block=7)
    assert np.array_equal(a, blk)
    reps = saver_b200.sample_saver(sub, rep=3, seed=1, handle=handle)
    assert len(reps) == 3 and not np.array_equal(reps[0], reps[1])
    # distribution: many draws of a few (shape, rate) pairs, KS against scipy
    rng = np.random.default_rng(0)
    # (scales >> 1e-3 so the 3-decimal rounding of R/sample_saver.R:57 is invisible to the test)
    for shape, rate in ((0.6, 0.002), (1.0, 0.01), (4.5, 0.03), (120.0, 0.1)):
        est = np.full((1, 40000), shape / rate)
        se = np.full((1, 40000), np.sqrt(shape) / rate)
        d = ops.sample(est, se, seed=int(rng.integers(1 << 30)), handle=handle)[0]
        assert stats.kstest(d, "gamma", args=(shape, 0, 1.0 / rate)).pvalue > 1e-3
        k = ops.sample(est, se, nb_mode=True, seed=7, handle=handle)[0]
        mu, size = shape / rate, shape
        assert abs(k.mean() - mu) < 5 * np.sqrt((mu + mu * mu / size) / k.size)
        assert abs(k.var() - (mu + mu * mu / size)) < 0.08 * (mu + mu * mu / size)
        assert (k == np.rint(k)).all() and (k >= 0).all()
    # all-zero genes (estimate = se = 0): NaN, as rgamma(n, NaN, NaN) in R
    z = ops.sample(np.zeros((1, 8)), np.zeros((1, 8)), seed=1, handle=handle)
    assert np.isnan(z).all()
    with pytest.raises(saver_b200.SaverError):
        saver_b200.sample_saver(sub, rep=0)


def test_get_mu_and_combine_saver(golden):
    obj = _golden_obj(golden)
    x = golden["counts"].astype(np.float64)
    mu = saver_b200.get_mu(x, obj)
    ref = posthoc.get_mu(x, obj.estimate, obj.se, golden["sf"])
    ok = np.isfinite(ref)
    assert np.array_equal(mu[ok], ref[ok])
    info = {k: np.zeros(5) for k in ("maxcor", "lambda.max", "lambda.min", "sd.cv", "pred.time", "var.time")}
    info.update({"size.factor": golden["sf"], "cutoff": 0.3, "lambda.coefs": np.zeros(2), "total.time": 1.5})
    a = saver_b200.SaverResult(obj.estimate[:5], obj.se[:5], info)
    b = saver_b200.SaverResult(obj.estimate[5:10], obj.se[5:10], info)
    c = saver_b200.combine_saver([a, b])
    assert c.estimate.shape == (10, 200) and c.info["total.time"] == 3.0 and c.info["maxcor"].size == 10


def test_legacy_alpha_beta_objects_use_the_same_kernels(handle):
    """sample.saver.old / cor.genes.old / cor.cells.old (R/utils.R:137-214): a legacy object
    list(estimate, alpha, beta) must give what the current object with se = sqrt(alpha)/beta gives."""
    rng = np.random.default_rng(4)
    alpha = rng.gamma(3.0, 1.0, (40, 60)) + 0.05
    beta = rng.gamma(2.0, 0.5, (40, 60)) + 0.05
    leg = saver_b200.LegacySaverResult(alpha / beta, alpha, beta, {})
    cur = _Obj(alpha / beta, np.sqrt(alpha) / beta, None)
    assert np.array_equal(saver_b200.sample_saver(leg, seed=3, handle=handle),
                          saver_b200.sample_saver(cur, seed=3, handle=handle))
    assert np.array_equal(saver_b200.cor_genes(leg, handle=handle), saver_b200.cor_genes(cur, handle=handle))
    assert np.array_equal(saver_b200.cor_cells(leg, handle=handle), saver_b200.cor_cells(cur, handle=handle))
    ref = posthoc.cor_adjust(alpha / beta, np.sqrt(alpha) / beta)
    assert np.abs(saver_b200.cor_genes(leg, handle=handle) - ref).max() < 1e-10
    # NB mode: rnbinom(mu = estimate, size = alpha) -- mean and variance of many draws of one pair
    a1, b1 = np.full((1, 40000), 4.5), np.full((1, 40000), 0.03)
    k = saver_b200.sample_saver(saver_b200.LegacySaverResult(a1 / b1, a1, b1, {}), efficiency_known=True,
                                seed=7, handle=handle)[0]
    mu, size = 4.5 / 0.03, 4.5
    assert abs(k.mean() - mu) < 5 * np.sqrt((mu + mu * mu / size) / k.size)
    assert (k == np.rint(k)).all()


def test_cor_prep_and_blocks_tile_cor_genes(golden, handle):
    """saver_cuda_cor_prep + saver_cuda_cor_block (the gene-sharded cor.genes of SURVEY.md 8(e)(4)):
    blocks filled from a standardised panel reproduce saver_cuda_cor_adjust, upper-triangle blocks +
    their transposes give the whole symmetric matrix."""
    import torch
    obj = _golden_obj(golden)
    est, se = obj.estimate[:333], obj.se[:333]
    G, n = est.shape
    want = ops.cor_adjust(est, se, handle=handle)
    dev = torch.device("cuda", handle.device)
    Z, adj = ops.cor_prep(torch.from_numpy(est).to(dev), torch.from_numpy(se).to(dev), handle=handle)
    handle.synchronize()
    assert Z.shape == (G, (n + 15) // 16 * 16) and not Z[:, n:].any()
    assert np.allclose((Z * Z).sum(1).cpu().numpy(), 1.0, rtol=1e-12)
    Zp = torch.cat([Z, torch.zeros((128, Z.shape[1]), dtype=Z.dtype, device=dev)], 0)
    full = torch.zeros((G, G), dtype=torch.float64, device=dev)
    ed = [0, 100, 101, 260, G]
    torch.cuda.synchronize()
    for i in range(4):
        for j in range(i, 4):
            ops.cor_block(Zp, adj, G, ed[j], ed[j + 1] - ed[j], ed[i], ed[i + 1] - ed[i],
                          full[ed[i]:ed[i + 1]], handle=handle)
    handle.synchronize()
    got = full.cpu().numpy()
    got = np.triu(got) + np.triu(got, 1).T
    assert np.abs(got - want).max() < 1e-12
    with pytest.raises(ValueError):
        ops.cor_block(Z, adj, G, 0, G, 0, G, full, handle=handle)   # panel without the zero tile rows
