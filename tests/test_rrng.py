"""R's random stream restated (saver_b200/rrng.py): known answers of set.seed() + runif() /
sample() that any R session reproduces, and the host functions built on them."""
import numpy as np
import pytest

from saver_b200 import host, rrng


def test_set_seed_runif_known_answers():
    # set.seed(1); runif(3)   /   set.seed(42); runif(2)   /   set.seed(123); runif(3)
    r = rrng.RRng(1)
    assert np.allclose([r.unif_rand() for _ in range(3)], [0.2655087, 0.3721239, 0.5728534], atol=5e-8)
    r = rrng.RRng(42)
    assert np.allclose(r.unif_rand(2), [0.9148060, 0.9370754], atol=5e-8)
    r = rrng.RRng(123)
    assert np.allclose(r.unif_rand(3), [0.2875775, 0.7883051, 0.4089769], atol=5e-8)


def test_sample_known_answers_both_samplers():
    # R >= 3.6 (sample.kind = "Rejection")
    assert rrng.RRng(1).sample_int(10).tolist() == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]
    assert rrng.RRng(42).sample_int(10).tolist() == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]
    assert rrng.RRng(123).sample_int(10).tolist() == [3, 10, 2, 8, 6, 9, 1, 7, 5, 4]
    # R < 3.6 (sample.kind = "Rounding"): the classic set.seed(1); sample(10)
    assert rrng.RRng(1, "rounding").sample_int(10).tolist() == [3, 4, 5, 7, 2, 8, 9, 6, 10, 1]
    with pytest.raises(ValueError):
        rrng.RRng(1, "walker")


def test_cv_folds_are_a_balanced_permutation():
    for kind in ("rejection", "rounding"):
        f = rrng.cv_folds(7, 203, 5, kind)
        assert f.dtype == np.int32 and sorted(np.bincount(f)[1:].tolist()) == [40, 40, 41, 41, 41]
        assert np.array_equal(f, rrng.cv_folds(7, 203, 5, kind)) and not np.array_equal(f, rrng.cv_folds(8, 203, 5, kind))
    # sample(rep(seq(5), length = 10)) after set.seed(1) in R >= 3.6: x[c(9,4,7,1,2,5,3,10,6,8)]
    base = np.arange(10) % 5 + 1
    assert rrng.cv_folds(1, 10, 5).tolist() == base[np.array([9, 4, 7, 1, 2, 5, 3, 10, 6, 8]) - 1].tolist()


def test_host_draws_reference_folds_and_gene_order():
    f = host.draw_folds([1, 2, 3], 50, 5, method="R")
    assert f.shape == (3, 50) and np.array_equal(f[1], rrng.cv_folds(2, 50, 5, "rejection"))
    assert np.array_equal(host.draw_folds([2], 50, 5, method="R-3.5")[0], rrng.cv_folds(2, 50, 5, "rounding"))
    with pytest.raises(host.SaverError):
        host.draw_folds([1], 50, method="mt")
    # every gene predictable: ind = sample(1:ngenes, ngenes) after set.seed(1)
    ind = host.gene_order(("R", 1), np.arange(10), np.zeros(0, np.int64), 10)
    assert (ind + 1).tolist() == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]
    # some genes not predictable: two consecutive sample() calls on one stream
    pred1, rest = np.array([0, 2, 3, 5, 6, 8]), np.array([1, 4, 7, 9])
    ind = host.gene_order(("R", 1), pred1, rest, 10)
    r = rrng.RRng(1)
    assert np.array_equal(ind, np.concatenate([r.sample(pred1), r.sample(rest)]))
    assert sorted(ind[:6].tolist()) == pred1.tolist() and sorted(ind[6:].tolist()) == rest.tolist()
    assert host.gene_order("identity", pred1, rest, 10).tolist() == [0, 2, 3, 5, 6, 8, 1, 4, 7, 9]
