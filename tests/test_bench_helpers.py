"""bench.py's host-side bookkeeping (no GPU): how a step's gene block is picked, dealt over the ranks
and seeded, and that both arms describe the same workload."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_block_is_dealt_once_with_stage_local_seeds():
    for B in (5, 8, 9, 256, 512):
        layout = bench.stage_layout(B)
        assert layout[0][0] == 0 and layout[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(layout, layout[1:]))
        for world in (1, 2, 3, 8):
            seen = np.zeros(B, dtype=int)
            for rank in range(world):
                pos, seeds = bench.deal(B, rank, world)
                seen[pos] += 1
                for p, s in zip(pos, seeds):   # cv seed = 1-based position inside the gene's stage
                    a = [lo for lo, hi in layout if lo <= p < hi][0]
                    assert s == p - a + 1
            assert (seen == 1).all()


def test_pick_block_is_a_fixed_spread_of_the_predictable_genes():
    pred = np.arange(3, 2000, 3)
    blk = bench.pick_block(pred, 64)
    assert blk.size == 64 and blk[0] == pred[0] and blk[-1] == pred[-1] and (np.diff(blk) > 0).all()
    assert np.array_equal(bench.pick_block(pred, 0), pred) and np.array_equal(bench.pick_block(pred, 10 ** 6), pred)


def test_both_arms_name_the_same_config():
    w = bench.WORKLOADS["cfg3"]
    a = bench.config_dict("cfg3", w, 512, 18385, 4)
    b = bench.config_dict("cfg3", w, 512, 18385, 4)
    assert a == b and a["genes_per_step"] == 512 and a["genes_per_step_per_gpu"] == 128
    assert "20000 genes x 10000 cells" in a["workload"] and "512 genes per step" in a["workload"]
    assert w["block"] == 512 and set(bench.WORKLOADS) >= {"cfg2", "cfg3", "cfg4", "cfg5", "null", "fast"}


def test_sparse_generator_builds_a_canonical_dgcmatrix():
    """bench_extra.sparse_nb_counts assembles the CSC slots chunk by chunk (cfg4's 30k x 50k matrix
    never exists densely on the host): the result must be a canonical dgCMatrix -- column pointers
    consistent, row indices sorted inside a column, no stored zeros -- with the model's sparsity."""
    import torch

    import bench_extra
    x = bench_extra.sparse_nb_counts(300, 170, 7, torch.device("cpu"), torch, chunk=64)
    assert x.shape == (300, 170) and x.format == "csc"
    assert x.indptr[0] == 0 and x.indptr[-1] == x.nnz == x.data.size and (np.diff(x.indptr) >= 0).all()
    assert x.has_canonical_format and (x.data > 0).all() and (x.data == np.rint(x.data)).all()
    assert 0.1 < x.nnz / (300 * 170) < 0.7
    y = bench_extra.sparse_nb_counts(300, 170, 7, torch.device("cpu"), torch, chunk=64)
    assert (x != y).nnz == 0                      # seeded
