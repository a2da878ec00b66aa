"""world_size-2 gloo tests of the gene-sharded path's host logic (no GPU): the collectives of
saver_b200.dist and rank-count invariance of saver.fit (sharding is result-
invariant, bitwise)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import saver_b200
    from helpers import OracleBackend
    from saver_b200 import dist, synth
    comm = dist.Comm(backend="gloo")
    # collectives
    parts = comm.allgather_ragged(np.full((2, 3 + rank), float(rank)))
    assert [p.shape for p in parts] == [(2, 3), (2, 4)] and parts[1][0, 0] == 1.0
    arr = comm.broadcast_array(np.arange(12.0).reshape(3, 4) if rank == 0 else None)
    assert arr.shape == (3, 4) and arr[2, 3] == 11.0
    assert comm.max_scalar(rank + 0.5) == 1.5
    mat = np.zeros((5, 2))
    owner = np.array([0, 1, 1, 0, 1])
    mat[owner == rank] = rank + 1
    merged = comm.gather_rows(mat, owner)
    assert merged[:, 0].tolist() == [1, 2, 2, 1, 2]
    assert dist.shard(7, rank, world).tolist() == list(range(rank, 7, world))
    # the sharded schedule
    x = synth.nb_counts(90, 50, seed=8).astype(np.float64)
    res = saver_b200.saver(x, do_fast=False, backend=OracleBackend(threads=1), comm=comm, fold_seed=2)
    # root-only call: only rank 0 holds the matrix (x = None elsewhere); the design and the counts of
    # the requested genes arrive by broadcast; gather=False leaves every rank with its own rows
    pg = np.arange(3, 40, 3)
    be = OracleBackend(threads=1)
    ro = saver_b200.saver(x if rank == 0 else None, do_fast=False, pred_genes=pg, pred_genes_only=True,
                          backend=be, comm=comm, fold_seed=2, gather=False)
    assert be.xest is not None and be.xest.shape[1] == 50        # the broadcast design
    own = ro.owner == rank
    assert own.sum() in (pg.size // 2, pg.size - pg.size // 2)
    assert (ro.estimate[~own] == 0).all() and (ro.estimate[own].sum(1) > 0).any()
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), est=res.estimate, se=res.se,
             lmin=res.info["lambda.min"], status=res.status, ro_est=ro.estimate, ro_own=own,
             ro_lmin=ro.info["lambda.min"])
    comm.barrier()


class _Obj:
    def __init__(self, est, se):
        self.estimate, self.se = est, se


def _cpu_posthoc_ops(ops):
    """Stand-ins with the contract of the device ops (saver_cuda_sample: the draw depends on
    (seed, rep, global gene, cell) only; saver_cuda_cor_adjust: column blocks of one symmetric
    matrix), so the rank-sharded host logic can run without a GPU."""
    from oracle import posthoc

    def sample(est, se, nb_mode=False, seed=0, rep_idx=0, gene0=0, handle=None):
        out = np.empty_like(est)
        for g in range(est.shape[0]):
            rng = np.random.default_rng([int(seed), int(rep_idx), int(gene0) + g, int(nb_mode)])
            out[g] = np.round(rng.gamma(est[g] ** 2 / se[g] ** 2, se[g] ** 2 / est[g]), 3)
        return out

    def cor_adjust(est, se, by_cells=False, cor_in=None, col0=0, ncols=None, handle=None):
        full = posthoc.cor_adjust(est, se, by_cells=by_cells)
        ncols = full.shape[0] - col0 if ncols is None else ncols
        blk = full[col0:col0 + ncols]
        if cor_in is not None:  # a user-supplied cor.mat block is only rescaled
            adj = np.sqrt(np.diag(full))
            blk = np.asarray(cor_in) * np.outer(adj[col0:col0 + ncols], adj)
        return blk

    ops.sample, ops.cor_adjust = sample, cor_adjust


def _posthoc_inputs():
    rng = np.random.default_rng(12)
    est = rng.gamma(2.0, 1.0, (13, 9)) + 0.1
    se = rng.gamma(1.0, 0.3, (13, 9)) + 0.05
    return est, se


def _posthoc_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import saver_b200
    from saver_b200 import dist, host, ops
    _cpu_posthoc_ops(ops)
    comm = dist.Comm(backend="gloo")
    est, se = _posthoc_inputs()
    obj = _Obj(est, se)
    leg = saver_b200.LegacySaverResult(est, (est / se) ** 2, est / se ** 2, {})
    out = {
        "samp": host.sample_saver(obj, seed=50, comm=comm, block=4),
        "reps": np.stack(host.sample_saver(obj, rep=2, seed=7, comm=comm)),
        "noseed": host.sample_saver(obj, comm=comm),            # ranks must agree on one seed
        "leg": host.sample_saver(leg, seed=50, comm=comm),
        "cg": host.cor_genes(obj, comm=comm),
        "cc": host.cor_cells(obj, comm=comm),
        "cg_user": host.cor_genes(obj, cor_mat=np.corrcoef(est), comm=comm),
    }
    blk, (c0, c1) = host.cor_genes(obj, comm=comm, gather=False)
    assert blk.shape == (c1 - c0, est.shape[0]) and np.array_equal(blk, out["cg"][c0:c1])
    tiny = _Obj(est[:3], se[:3])  # fewer genes than ranks when world = 4: some blocks are empty
    out["tiny_samp"] = host.sample_saver(tiny, seed=3, comm=comm)
    out["tiny_cg"] = host.cor_genes(tiny, comm=comm)
    np.savez(os.path.join(outdir, "ph%d.npz" % rank), **out)
    comm.barrier()


@pytest.mark.parametrize("world", [2, 4])
def test_posthoc_sharding_is_result_invariant(tmp_path, world):
    """sample.saver / cor.genes / cor.cells sharded over ranks: contiguous
    gene blocks with global gene offsets, column panels of the adjusted correlation, row gather."""
    port = 29900 + (os.getpid() % 90) + world
    mp.spawn(_posthoc_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from saver_b200 import host, ops
    saved = ops.sample, ops.cor_adjust
    try:
        _cpu_posthoc_ops(ops)
        est, se = _posthoc_inputs()
        obj = _Obj(est, se)
        one = {"samp": host.sample_saver(obj, seed=50), "reps": np.stack(host.sample_saver(obj, rep=2, seed=7)),
               "leg": host.sample_saver(obj, seed=50), "cg": host.cor_genes(obj), "cc": host.cor_cells(obj),
               "cg_user": host.cor_genes(obj, cor_mat=np.corrcoef(est)),
               "tiny_samp": host.sample_saver(_Obj(est[:3], se[:3]), seed=3),
               "tiny_cg": host.cor_genes(_Obj(est[:3], se[:3]))}
    finally:
        ops.sample, ops.cor_adjust = saved
    d = [np.load(os.path.join(str(tmp_path), "ph%d.npz" % r)) for r in range(world)]
    for k, v in one.items():
        for r in range(world):
            assert (np.array_equal(d[r][k], v) if k in ("samp", "reps", "tiny_samp")
                    else np.allclose(d[r][k], v, rtol=1e-12)), k
    assert all(np.array_equal(d[0]["noseed"], d[r]["noseed"]) for r in range(world))
    assert one["cg"].shape == (13, 13) and one["cc"].shape == (9, 9)


def test_two_rank_sharding_is_result_invariant(tmp_path):
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import saver_b200
    from helpers import OracleBackend
    from saver_b200 import synth
    x = synth.nb_counts(90, 50, seed=8).astype(np.float64)
    one = saver_b200.saver(x, do_fast=False, backend=OracleBackend(), fold_seed=2)
    pg = np.arange(3, 40, 3)
    sub = saver_b200.saver(x, do_fast=False, pred_genes=pg, pred_genes_only=True, backend=OracleBackend(),
                           fold_seed=2)
    merged = np.zeros_like(sub.estimate)
    for r in range(2):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(d["est"], one.estimate) and np.array_equal(d["se"], one.se)
        assert np.array_equal(d["lmin"], one.info["lambda.min"])
        merged[d["ro_own"]] = d["ro_est"][d["ro_own"]]
        assert np.array_equal(d["ro_lmin"], sub.info["lambda.min"])   # stage scalars reach every rank
    assert np.array_equal(merged, sub.estimate)                        # the slices tile the one-rank result


def _cpu_cor_ops(ops):
    """numpy stand-ins with the contract of saver_cuda_cor_prep / saver_cuda_cor_block."""
    import torch

    def cor_prep(est, se, handle=None):
        e, s = est.numpy(), se.numpy()
        B, n = e.shape
        ldz = (n + 15) // 16 * 16
        d = e - e.mean(1, keepdims=True)
        ss = (d ** 2).sum(1)
        Z = np.zeros((B, ldz))
        Z[:, :n] = d / np.sqrt(ss)[:, None]
        var = ss / (n - 1)
        return torch.from_numpy(Z), torch.from_numpy(np.sqrt(var / (var + (s ** 2).mean(1))))

    def cor_block(Z, adj, M, row0, nrows, col0, ncols, out, handle=None):
        assert Z.shape[0] >= M + 128 and not Z[M:].any()
        z, a = Z.numpy(), adj.numpy()
        c = np.clip(z[col0:col0 + ncols] @ z[row0:row0 + nrows].T, -1.0, 1.0)
        out[:, row0:row0 + nrows] = torch.from_numpy(c * np.outer(a[col0:col0 + ncols], a[row0:row0 + nrows]))
        return out

    ops.cor_prep, ops.cor_block = cor_prep, cor_block


def _cor_sharded_worker(rank, world, port, outdir, equal):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from saver_b200 import dist, host, ops
    _cpu_cor_ops(ops)
    comm = dist.Comm(backend="gloo")
    est, se = _posthoc_inputs()
    if equal:
        est, se = est[:12], se[:12]
    rows = np.array_split(np.arange(est.shape[0]), world)[rank]
    panel, (c0, c1), blocks = host.cor_genes_sharded(est[rows], se[rows], comm)
    assert (c0, c1) == (rows[0], rows[-1] + 1) and panel.shape == (rows.size, est.shape[0])
    np.savez(os.path.join(outdir, "cs%d.npz" % rank), panel=panel.numpy(), c=np.array([c0, c1]),
             blocks=np.array(blocks))
    comm.barrier()


@pytest.mark.parametrize("world,equal", [(2, False), (3, False), (4, True)])
def test_cor_genes_sharded_tiles_the_symmetric_matrix(tmp_path, world, equal):
    """cor.genes on a gene-sharded object: all-gather of the standardised rows,
    every unordered pair of gene blocks computed by exactly one rank; the blocks and their
    transposes tile the one-rank result."""
    from oracle import posthoc
    port = 29700 + (os.getpid() % 90) + 7 * world
    mp.spawn(_cor_sharded_worker, args=(world, port, str(tmp_path), equal), nprocs=world, join=True)
    est, se = _posthoc_inputs()
    if equal:
        est, se = est[:12], se[:12]
    want = posthoc.cor_adjust(est, se)
    G = est.shape[0]
    full, seen = np.zeros((G, G)), np.zeros((G, G), dtype=int)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "cs%d.npz" % r))
        c0, c1 = d["c"]
        for r0, r1 in d["blocks"]:
            full[c0:c1, r0:r1] = d["panel"][:, r0:r1]
            seen[c0:c1, r0:r1] += 1
            if (r0, r1) != (c0, c1):
                full[r0:r1, c0:c1] = d["panel"][:, r0:r1].T
                seen[r0:r1, c0:c1] += 1
    assert (seen == 1).all()
    assert np.allclose(full, want, rtol=1e-12, atol=1e-14)


def test_cyclic_window_covers_every_block_pair_once():
    from saver_b200 import host
    for world in range(1, 10):
        pairs = [tuple(sorted((r, b))) for r in range(world) for b in host.cyclic_window(r, world)]
        assert sorted(pairs) == sorted((a, b) for a in range(world) for b in range(a, world))
        work = [len(host.cyclic_window(r, world)) for r in range(world)]
        assert max(work) - min(work) <= 1
