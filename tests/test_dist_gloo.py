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
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), est=res.estimate, se=res.se,
             lmin=res.info["lambda.min"], status=res.status)
    comm.barrier()


def test_two_rank_sharding_is_result_invariant(tmp_path):
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import saver_b200
    from helpers import OracleBackend
    from saver_b200 import synth
    x = synth.nb_counts(90, 50, seed=8).astype(np.float64)
    one = saver_b200.saver(x, do_fast=False, backend=OracleBackend(), fold_seed=2)
    for r in range(2):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(d["est"], one.estimate) and np.array_equal(d["se"], one.se)
        assert np.array_equal(d["lmin"], one.info["lambda.min"])
