"""torchrun --nproc-per-node 2 scripts/run_dist_saver.py : saver() sharded over ranks (NCCL design
broadcast, per-stage all_gather of the scalars, final row gather) equals the one-rank result."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saver_b200
from saver_b200 import dist, host, synth
comm = dist.Comm()
h = saver_b200.Handle(int(os.environ.get("LOCAL_RANK", "0")))
x = synth.nb_counts(300, 400, seed=4)
res = saver_b200.saver(x, do_fast=True, backend=host.CudaBackend(handle=h), comm=comm, fold_seed=3)
if comm.rank == 0:
    one = saver_b200.saver(x, do_fast=True, backend=host.CudaBackend(handle=h), fold_seed=3)
    same = np.array_equal(one.estimate, res.estimate) and np.array_equal(one.se, res.se)
    print("ranks", comm.world, "backend", comm.backend, "identical to one-rank run:", same,
          "max |diff|", float(np.abs(one.estimate - res.estimate).max()), flush=True)
# the consumers of the object, sharded the same way (contiguous gene blocks / column panels)
samp = saver_b200.sample_saver(res, seed=50, handle=h, comm=comm)
cg = saver_b200.cor_genes(res, handle=h, comm=comm)
if comm.rank == 0:
    s1 = saver_b200.sample_saver(res, seed=50, handle=h)
    c1 = saver_b200.cor_genes(res, handle=h)
    print("sample.saver identical:", np.array_equal(s1, samp, equal_nan=True),
          " cor.genes identical:", np.array_equal(c1, cg, equal_nan=True), flush=True)
comm.barrier()
import torch.distributed as _d
_d.destroy_process_group()
