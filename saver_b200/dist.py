"""One process per GPU: the torch.distributed plumbing of the gene-sharded path.

Genes are independent given the design (R/calc_estimate.R:69-74 maps row chunks over workers), so
the data path has no collective: every rank runs its share of each stage's genes.  The only
exchanges are (1) the design, broadcast once (NCCL over NVLink on a GPU box), (2) the per-gene
scalars of a finished stage, needed by every rank for the schedule's two regressions
(R/saver_fit.R:158,244), (3) the final gather of estimate / se rows for the caller.
"""
import os

import numpy as np


class Comm:
    def __init__(self, backend=None, init=True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        if init and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if backend == "nccl":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group(backend=backend)
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.backend = dist.get_backend()
        self.device = (torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
                       if self.backend == "nccl" else torch.device("cpu"))
        self.bytes_broadcast = 0   # payload bytes this rank sent / received in broadcasts (NVLink)
        self.bytes_gathered = 0    # payload bytes of all_gather results

    def _t(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def barrier(self):
        self.dist.barrier()

    def allgather_ragged(self, arr):
        """arr: (k, m_r) float64 with rank-dependent m_r -> list of the world's arrays."""
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        k = arr.shape[0]
        sizes = self.torch.zeros(self.world, dtype=self.torch.int64, device=self.device)
        sizes[self.rank] = arr.shape[1]
        self.dist.all_reduce(sizes)
        sizes = sizes.cpu().numpy()
        mx = int(sizes.max())
        pad = np.zeros((k, mx))
        pad[:, :arr.shape[1]] = arr
        outs = [self.torch.empty((k, mx), dtype=self.torch.float64, device=self.device)
                for _ in range(self.world)]
        self.dist.all_gather(outs, self._t(pad))
        self.bytes_gathered += k * mx * 8 * self.world
        return [o.cpu().numpy()[:, :int(sizes[r])] for r, o in enumerate(outs)]

    def bcast_obj(self, obj, src=0):
        """Small picklable object (the plan of a root-only saver() call) from src to every rank."""
        box = [obj if self.rank == src else None]
        self.dist.broadcast_object_list(box, src=src, device=self.device)
        return box[0]

    def broadcast_array(self, arr, shape=None, src=0, keep_on_device=False):
        """Broadcast a float64 array from src; other ranks pass arr=None.  A CUDA tensor on src is
        sent as it is (no host round trip)."""
        if self.rank == src:
            if isinstance(arr, np.ndarray) or not hasattr(arr, "is_cuda"):
                t = self._t(np.asarray(arr, dtype=np.float64))
            else:
                t = arr.to(self.device).contiguous()
            shp = self.torch.tensor(list(t.shape) + [0] * (4 - t.dim()), dtype=self.torch.int64,
                                    device=self.device)
        else:
            shp = self.torch.zeros(4, dtype=self.torch.int64, device=self.device)
        self.dist.broadcast(shp, src)
        dims = [int(v) for v in shp.cpu().numpy() if v > 0]
        if self.rank != src:
            t = self.torch.empty(dims, dtype=self.torch.float64, device=self.device)
        self.dist.broadcast(t, src)
        self.bytes_broadcast += t.numel() * t.element_size()
        return t if keep_on_device else t.cpu().numpy()

    def allgather_blocks(self, t, sizes):
        """Row blocks of one matrix, block r ((sizes[r], k) float64 tensor on self.device) held by rank
        r -> the stacked matrix plus `pad` zero rows on every rank, as ONE all-gather into a tensor
        (equal blocks) or a list all-gather of padded blocks.  Stays on the device."""
        torch = self.torch
        sizes = [int(v) for v in sizes]
        k = t.shape[1] if t.dim() == 2 else 1
        tot, mx = sum(sizes), max(sizes)
        if min(sizes) == mx:
            out = torch.empty((tot,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
            self.dist.all_gather_into_tensor(out, t.contiguous())
        else:
            padded = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
            padded[:t.shape[0]] = t
            parts = [torch.empty_like(padded) for _ in range(self.world)]
            self.dist.all_gather(parts, padded)
            out = torch.cat([p[:sizes[r]] for r, p in enumerate(parts)], 0)
        self.bytes_gathered += tot * k * t.element_size()
        return out

    def max_scalar(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_rows(self, mat, owner):
        """mat: (G, n) with only this rank's rows filled; owner[g] = owning rank.  Returns the
        merged matrix on every rank (row blocks exchanged with all_gather)."""
        owner = np.asarray(owner)
        mine = np.where(owner == self.rank)[0]
        parts = self.allgather_ragged(mat[mine].T.copy() if mine.size else np.zeros((mat.shape[1], 0)))
        for r, part in enumerate(parts):
            rows = np.where(owner == r)[0]
            if rows.size:
                mat[rows] = part.T
        return mat


def shard(count, rank, world):
    """Positions of a stage's genes owned by `rank` (round-robin deal)."""
    return np.arange(count)[rank::world]
