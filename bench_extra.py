"""bench_extra.py -- the non-headline workloads of bench.py (same JSON contract, one line, rank 0).

    python bench.py --workload null     saver(x, null.model = TRUE, estimates.only = TRUE), 20k x 10k
                                         = calc.b + Gamma posterior for every gene (variance.cu alone)
    python bench.py --workload fast     saver(x, do.fast = TRUE) on 5,000 x 2,000 (the default schedule)
    python bench.py --workload cfg5     sample.saver (gamma + NB) and cor.genes on 20k x 10k est / se

N > 1: genes (null, sample) or column panels (cor.genes) are dealt in contiguous blocks over the
ranks, no data-path collective; time = max over ranks.
"""
import json
import os
import time

import numpy as np


def _env():
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libsaver_cuda has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        import torch.distributed as tdist
        from saver_b200 import dist as sdist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", device_id=dev)
        comm = sdist.Comm(init=False)
    return torch, rank, world, local, dev, comm


def _max_over_ranks(comm, vals, dev, torch):
    if comm is None:
        return [float(v) for v in vals]
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=dev)
    import torch.distributed as tdist
    tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    return [float(v) for v in t.cpu().numpy()]


def _finish(comm):
    if comm is not None:
        import torch.distributed as tdist
        comm.barrier()
        tdist.destroy_process_group()


def _timed(fn, steps, warmup, stream, flush, torch, barrier):
    for _ in range(warmup):
        flush.zero_()
        fn()
    barrier()
    evs = []
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    return sum(a.elapsed_time(b) for a, b in evs) / 1e3


def run(args):
    if args.workload == "null":
        run_null(args)
    elif args.workload == "fast":
        run_fast(args)
    elif args.workload == "cfg4":
        run_cfg4(args)
    else:
        run_cfg5(args)


# ------------------------------------------------------------------------------------------------
def run_null(args):
    import bench
    import saver_b200
    from saver_b200 import host, ops
    torch, rank, world, local, dev, comm = _env()
    h = saver_b200.Handle(local)
    stream = torch.cuda.ExternalStream(h.stream(), device=dev)
    w = bench.WORKLOADS["null"]
    G, n = (w["G"], w["n"]) if args.block is None else (args.block, w["n"])
    from saver_b200 import synth
    x = synth.nb_counts(G, n, w["seed"])          # every rank draws the same matrix, keeps its block
    xc, _ = host.clean_data(x)
    sf, scale_sf = host.calc_size_factor(xc, None, n)
    e = [(G * r) // world for r in range(world + 1)]
    mine = np.arange(e[rank], e[rank + 1])
    Xd = torch.from_numpy(np.ascontiguousarray(xc[mine])).to(dev)
    sfd = torch.from_numpy(sf).to(dev)
    mud = (Xd / sfd).mean(1).contiguous()           # null model: mu = mean(y / sf) per gene
    est = torch.empty_like(Xd)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    blk = 4096

    def barrier():
        h.synchronize()
        torch.cuda.synchronize()
        if comm is not None:
            comm.barrier()

    def step():
        for b0 in range(0, mine.size, blk):
            sl = slice(b0, min(b0 + blk, mine.size))
            ops.calc_post(Xd[sl], mud[sl], sfd, scale_sf, want_se=False, want_info=False,
                          out={"est": est[sl]}, handle=h)

    clocks = bench.ClockSampler(local)
    l0 = h.launch_count()
    clocks.start()
    t = _timed(step, args.steps, args.warmup, stream, flush, torch, barrier)
    clk = clocks.stop()
    launches = h.launch_count() - l0
    # e2e through the public API from host buffers
    be = host.CudaBackend(handle=h)
    sub = xc[mine]
    saver_b200.saver(sub[:64], null_model=True, estimates_only=True, backend=be)
    barrier()
    t0 = time.time()
    est_h = saver_b200.saver(sub, null_model=True, estimates_only=True, backend=be)
    h.synchronize()
    e2e = time.time() - t0
    tmax, e2e_max = _max_over_ranks(comm, [t, e2e], dev, torch)
    parity = None
    if rank == 0:
        import oracle
        k = min(16, mine.size)
        sel = np.linspace(0, mine.size - 1, k).astype(int)
        t0 = time.time()
        ref = oracle.calc_post_batch(sub[sel], np.repeat((sub[sel] / sf).mean(1)[:, None], n, 1), sf, scale_sf)
        dt = time.time() - t0
        got = est[torch.from_numpy(sel).to(dev)].cpu().numpy()
        rel = np.abs(got - ref["est"]) / np.maximum(ref["est"], 1e-12)
        parity = {"genes": int(k), "max_rel_err_estimate": float(rel.max()),
                  "quantum": "values are ceiling(v * 1000) / 1000: one quantum = 1e-3 absolute",
                  "max_abs_err": float(np.abs(got - ref["est"]).max()),
                  "e2e_equal_resident": bool(np.array_equal(est_h[sel], got))}
        cpu = {"value": k / dt * (os.cpu_count() or 1) / min(k, os.cpu_count() or 1), "unit": "genes/s",
               "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "%d genes, one gene per thread, %.2f s" % (k, dt)}
        hbm, which = bench.peaks()
        bytes_step = 16.0 * n * G    # read y + write est (scalar mu, estimates.only); sf amortised
        line = {
            "metric": "genes/sec, saver(null.model = TRUE, estimates.only = TRUE), synthetic NB",
            "value": G * args.steps / tmax, "unit": "genes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tmax * 1e3 / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "null: %d genes x %d cells, calc.b (Brent + grid branch) + Gamma posterior "
                                   "for every gene" % (G, n), "l2": "flushed between steps (256 MiB write)"},
            "clocks": clk,
            "e2e": {"value": G / e2e_max, "unit": "genes/s", "h2d_bytes_per_step": int(sub.nbytes * world),
                    "d2h_bytes_per_step": int(sub.nbytes * world), "s_per_step": e2e_max,
                    "api": "saver_b200.saver(x, null_model=True, estimates_only=True)"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "calc_post_kernel", "bound": "hbm", "unit": "GB/s",
                         "achieved": bytes_step * args.steps / tmax / 1e9, "peak": hbm,
                         "frac": bytes_step * args.steps / tmax / 1e9 / hbm, "peak_source": which,
                         "traffic": None,
                         "algorithmic": "16*n B/gene (read y, write est; scalar mu, estimates.only)",
                         "note": "the kernel is FP64-transcendental bound (one log per cell and one "
                                 "log-of-product per non-zero cell per likelihood evaluation); see "
                                 "profiles/ for the FP64-pipe utilisation"},
            "parity": parity, "cpu_baseline": cpu if world == 1 else None,
        }
        print(json.dumps(line), flush=True)
    _finish(comm)


# ------------------------------------------------------------------------------------------------
def run_fast(args):
    import bench
    import saver_b200
    from saver_b200 import host
    torch, rank, world, local, dev, comm = _env()
    h = saver_b200.Handle(local)
    w = bench.WORKLOADS["fast"]
    x = bench.make_counts("fast") if rank == 0 or comm is None else None
    be = host.CudaBackend(handle=h, comm=comm)

    def barrier():
        h.synchronize()
        torch.cuda.synchronize()
        if comm is not None:
            comm.barrier()

    kw = dict(do_fast=True, estimates_only=False, fold_seed=args.fold_seed, perm=args.fold_seed,
              backend=be, comm=comm, gather=False)
    clocks = bench.ClockSampler(local)
    times = []
    l0 = h.launch_count()
    res = None
    for it in range(args.warmup + args.steps):
        barrier()
        if it == args.warmup:
            clocks.start()
            l0 = h.launch_count()
        t0 = time.time()
        res = saver_b200.saver(x, **kw)
        h.synchronize()
        if it >= args.warmup:
            times.append(time.time() - t0)
    clk = clocks.stop()
    launches = h.launch_count() - l0
    (tmax,) = _max_over_ranks(comm, [float(np.sum(times))], dev, torch)
    if rank == 0:
        info = res.info
        G = res.x.shape[0]
        ncv = int((np.isfinite(info["sd.cv"]) & (info["lambda.max"] > 0)).sum())
        nfix = int(np.isnan(info["sd.cv"]).sum())
        line = {
            "metric": "genes/sec, saver(do.fast = TRUE) (default schedule), synthetic NB",
            "value": G * args.steps / tmax, "unit": "genes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tmax * 1e3 / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "fast: %d genes x %d cells, do.fast=TRUE: %d genes cross-validated, %d "
                                   "fixed-lambda paths, the rest means; estimate + se" % (G, w["n"], ncv, nfix)},
            "clocks": clk,
            "e2e": {"value": G * args.steps / tmax, "unit": "genes/s",
                    "h2d_bytes_per_step": int(res.x.nbytes * 2), "d2h_bytes_per_step": int(res.x.nbytes * 2),
                    "s_per_step": tmax / args.steps,
                    "api": "saver_b200.saver(x, do_fast=True): value IS the end-to-end number (the stage "
                           "schedule needs the host regressions between device passes)"},
            "gpu_launches": int(launches),
            "roofline": None,
        }
        print(json.dumps(line), flush=True)
    _finish(comm)


# ------------------------------------------------------------------------------------------------
def sparse_nb_counts(G, n, seed, dev, torch, chunk=2000, k=12):
    """synth.nb_counts' model drawn on the device cell-chunk by cell-chunk and assembled straight into
    the three slots of a dgCMatrix (genes x cells, compressed by cell) -- a 30k x 50k matrix never
    exists densely on the host."""
    import scipy.sparse as sp
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    f64 = torch.float64
    W = torch.randn((G, k), generator=gen, device=dev, dtype=f64) * 0.5
    W[torch.rand(G, generator=gen, device=dev) < 0.3] = 0.0
    b = torch.randn(G, generator=gen, device=dev, dtype=f64) * 1.5 - 1.0
    phi = torch.clamp(torch._standard_gamma(torch.full((G,), 2.0, device=dev, dtype=f64), generator=gen), 0.1, 50.0)
    data, idx, counts = [], [], []
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        Z = torch.randn((c1 - c0, k), generator=gen, device=dev, dtype=f64)
        depth = torch.exp(torch.randn(c1 - c0, generator=gen, device=dev, dtype=f64) * 0.35)
        rate = torch.exp(b[None, :] + Z @ W.T) * depth[:, None]            # cells x genes
        lam = torch._standard_gamma(phi[None, :].expand_as(rate).contiguous(), generator=gen) * rate / phi[None, :]
        xt = torch.poisson(lam, generator=gen)
        nz = xt != 0
        counts.append(nz.sum(1).cpu())
        cg = nz.nonzero()                                                    # sorted by cell, then gene
        idx.append(cg[:, 1].to(torch.int32).cpu())
        data.append(xt[nz].cpu())
        del rate, lam, xt, nz, cg
    indptr = np.concatenate([[0], np.cumsum(torch.cat(counts).numpy())]).astype(np.int64)
    return sp.csc_matrix((torch.cat(data).numpy(), torch.cat(idx).numpy(), indptr), shape=(G, n))


def run_cfg4(args):
    """BASELINE.json configs[3]: sparse dgCMatrix 30k genes x 50k cells, library-size size factors,
    estimates.only = FALSE.  A step = saver(x, pred.genes = block, pred.genes.only = TRUE) with the
    default do.fast schedule on a fixed gene block of the matrix (the whole matrix is the design)."""
    import bench
    import saver_b200
    from saver_b200 import host
    torch, rank, world, local, dev, comm = _env()
    if comm is not None:
        raise SystemExit("cfg4 is a one-GPU line")
    h = saver_b200.Handle(local)
    G, n = int(os.environ.get("CFG4_G", 30000)), int(os.environ.get("CFG4_N", 50000))   # env: smoke sizes
    block = 2048 if args.block is None else args.block
    t0 = time.time()
    x = sparse_nb_counts(G, n, 20264, dev, torch)
    t_gen = time.time() - t0
    torch.cuda.empty_cache()
    pred = np.where(np.asarray(x.sum(1)).ravel() > 0)[0]
    genes = bench.pick_block(pred, block)
    be = host.CudaBackend(handle=h)
    kw = dict(do_fast=True, estimates_only=False, fold_seed=args.fold_seed, perm=args.fold_seed,
              pred_genes_only=True, backend=be)
    clocks = bench.ClockSampler(local)
    times, res = [], None
    l0 = h.launch_count()
    for it in range(args.steps + 1):   # the first pass (a few genes) warms the path
        h.synchronize()
        if it == 1:
            clocks.start()
            l0 = h.launch_count()
        t0 = time.time()
        res = saver_b200.saver(x, pred_genes=genes if it else genes[:32], **kw)
        h.synchronize()
        if it:
            times.append(time.time() - t0)
    clk = clocks.stop()
    info = res.info
    ncv = int((np.isfinite(info["sd.cv"]) & (info["lambda.max"] > 0)).sum())
    nfix = int(np.isnan(info["sd.cv"]).sum())
    tot = float(np.sum(times))
    B = genes.size
    line = {
        "metric": "genes/sec, saver(dgCMatrix, estimates.only = FALSE), synthetic NB %dk x %dk" % (G // 1000, n // 1000),
        "value": B * args.steps / tot, "unit": "genes/s", "n_gpus": 1, "steps": args.steps,
        "warmup": 1, "ms_per_step": tot * 1e3 / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg4: sparse dgCMatrix %d genes x %d cells (nnz %d = %.0f %% dense), library-size "
                               "size factors, do.fast=TRUE, estimates.only=FALSE, %d predictors; a step = a fixed "
                               "block of %d genes: %d cross-validated, %d fixed-lambda paths, the rest means"
                               % (G, n, x.nnz, 100.0 * x.nnz / (G * n), int((np.asarray(x.mean(1)).ravel() >= 0.1).sum()),
                                  B, ncv, nfix)},
        "clocks": clk,
        "e2e": {"value": B * args.steps / tot, "unit": "genes/s",
                "h2d_bytes_per_step": int(x.data.nbytes + x.indices.nbytes + x.indptr.nbytes),
                "d2h_bytes_per_step": int(res.estimate.nbytes * 2), "s_per_step": tot / args.steps,
                "api": "saver_b200.saver(scipy CSC on the host, pred_genes=block, pred_genes_only=True): the value "
                       "IS the end-to-end number (CSC upload, device sums / design, stage schedule, est + se D2H)"},
        "gpu_launches": int(h.launch_count() - l0),
        "roofline": None,
        "status_counts": np.bincount(res.status, minlength=3).tolist() if hasattr(res, "status") else None,
        "generate_s": t_gen,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_cfg5(args):
    import bench
    import saver_b200
    from saver_b200 import host, ops
    torch, rank, world, local, dev, comm = _env()
    h = saver_b200.Handle(local)
    stream = torch.cuda.ExternalStream(h.stream(), device=dev)
    G, n = (20000, 10000) if args.block is None else (args.block, 10000)
    # est / se of the shape saver() returns at cfg3 (posterior mean / sd of a Gamma): synthetic; every
    # rank holds only ITS genes (the layout saver(comm=, gather=False) leaves behind)
    e = [(G * r) // world for r in range(world + 1)]
    g0, g1 = e[rank], e[rank + 1]
    gen = torch.Generator(device=dev)
    gen.manual_seed(20265 + rank)
    est = torch.rand((g1 - g0, n), generator=gen, device=dev, dtype=torch.float64) * 3.0 + 0.05
    se = est * (0.2 + 0.5 * torch.rand((g1 - g0, n), generator=gen, device=dev, dtype=torch.float64))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = torch.empty((g1 - g0, n), dtype=torch.float64, device=dev)

    def barrier():
        h.synchronize()
        torch.cuda.synchronize()
        if comm is not None:
            comm.barrier()

    def samp(nb):
        def f():
            blk = 4096
            for b0 in range(0, g1 - g0, blk):
                b1 = min(b0 + blk, g1 - g0)
                out[b0:b1] = ops.sample(est[b0:b1], se[b0:b1], nb_mode=nb, seed=50, rep_idx=0,
                                        gene0=g0 + b0, handle=h)
        return f

    clocks = bench.ClockSampler(local)
    l0 = h.launch_count()
    clocks.start()
    t_gamma = _timed(samp(False), args.steps, args.warmup, stream, flush, torch, barrier)
    t_nb = _timed(samp(True), args.steps, args.warmup, stream, flush, torch, barrier)
    del out
    # cor.genes.  N = 1: the upper-triangle blocks of an 8 x 8 blocking, mirrored.  N > 1: all-gather of
    # the standardised rows (NCCL), every rank fills its G/N x G row panel, of which only the blocks of
    # its cyclic window are computed (the others are transposes of blocks other ranks own).
    res = {}

    def cor_one():
        nblk = 8
        Z, adj = ops.cor_prep(est, se, handle=h)
        Zp = torch.cat([Z, torch.zeros((128, Z.shape[1]), dtype=Z.dtype, device=dev)], 0)
        del Z
        full = torch.empty((G, G), dtype=torch.float64, device=dev)
        ed = [(G * b) // nblk for b in range(nblk + 1)]
        torch.cuda.synchronize()
        for i in range(nblk):
            for j in range(i, nblk):
                ops.cor_block(Zp, adj, G, ed[j], ed[j + 1] - ed[j], ed[i], ed[i + 1] - ed[i],
                              full[ed[i]:ed[i + 1]], handle=h)
        h.synchronize()
        for i in range(nblk):
            for j in range(i + 1, nblk):
                full[ed[j]:ed[j + 1], ed[i]:ed[i + 1]] = full[ed[i]:ed[i + 1], ed[j]:ed[j + 1]].T
        torch.cuda.synchronize()
        res["panel"], res["blocks"], res["flops"] = full, nblk * (nblk + 1) // 2, None

    def cor_many():
        res.pop("panel", None)
        panel, (c0, c1), blocks = host.cor_genes_sharded(est, se, comm, handle=h)
        res["panel"], res["blocks"] = panel, len(blocks)

    csteps = max(1, args.steps // 2)
    gathered0 = comm.bytes_gathered if comm is not None else 0
    t0 = None
    corf = cor_one if comm is None else cor_many
    corf()
    barrier()
    t0 = time.time()
    for _ in range(csteps):
        corf()
    barrier()
    t_cor = time.time() - t0
    clk = clocks.stop()
    launches = h.launch_count() - l0
    gathered = ((comm.bytes_gathered - gathered0) // (csteps + 1)) if comm is not None else 0
    tg, tn, tc = _max_over_ranks(comm, [t_gamma, t_nb, t_cor], dev, torch)
    if rank == 0:
        hbm, which = bench.peaks()
        pan = res["panel"]
        d = torch.diagonal(pan[:, g0:g1]).cpu().numpy()
        sb = 24.0 * n * G
        useful = 1.0 * G * G * n        # 2 n flop per entry of the upper triangle
        line = {
            "metric": "genes/sec, sample.saver (gamma) on saver() output, 20k x 10k",
            "value": G * args.steps / tg, "unit": "genes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tg * 1e3 / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg5: sample.saver(rep = 1, seed = 50) gamma and NB modes + cor.genes on "
                                   "est / se of %d genes x %d cells, genes in contiguous blocks over the "
                                   "ranks" % (G, n),
                       "l2": "flushed between steps (256 MiB write)"},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": None,
            "roofline": {"kernel": "sample_kernel (gamma mode)", "bound": "hbm", "unit": "GB/s",
                         "achieved": sb * args.steps / tg / 1e9, "peak": hbm * world,
                         "frac": sb * args.steps / tg / 1e9 / (hbm * world),
                         "peak_source": which, "traffic": None, "algorithmic": "24*n B/gene/rep"},
            "sample_nb": {"genes_per_s": G * args.steps / tn, "GBps": sb * args.steps / tn / 1e9,
                          "frac_hbm": sb * args.steps / tn / 1e9 / (hbm * world)},
            "cor_genes": {"s_per_call": tc / csteps,
                          "useful_tflops": useful * csteps / tc / 1e12,
                          "flops": "G^2 n (2 n per entry of the upper triangle; the blocks computed cover it "
                                   "plus half of the diagonal blocks); wall time of the whole call incl. "
                                   "standardisation, all-gather%s" % ("" if comm is not None else ", mirror copy"),
                          "blocks_on_rank0": res["blocks"],
                          "allgather_bytes": int(gathered),
                          "out_bytes_per_rank": int(pan.numel()) * 8,
                          "diag_is_adj2_le_1": bool((d <= 1.0 + 1e-12).all() and (d > 0).all())},
        }
        print(json.dumps(line), flush=True)
    _finish(comm)
