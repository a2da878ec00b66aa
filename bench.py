#!/usr/bin/env python
"""bench.py -- genes/sec of full saver() (cross-validated Poisson lasso + variance fit + posterior).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--workload cfg3]

Default workload "cfg3" = BASELINE.json's metric: synthetic negative-binomial 20,000 genes x
10,000 cells, do.fast = FALSE (every gene gets 5-fold cv.glmnet), estimates.only = TRUE, the full
design (every gene with mean >= 0.1 is a predictor).  The whole matrix is hours of work, so one
"step" = one pass of the hot path -- calc.estimate, R/calc_estimate.R:75-141: cv Poisson lasso on
the shared design, lambda.min prediction, calc.post -- over a FIXED block of genes of that matrix
(`--block`, evenly spaced through the predictable genes); i.e. saver(x, pred.genes = block,
pred.genes.only = TRUE, do.fast = FALSE, estimates.only = TRUE) of the reference.  With N > 1
(torchrun, one rank per GPU) the SAME block is dealt over the ranks (strong scaling): rank 0 holds
x, the design goes to the other ranks with one NCCL broadcast, every rank keeps its own estimate
rows (R/calc_estimate.R:69-74,147-160: one x, chunks over workers, x.est exported to each).

Printed JSON (one line, rank 0):
  value     genes/s with design, counts and folds resident in HBM (CUDA events on the library's
            stream, max over ranks, L2 flushed between steps);
  e2e       the same through saver_b200.saver() from HOST buffers on rank 0: host preparation, H2D of
            the design, NCCL broadcast, per-rank counts / folds H2D, D2H of each rank's estimate rows
            all inside the timed region;
  roofline  the prediction phase against the FP64 GEMM peak measured in this run (SURVEY.md 8(d):
            2 n p flop per full-gradient evaluation), with the per-kernel numbers beside it;
  cpu_baseline / parity  the CPU oracle (oracle/, a C restatement of glmnet's fishnet + cv.glmnet +
            R's Brent; R and glmnet cannot be installed here) on a few genes of the block: its time is
            the CPU baseline, its results are the parity check of this very run.

Other workloads: cfg2 (5,000 x 2,000, BASELINE.json configs[1]), cfg4 (dgCMatrix 30,000 x 50,000,
do.fast = TRUE, estimate + SE, a 2,048-gene block), cfg5 (sample.saver + cor.genes on 20,000 x 10,000
est / se, genes sharded over the ranks), null (saver(null.model = TRUE) on 20,000 x 10,000: the
variance-fit / posterior kernel alone), fast (do.fast = TRUE schedule on 5,000 x 2,000), tiny.
--impl reference times the CPU restatement as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: genes, cells, seed (seed = 20260 + config#), default gene block per step
    "cfg3": dict(G=20000, n=10000, seed=20263, block=512),
    "cfg2": dict(G=5000, n=2000, seed=20262, block=0),      # 0 = every predictable gene
    "tiny": dict(G=256, n=300, seed=20299, block=0),
    "null": dict(G=20000, n=10000, seed=20263, block=0),
    "fast": dict(G=5000, n=2000, seed=20262, block=0),
    "cfg5": dict(G=20000, n=10000, seed=20265, block=0),
    "cfg4": dict(G=30000, n=50000, seed=20264, block=2048),
}
METRIC = "genes/sec, full saver() (cv Poisson lasso + variance fit + posterior), synthetic NB"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 2 + k and r[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------
def make_counts(name, gene_stream=0):
    from saver_b200 import synth
    w = WORKLOADS[name]
    return synth.nb_counts(w["G"], w["n"], w["seed"], gene_stream=gene_stream)


def prepare(x):
    """Host preparation exactly as saver() does it (R/saver.R:105-157)."""
    from saver_b200 import host
    xc, _ = host.clean_data(x)
    sf, scale_sf = host.calc_size_factor(xc, None, xc.shape[1])
    good = np.where(xc.mean(1) >= 0.1)[0]
    x_est = np.log((xc[good] + 1.0) / sf)
    self_col = np.full(xc.shape[0], -1, dtype=np.int32)
    self_col[good] = np.arange(good.size, dtype=np.int32)
    pred = np.where(xc.sum(1) > 0)[0]
    return xc, sf, scale_sf, x_est, self_col, pred


def pick_block(pred, block):
    if block <= 0 or block >= pred.size:
        return pred
    return pred[np.linspace(0, pred.size - 1, block).astype(np.int64)]


def stage_layout(B):
    """Stages of saver.fit's slow schedule (R/saver_fit.R:333-462: n1 = 8, a quarter of the rest,
    the rest) for a block of B genes in matrix order: [(start, stop)]."""
    n1 = min(8, B)
    n2 = min(int(np.ceil((B - n1) / 4)) + n1, B)
    return [(a, b) for a, b in ((0, n1), (n1, n2), (n2, B)) if b > a]


def deal(B, rank, world):
    """Block positions this rank computes and their cv seeds j -- exactly what saver() does with the
    block: genes dealt round-robin over the ranks, j = 1-based position inside the gene's stage
    (R/calc_estimate.R:98), so the resident arm, the e2e arm and the oracle see the same folds."""
    seed_of = np.concatenate([np.arange(1, b - a + 1) for a, b in stage_layout(B)])
    pos = np.arange(B)[rank::world]
    return pos, seed_of[pos]


def counters(h):
    from saver_b200 import ops
    c, e = np.zeros(8), np.zeros(4)
    h.lib.saver_cuda_counters(h.h, ops.ptr(c))
    h.lib.saver_cuda_counters_ext(h.h, ops.ptr(e))
    return c, e


def fp64_gemm_peak(dev, m=6144, reps=4):
    """cuBLAS DGEMM TF/s on this box, in this run (no FP64 figure is in MEASURED_PEAKS.json)."""
    import torch
    a = torch.randn(m, m, dtype=torch.float64, device=dev)
    b = torch.randn(m, m, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * m ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    return best


def oracle_threads():
    cores = os.cpu_count() or 1
    outer = 5 if cores >= 10 else 1
    return cores, outer, max(1, cores // outer)


def oracle_sample(x_est, xc, sf, scale_sf, self_col, genes, seeds, fold_seed, inside_gene):
    """CPU oracle on a few genes: (seconds, dict of results).  inside_gene: threads split the column
    loops and fold fits of ONE gene at a time (20k x 10k: a gene is minutes of scalar work);
    otherwise one gene per thread."""
    import oracle
    from saver_b200 import host
    n = xc.shape[1]
    cores, outer, inner = oracle_threads()
    fold = host.draw_folds(seeds, n, 5, fold_seed=fold_seed)
    y = xc[genes] / sf
    t0 = time.time()
    if inside_gene:
        oracle.set_threads(outer, inner)
        r = oracle.predict_cv_batch(x_est, y, fold, self_col=self_col[genes], threads=1)
        oracle.set_threads(1, 1)
    else:
        r = oracle.predict_cv_batch(x_est, y, fold, self_col=self_col[genes], threads=cores)
    post = oracle.calc_post_batch(xc[genes], r["mu"], sf, scale_sf, threads=cores)
    return time.time() - t0, dict(est=post["est"], idx=r["idx"], lmin=r["lambda_min"], rc=r["rc"])


# ------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import saver_b200
    from saver_b200 import host, ops
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libsaver_cuda has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        from saver_b200 import dist as sdist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=dev)
        comm = sdist.Comm(init=False)
    h = saver_b200.Handle(local)
    stream = torch.cuda.ExternalStream(h.stream(), device=dev)
    w = WORKLOADS[args.workload]
    n = w["n"]
    block = w["block"] if args.block is None else args.block

    def barrier():
        h.synchronize()
        torch.cuda.synchronize()
        if comm is not None:
            comm.barrier()

    # ---------------- inputs: rank 0 holds the matrix ----------------
    t_setup = time.time()
    x_host = xc = x_est = None
    if rank == 0:
        x_host = make_counts(args.workload, args.gene_stream)
        xc, sf, scale_sf, x_est, self_col, pred = prepare(x_host)
        genes = pick_block(pred, block)
        plan = dict(sf=sf, scale_sf=scale_sf, genes=genes, self_col=self_col[genes], p=x_est.shape[0])
    else:
        plan = None
    if comm is not None:
        plan = comm.bcast_obj(plan)
    sf, scale_sf, genes, p = plan["sf"], plan["scale_sf"], plan["genes"], plan["p"]
    B = genes.size
    mine, seeds = deal(B, rank, world)
    # resident arm set-up (untimed): design upload + one NCCL broadcast, the block's counts
    bc_ms = None
    if comm is not None:
        xd = torch.from_numpy(x_est).to(dev) if rank == 0 else None
        comm.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        xd = comm.broadcast_array(xd, keep_on_device=True)
        e1.record()
        torch.cuda.synchronize()
        bc_ms = e0.elapsed_time(e1)
        ops.set_design(xd, handle=h)
        del xd
        xblk = comm.broadcast_array(xc[genes] if rank == 0 else None)
    else:
        ops.set_design(x_est, handle=h)
        xblk = xc[genes]
    h.synchronize()
    fold_h = host.draw_folds(seeds, n, 5, fold_seed=args.fold_seed)
    Xd = torch.from_numpy(np.ascontiguousarray(xblk[mine])).to(dev)
    sfd = torch.from_numpy(sf).to(dev)
    scd = torch.from_numpy(np.ascontiguousarray(plan["self_col"][mine])).to(dev)
    fold_d = torch.from_numpy(fold_h).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dgemm_tf = fp64_gemm_peak(dev) if rank == 0 else None
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup

    def resident_step():
        if mine.size == 0:
            return None
        return ops.calc_estimate(Xd, sfd, scale_sf, 1, self_col=scd, foldid=fold_d, nfolds=5,
                                 calc_maxcor=False, want_se=False, handle=h)

    # ---------------- resident arm ----------------
    for _ in range(args.warmup):
        flush.zero_()
        resident_step()
    barrier()
    c0, ext0 = counters(h)
    clocks = ClockSampler(local)
    clocks.start()
    evs = []
    t_host0 = time.time()
    out = None
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = resident_step()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_host = time.time() - t_host0
    clk = clocks.stop()
    c1, ext1 = counters(h)
    ms = sum(a.elapsed_time(b) for a, b in evs)
    dc = c1 - c0
    est_res = out["est"].cpu().numpy() if out is not None else np.zeros((0, n))
    status_res = out["status"].cpu().numpy() if out is not None else np.zeros(0, dtype=np.int32)
    lmin_res = out["lambda_min"].cpu().numpy() if out is not None else np.zeros(0)

    # ---------------- e2e arm: the public API, host buffers on rank 0 ----------------
    e2e_t, h2d, d2h, nvl = [], 0, 0, 0
    e2e_est = None
    if not args.no_e2e:
        be = host.CudaBackend(handle=h, comm=comm)
        kw = dict(do_fast=False, estimates_only=True, fold_seed=args.fold_seed, perm="identity",
                  pred_genes_only=True, backend=be, comm=comm, gather=False)
        x_pin = None
        if rank == 0:
            x_pin = torch.from_numpy(np.ascontiguousarray(x_host)).pin_memory().numpy()
        for it in range(args.e2e_steps + 1):   # the first pass warms the host-buffer path
            nv0 = comm.bytes_broadcast if comm is not None else 0
            barrier()
            t0 = time.time()
            res = saver_b200.saver(x_pin, pred_genes=genes if it else genes[:max(2 * world, 4)], **kw)
            h.synchronize()
            dt = time.time() - t0
            if it:
                e2e_t.append(dt)
                own = np.where(res.owner == rank)[0]
                e2e_est = res.estimate[own]
                assert np.isfinite(e2e_est).all()
                d2h = e2e_est.nbytes
                nvl = (comm.bytes_broadcast - nv0) if comm is not None else 0
                # H2D on this rank: the design (rank 0 only), its genes' counts and folds, sf, self_col
                h2d = (p * n * 8 if rank == 0 else 0) + own.size * n * (8 + 4) + n * 8 + own.size * 4
    e2e_s = float(np.mean(e2e_t)) if e2e_t else float("nan")

    tmax, e2e_max = ms / 1e3, e2e_s
    sums = np.array([dc[0], dc[1], dc[2], dc[3], dc[5], dc[7], float(h2d), float(d2h)])
    if comm is not None:
        t = torch.tensor([tmax, e2e_max], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        tmax, e2e_max = float(t[0]), float(t[1])
        s = torch.from_numpy(sums).to(dev)
        tdist.all_reduce(s)
        sums = s.cpu().numpy()
        nranks_seen = int(comm.max_scalar(world))
    if rank != 0:
        if comm is not None:
            comm.barrier()
            tdist.destroy_process_group()
        return

    # ---------------- parity sample + CPU baseline (rank 0, outside every timed region) ----------
    parity, cpu = None, None
    k = args.parity if args.parity is not None else (0 if (world > 1 or args.no_cpu) else
                                                     (2 if n >= 5000 else 2 * (os.cpu_count() or 1)))
    k = min(k, mine.size)
    if k > 0:
        sel = np.linspace(0, mine.size - 1, k).astype(int)
        inside = n >= 5000
        dt, ref = oracle_sample(x_est, xc, sf, scale_sf, self_col, genes[mine[sel]], seeds[sel],
                                args.fold_seed, inside)
        got = est_res[sel]
        rel = np.abs(got - ref["est"]) / np.maximum(ref["est"], 1e-12)
        parity = {"genes": int(k), "median_rel_err_estimate": float(np.median(rel)),
                  "per_gene_median_rel_err": [float(v) for v in np.median(rel, 1)][:16],
                  "lambda_min_equal": int(np.isclose(ref["lmin"], lmin_res[sel], rtol=1e-9, atol=0).sum()),
                  "status_equal": int((ref["rc"] == status_res[sel]).sum()),
                  "tolerance": "median <= 1e-3 (north_star)", "ok": bool(np.median(rel) <= 1e-3),
                  "against": "oracle/saver_oracle.c on the same genes and folds, outside the timed region"}
        if e2e_est is not None and world == 1:   # same genes, same folds: the e2e result is checked too
            rel2 = np.abs(e2e_est[mine[sel]] - ref["est"]) / np.maximum(ref["est"], 1e-12)
            parity["e2e_median_rel_err_estimate"] = float(np.median(rel2))
        cores, outer, inner = oracle_threads()
        cpu = {"value": k / dt, "unit": "genes/s", "cores": cores, "kind": "port",
               "sample": "%d genes of the block, full design, %.1f s%s" % (
                   k, dt, " (threads inside a gene: %d fold fits x %d column threads)" % (outer, inner)
                   if inside else " (one gene per thread)"),
               "note": "oracle/saver_oracle.c (glmnet-style IRLS + coordinate descent, R's Brent); the "
                       "reference itself is R + glmnet, not installable in this image"}

    # ---------------- the line ----------------
    hbm, which = peaks()
    round_s, gemm_s, setup_s, post_s, gemm_cols, launches = (sums[0] / 1e3, sums[1] / 1e3, sums[2] / 1e3,
                                                              sums[3] / 1e3, sums[4], sums[5])
    genes_done = B * args.steps
    value = genes_done / tmax
    pred_flops = 2.0 * n * p * gemm_cols                  # 2 n p per gradient evaluation
    pred_busy = round_s + gemm_s + setup_s                # GPU-seconds summed over ranks
    pred_tf = pred_flops / pred_busy / 1e12 if pred_busy > 0 else None
    traffic, tnote = None, None
    tp = os.path.join(ROOT, "profiles", "r2_round_kernel_traffic.json")
    if os.path.exists(tp):
        td = json.load(open(tp))
        traffic, tnote = td.get("dram_bytes_per_launch"), td.get("capture")
    post_bytes = 24.0 * n * genes_done                   # read y, mu + write est (estimates.only)
    roof = {
        "kernel": "prediction phase = lasso_round_gram_kernel + gemm_tn_f64_kernel",
        "bound": "tensor", "unit": "TFLOP/s",
        "achieved": pred_tf, "peak": dgemm_tf, "frac": pred_tf / dgemm_tf if pred_tf and dgemm_tf else None,
        "peak_source": "cuBLAS DGEMM %d^3 via torch.matmul, measured in this run" % 6144,
        "traffic": traffic, "traffic_source": tnote,
        "algorithmic": "2*n*p flop per full-gradient evaluation x %.0f evaluations/gene "
                       "(floor (F+1)*L_g)" % (gemm_cols / max(genes_done, 1)),
        "share_of_gpu_time": {"lasso_round_gram_kernel": round_s / max(pred_busy + post_s, 1e-9),
                              "gemm_tn_f64_kernel": gemm_s / max(pred_busy + post_s, 1e-9),
                              "lasso_setup": setup_s / max(pred_busy + post_s, 1e-9),
                              "calc_post_kernel": post_s / max(pred_busy + post_s, 1e-9)},
        "kernels": {
            "gemm_tn_f64_kernel": {"bound": "tensor", "unit": "TFLOP/s",
                                   "achieved": pred_flops / gemm_s / 1e12 if gemm_s > 0 else None,
                                   "peak": dgemm_tf,
                                   "frac": pred_flops / gemm_s / 1e12 / dgemm_tf if gemm_s > 0 and dgemm_tf else None},
            "calc_post_kernel": {"bound": "hbm", "unit": "GB/s",
                                 "achieved": post_bytes / post_s / 1e9 if post_s > 0 else None,
                                 "peak": hbm, "peak_source": which,
                                 "frac": post_bytes / post_s / 1e9 / hbm if post_s > 0 else None,
                                 "algorithmic": "24*n B/gene (estimates.only); the kernel is FP64-"
                                                "transcendental bound, see profiles/"},
        },
    }
    line = {
        "metric": METRIC, "value": value, "unit": "genes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tmax * 1e3 / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.workload, w, B, p, world),
        "clocks": clk,
        "e2e": {"value": B / e2e_max if e2e_max == e2e_max else None, "unit": "genes/s",
                "h2d_bytes_per_step": int(sums[6]), "d2h_bytes_per_step": int(sums[7]),
                "nvlink_bytes_per_step": int(nvl) * max(world - 1, 0),
                "api": "saver_b200.saver(x on rank 0, pred_genes=block, pred_genes_only=True, "
                       "do_fast=False, estimates_only=True, gather=False)",
                "s_per_step": e2e_max},
        "gpu_launches": int(launches),
        "roofline": roof,
        "fp64_gemm_peak_tflops": dgemm_tf,
        "design_broadcast": None if bc_ms is None else {"ms": bc_ms, "bytes": int(p) * n * 8,
                                                        "GBps": p * n * 8 / bc_ms / 1e6},
        "comm_nranks_seen": nranks_seen if comm is not None else 1,
        "status_counts": np.bincount(status_res, minlength=3).tolist(),
        "host_s_timed_region": t_host, "setup_s": t_setup,
    }
    if int(os.environ.get("SAVER_LASSO_DBG", "0")) & 128:   # per-phase SM cycles of the round kernel (diagnostic)
        ph = np.zeros(16)
        h.lib.saver_cuda_phase_cycles(h.h, ops.ptr(ph))
        names = ["phaseA", "vectors", "gramP", "sweeps", "material", "inactive", "relin"]
        tot = ph[:7].sum() + ph[13]
        line["phases"] = {k: round(float(v / tot), 4) for k, v in zip(names, ph[:7])}
        line["phases"].update(activate=round(float(ph[13] / tot), 4), irls_per_gene=float(ph[7] / max(1, B * (args.steps + args.warmup))),
                              mean_m=float(ph[12] / max(ph[7], 1)), inactive_cols_per_irls=float(ph[10] / max(ph[7], 1)),
                              activations_per_irls=float(ph[14] / max(ph[7], 1)), sweep_steps_per_irls=float(ph[8] / max(ph[7], 1)))
    if parity is not None:
        line["parity"] = parity
    if cpu is not None and world == 1:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if comm is not None:
        comm.barrier()
        tdist.destroy_process_group()


def config_dict(name, w, B, p, world):
    """The `config` of the JSON line -- the same dict from both arms (the reference arm describes the
    CUDA arm's workload; the L2 / parallelism entries say how the CUDA arm runs it)."""
    return {"workload": workload_name(name, w, B, p),
            "genes_per_step": int(B), "genes_per_step_per_gpu": int(np.ceil(B / world)),
            "l2": "flushed between steps (256 MiB write)",
            "parallelism": "one gene block dealt round-robin over %d rank(s); design = one NCCL "
                           "broadcast; no data-path collective" % world}


def workload_name(name, w, B, p):
    return ("%s: synthetic negative-binomial %d genes x %d cells, do.fast=FALSE (5-fold cv.glmnet "
            "Poisson lasso for every gene of the step's block), estimates.only=TRUE, %d predictors, "
            "%d genes per step" % (name, w["G"], w["n"], p, B))


# ------------------------------------------------------------------------------------------------
# the reference arm: the CPU restatement on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import oracle
    from saver_b200 import host
    w = WORKLOADS[args.workload]
    n = w["n"]
    block = w["block"] if args.block is None else args.block
    x_host = make_counts(args.workload, args.gene_stream)
    xc, sf, scale_sf, x_est, self_col, pred = prepare(x_host)
    genes = pick_block(pred, block)
    B, p = genes.size, x_est.shape[0]
    cores, outer, inner = oracle_threads()
    inside = n >= 5000
    nsteps = args.warmup + args.steps
    times, done = [], []
    last = 0.0
    mine, seeds = deal(B, 0, 1)
    if inside:
        # A gene is minutes of scalar work at this shape, so a step is ONE glmnet fit of the gene's
        # six (the master fit, then the five fold fits on the master's lambda list), with every host
        # thread working inside the fit (column loops of the KKT / curvature passes); 6 steps = 1 gene.
        oracle.set_threads(1, cores)
        sample = ("one glmnet fit per step (master fit, then the 5 fold fits on its lambda list; 6 steps "
                  "= 1 gene), %d threads inside the fit" % cores)
        state = {}
        for it in range(nsteps):
            gi, f = divmod(it, 6)
            pos = (gi * max(1, B // (nsteps // 6 + 1))) % B
            g = genes[pos]
            y = xc[g] / sf
            t0 = time.time()
            if f == 0 or "lam" not in state:
                r = oracle.glmnet_poisson(x_est, y, excl=int(self_col[g]))
                state = {"lam": r["lam"], "fold": host.draw_folds([seeds[pos]], n, 5, args.fold_seed)[0]}
            else:
                rows = np.where(state["fold"] != f)[0]
                r = oracle.glmnet_poisson(x_est, y, rows=rows, excl=int(self_col[g]), lam=state["lam"])
            dt = time.time() - t0
            if it >= args.warmup:
                times.append(dt)
                done.append(1.0 / 6.0)
            last = dt
        oracle.set_threads(1, 1)
    else:
        per_step = 2 * cores
        sample = "%d genes of the block per step, one gene per thread (dynamic)" % per_step
        for it in range(nsteps):
            sel = (np.arange(per_step) * max(1, B // (per_step * nsteps)) + it * per_step) % B
            dt, _ = oracle_sample(x_est, xc, sf, scale_sf, self_col, genes[sel], seeds[sel],
                                  args.fold_seed, inside)
            if it >= args.warmup:
                times.append(dt)
                done.append(sel.size)
            last = dt
    tot = float(np.sum(times))
    value = float(np.sum(done)) / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "genes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot * 1e3 / max(len(times), 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(args.workload, w, B, p, world),
        "cpu_baseline": {"value": value, "unit": "genes/s", "cores": cores, "kind": "port",
                         "sample": sample + "; last step %.1f s" % last},
        "e2e": {"value": value, "unit": "genes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference path (oracle/saver_oracle.c) on all host threads; "
                "the reference is R + glmnet (Fortran) and neither is installable here (DESIGN.md)",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--block", type=int, default=None, help="genes per step (default: the workload's)")
    ap.add_argument("--gene-stream", type=int, default=0,
                    help="draw another block of genes over the same cells (synth.nb_counts)")
    ap.add_argument("--fold-seed", type=int, default=5)
    ap.add_argument("--parity", type=int, default=None, help="genes checked against the CPU oracle")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer (e2e) arm")
    ap.add_argument("--e2e-steps", type=int, default=1)
    args = ap.parse_args()
    try:
        if args.workload in ("null", "fast", "cfg5", "cfg4") and args.impl == "cuda":
            import bench_extra
            bench_extra.run(args)
        elif args.impl == "reference":
            run_reference(args)
        else:
            run_cuda(args)
    except BaseException:
        # torchrun swallows a rank's traceback: print it (and the library's last error is part of
        # the exception text) before the non-zero exit
        sys.stderr.write("[bench.py rank %s] failed:\n%s\n" % (os.environ.get("RANK", "0"),
                                                                traceback.format_exc()))
        sys.stderr.flush()
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "bench_rank%s.err" % os.environ.get("RANK", "0")), "w") as f:
                f.write(traceback.format_exc())
        except OSError:
            pass
        raise


if __name__ == "__main__":
    main()
