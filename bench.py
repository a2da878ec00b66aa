#!/usr/bin/env python
"""bench.py -- genes/sec of full saver() (cross-validated Poisson lasso + variance fit + posterior).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--workload cfg2]

One "step" = one pass of the whole hot path over the workload's gene block: every gene of the
count matrix goes through calc.estimate (R/calc_estimate.R:75-141) -- 5-fold cv.glmnet Poisson
lasso on the shared design, lambda.min prediction, calc.post -- i.e. saver(x, do.fast = FALSE,
estimates.only = TRUE) of the reference.  Workload "cfg2" = BASELINE.json configs[1]: synthetic
negative-binomial 5,000 genes x 2,000 cells on one B200.  With N > 1 (torchrun, one rank per
GPU) every rank runs its own 5,000-gene block against the same design (weak scaling; genes are
independent given the design, so there is no data-path collective).

Printed JSON (one line, rank 0): `value` = genes/s with all inputs resident in HBM (CUDA events on
the library's stream, max over ranks); `e2e` = the same through saver_b200.saver() with host
buffers (pinned), host<->device copies inside the timed region; `roofline` for the dominant
kernel; `cpu_baseline` = the CPU oracle (oracle/, a C restatement of the reference -- R and glmnet
cannot be installed here) on a bounded gene sample with all host threads.

--impl reference times that CPU restatement as the reference arm (the reference itself is pure R;
there is no R in the image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (genes, cells, seed)   -- SURVEY.md 8(d): seed = 20260 + config#
    "cfg2": (5000, 2000, 20262),
    "cfg3": (20000, 10000, 20263),
    "tiny": (256, 300, 20299),
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


def make_workload(name, rank):
    """Counts for this rank's gene block + the shared design inputs (rank 0's block)."""
    from saver_b200 import synth
    G, n, seed = WORKLOADS[name]
    x0 = synth.nb_counts(G, n, seed)  # the design always comes from block 0
    x = x0 if rank == 0 else synth.nb_counts(G, n, seed, gene_stream=rank)
    return x0, x, G, n


def prepare(x0, x, rank):
    """Host preprocessing exactly as saver() does it (R/saver.R:105-157)."""
    from saver_b200 import host
    x0c, _ = host.clean_data(x0)
    sf, scale_sf = host.calc_size_factor(x0c, None, x0c.shape[1])
    good = np.where(x0c.mean(1) >= 0.1)[0]
    x_est = np.log((x0c[good] + 1.0) / sf)
    xc = x0c if rank == 0 else np.asarray(x, dtype=np.float64)
    self_col = np.full(xc.shape[0], -1, dtype=np.int32)
    if rank == 0:
        self_col[good] = np.arange(good.size, dtype=np.int32)
    return xc, sf, scale_sf, x_est, self_col


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
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    h = saver_b200.Handle(local)
    x0, x, G, n = make_workload(args.workload, rank)
    xc, sf, scale_sf, x_est, self_col = prepare(x0, x, rank)
    p = x_est.shape[0]
    dev = torch.device("cuda", local)
    stream = torch.cuda.ExternalStream(h.stream(), device=dev)

    # ---------------- resident arm: everything already in HBM ----------------
    ops.set_design(x_est, handle=h)
    rs = xc.sum(1)
    pred = np.where(rs > 0)[0]
    seeds = np.arange(1, pred.size + 1)
    fold_h = host.draw_folds(seeds, n, 5, fold_seed=args.fold_seed)
    Xd = torch.from_numpy(np.ascontiguousarray(xc[pred])).to(dev)
    sfd = torch.from_numpy(sf).to(dev)
    scd = torch.from_numpy(self_col[pred]).to(dev)
    fold_d = torch.from_numpy(fold_h).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def resident_step():
        return ops.calc_estimate(Xd, sfd, scale_sf, 1, self_col=scd, foldid=fold_d, nfolds=5,
                                 calc_maxcor=False, want_se=False, handle=h)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        h.synchronize()

    for _ in range(args.warmup):
        flush.zero_()
        resident_step()
    barrier()
    c0 = np.zeros(8)
    h.lib.saver_cuda_counters(h.h, ops.ptr(c0))
    ext0 = np.zeros(4)
    h.lib.saver_cuda_counters_ext(h.h, ops.ptr(ext0))
    clocks = ClockSampler(local)
    clocks.start()
    evs = []
    t_host0 = time.time()
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
    c1 = np.zeros(8)
    h.lib.saver_cuda_counters(h.h, ops.ptr(c1))
    ext1 = np.zeros(4)
    h.lib.saver_cuda_counters_ext(h.h, ops.ptr(ext1))
    ms = sum(a.elapsed_time(b) for a, b in evs)
    dc = c1 - c0
    ngenes_step = pred.size
    # ---------------- e2e arm: the public API with host buffers ----------------
    x_pin = torch.from_numpy(np.ascontiguousarray(xc)).pin_memory().numpy()
    e2e_t = []
    h2d = xc.nbytes + x_est.nbytes + fold_h.nbytes + sf.nbytes + self_col.nbytes
    d2h = xc.nbytes
    if args.no_e2e:
        e2e_t = [float("nan")]
    else:
        # one small pass warms the host-buffer path (pinned staging, allocator), then the timed pass
        saver_b200.saver(x_pin[:64], do_fast=False, estimates_only=True, fold_seed=args.fold_seed,
                         backend=host.CudaBackend(handle=h))
        for it in range(args.e2e_steps):
            barrier()
            t0 = time.time()
            est = saver_b200.saver(x_pin, do_fast=False, estimates_only=True,
                                   fold_seed=args.fold_seed, backend=host.CudaBackend(handle=h))
            h.synchronize()
            e2e_t.append(time.time() - t0)
            assert np.isfinite(est).all()
            d2h = est.nbytes
    e2e_s = float(np.mean(e2e_t))

    tmax = ms / 1e3
    e2e_max = e2e_s
    if dist is not None:
        t = torch.tensor([tmax, e2e_max], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tmax, e2e_max = float(t[0]), float(t[1])
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    total_genes = ngenes_step * world * args.steps
    value = total_genes / tmax
    hbm, which = peaks()
    # dominant kernel: the lasso round kernel.  Its algorithmic HBM-side work is the design columns
    # it streams (n x 8 B each, counted in-kernel); its FP64 tensor work is the Gram blocks (DMMA).
    gram_flops = 2.0 * (ext1[2] - ext0[2])
    col_bytes = dc[4] * n * 8.0
    round_s = dc[0] / 1e3
    gemm_s = dc[1] / 1e3
    gemm_flops = dc[5] * 2.0 * n * p
    kname = "lasso_round_gram_kernel" if os.environ.get("SAVER_LASSO_GRAM", "1") != "0" else "lasso_round_kernel"
    # DRAM bytes of one launch of that kernel from the committed ncu --set full capture
    traffic, tnote = None, ""
    tp = os.path.join(ROOT, "profiles", "r1b_round_kernel_traffic.json")
    if kname == "lasso_round_gram_kernel" and os.path.exists(tp):
        td = json.load(open(tp))
        traffic = td["dram_bytes_per_launch"]
        tnote = (" traffic = dram read + write of the captured launch (%s); the same launch moved %.0f GB "
                 "from L2 to the SMs" % (td["capture"], td["l2_to_sm_bytes_per_launch"] / 1e9))
    roof = {"kernel": kname, "bound": "hbm",
            "achieved": col_bytes / round_s / 1e9 if round_s > 0 else None, "peak": hbm,
            "unit": "GB/s", "frac": (col_bytes / round_s / 1e9 / hbm) if round_s > 0 else None,
            "traffic": traffic, "peak_source": which,
            "note": "algorithmic bytes = design columns streamed x n x 8 B (counted in-kernel) over the "
                    "CUDA-event time of the round kernels; the columns are served by L2 (80% hits), the "
                    "kernel is latency bound (profiles/README.md)." + tnote,
            "gram_dmma_tflops_inside_kernel": gram_flops / round_s / 1e12 if round_s > 0 else None,
            "share_of_step": {kname: round_s / tmax, "gemm_tn_f64_kernel": gemm_s / tmax,
                              "lasso_setup": dc[2] / 1e3 / tmax, "calc_post_kernel": dc[3] / 1e3 / tmax},
            "gemm": {"bound": "fp64-mma", "achieved_tflops": gemm_flops / gemm_s / 1e12 if gemm_s > 0 else None,
                     "flops_per_column": 2.0 * n * p,
                     "note": "no FP64 peak in MEASURED_PEAKS.json; cuBLAS DGEMM measured 36 TF/s on these shapes"}}
    line = {
        "metric": METRIC, "value": value, "unit": "genes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tmax * 1e3 / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: synthetic negative-binomial %d genes x %d cells per GPU, "
                               "do.fast=FALSE (5-fold cv.glmnet Poisson lasso for every gene), "
                               "estimates.only=TRUE, %d predictors" % (args.workload, G, n, p),
                   "genes_per_step_per_gpu": int(ngenes_step), "l2": "flushed between steps (256 MiB write)",
                   "parallelism": "gene-sharded x%d, no data-path collective" % world},
        "clocks": clk,
        "e2e": {"value": ngenes_step * world / e2e_max, "unit": "genes/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "api": "saver_b200.saver(x, do_fast=False, estimates_only=True)",
                "s_per_step": e2e_max},
        "gpu_launches": int(dc[7]),
        "roofline": roof,
        "host_s_timed_region": t_host,
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(xc, sf, scale_sf, x_est, self_col, pred, fold_h, args)
    print(json.dumps(line), flush=True)


def cpu_baseline(xc, sf, scale_sf, x_est, self_col, pred, fold_h, args, budget_genes=None):
    """The CPU oracle (C restatement of glmnet's fishnet + cv glue + R's Brent) on a bounded gene
    sample of the same workload, all host threads."""
    import oracle
    cores = os.cpu_count() or 1
    k = budget_genes or max(cores, 8)
    sel = np.linspace(0, pred.size - 1, k).astype(int)
    g = pred[sel]
    y = xc[g] / sf
    t0 = time.time()
    r = oracle.predict_cv_batch(x_est, y, fold_h[sel], self_col=self_col[g], threads=cores)
    oracle.calc_post_batch(xc[g], r["mu"], sf, scale_sf, threads=cores)
    dt = time.time() - t0
    return {"value": k / dt, "unit": "genes/s", "cores": cores, "kind": "port",
            "sample": "%d genes of the same block (evenly spaced), full design, %.1f s" % (k, dt),
            "note": "oracle/saver_oracle.c (glmnet-style IRLS + coordinate descent, R's Brent); the "
                    "reference itself is R + glmnet, not installable in this image"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from saver_b200 import host
    x0, x, G, n = make_workload(args.workload, 0)
    xc, sf, scale_sf, x_est, self_col = prepare(x0, x, 0)
    pred = np.where(xc.sum(1) > 0)[0]
    cores = os.cpu_count() or 1
    k = max(cores, 4)  # the oracle runs one gene per thread: a step keeps every host thread busy
    times = []
    last = None
    for it in range(args.warmup + args.steps):
        sel = np.linspace(it, pred.size - 1 - (args.warmup + args.steps - it), k).astype(int)
        fold = host.draw_folds(sel + 1, n, 5, fold_seed=args.fold_seed)
        import oracle
        t0 = time.time()
        g = pred[sel]
        r = oracle.predict_cv_batch(x_est, xc[g] / sf, fold, self_col=self_col[g], threads=cores)
        oracle.calc_post_batch(xc[g], r["mu"], sf, scale_sf, threads=cores)
        dt = time.time() - t0
        if it >= args.warmup:
            times.append(dt)
        last = dt
    tot = float(np.sum(times))
    value = k * len(times) / tot
    p = x_est.shape[0]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "genes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot * 1e3 / max(len(times), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: synthetic negative-binomial %d genes x %d cells per GPU, "
                               "do.fast=FALSE (5-fold cv.glmnet Poisson lasso for every gene), "
                               "estimates.only=TRUE, %d predictors" % (args.workload, G, n, p)},
        "cpu_baseline": {"value": value, "unit": "genes/s", "cores": cores, "kind": "port",
                         "sample": "%d genes per step (evenly spaced through the block), full design; "
                                   "last step %.1f s" % (k, last)},
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
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--fold-seed", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer (e2e) arm")
    ap.add_argument("--e2e-steps", type=int, default=1)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
