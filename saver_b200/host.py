"""Host-side mirror of the reference's R API for the expression-recovery path.

R is not installable in the build image, so the host layer that the reference
writes in R (R/saver.R, R/saver_fit.R, R/calc_estimate.R, R/utils.R) is mirrored here in Python
with the same names, argument meaning and error behaviour; INTEGRATION.md shows the R wrappers and
the ``.Call`` shim that bind the same C ABI.  All numerical work goes through libsaver_cuda
(``saver_cuda_calc_estimate`` = one call per stage block); there is no CPU fallback.

Mapping (reference file:line -> here):
    saver                R/saver.R:98-179          saver()
    saver.fit            R/saver_fit.R:56-463      saver_fit()
    saver.fit.mean/.null R/saver_fit.R:467-628     saver_fit_mean(), saver_fit_null()
    calc.estimate        R/calc_estimate.R:65-161  CudaBackend.calc_estimate()
    clean.data etc.      R/utils.R                 clean_data(), calc_size_factor(), ...
"""
import time

import numpy as np

from . import ops


class SaverError(ValueError):
    """R's stop(): argument validation failures of the reference API."""


# ----------------------------------------------------------------------------------------------
# R/utils.R
# ----------------------------------------------------------------------------------------------
def clean_data(x):
    """clean.data (R/utils.R:2-31): values < 0.001 -> 0, drop all-zero cells.

    Returns (x, kept_cells).  x is genes x cells; scipy.sparse matrices are densified per block by
    the caller, like ``as.matrix(x)`` in R/calc_estimate.R:69."""
    sparse = hasattr(x, "tocsc")
    if sparse:
        x = x.tocsc(copy=True).astype(np.float64)
        x.data[x.data < 0.001] = 0
        x.eliminate_zeros()
        if x.ndim != 2 or x.shape[1] <= 1:
            raise SaverError("x should be a matrix with 2 or more columns")
        cs = np.asarray(x.sum(0)).ravel()
    else:
        x = np.array(x, dtype=np.float64, copy=True)
        if x.ndim != 2 or x.shape[1] <= 1:
            raise SaverError("x should be a matrix with 2 or more columns")
        x[x < 0.001] = 0
        cs = x.sum(0)
    keep = cs != 0
    if not keep.all():
        x = x[:, np.where(keep)[0]]
    return x, keep


def check_mu(x, mu):
    """check.mu (R/utils.R:33-51)."""
    mu = np.asarray(mu, dtype=np.float64)
    if mu.shape != tuple(x.shape):
        raise SaverError("x and mu must have same dimensions")
    cs = np.asarray(x.sum(0)).ravel()
    if cs.min() == 0:
        mu = mu[:, cs != 0]
    return mu


def calc_size_factor(x, size_factor, ncells):
    """calc.size.factor (R/utils.R:70-86) -> (sf, scale_sf)."""
    if size_factor is None:
        cs = np.asarray(x.sum(0)).ravel().astype(np.float64)
        return cs / cs.mean(), 1.0
    sfv = np.atleast_1d(np.asarray(size_factor, dtype=np.float64))
    if sfv.size == ncells:
        if sfv.min() <= 0:
            raise SaverError("Size factor must be greater than 0")
        return sfv / sfv.mean(), float(sfv.mean())
    if sfv.size == 1 and sfv[0] == 1:
        return np.ones(ncells), 1.0
    if sfv.min() <= 0:
        raise SaverError("Size factor must be greater than 0")
    raise SaverError("Not a valid size factor")


def get_pred_cells(pred_cells, ncells):
    """get.pred.cells (R/utils.R:88-98); 0-based indices here."""
    if pred_cells is None:
        return np.arange(ncells)
    pc = np.asarray(pred_cells, dtype=np.int64)
    if pc.min() < 0 or pc.max() >= ncells:
        raise SaverError("pred.cells must be column indices of x")
    return pc


def get_pred_genes(x, pred_genes, npred, ngenes):
    """get.pred.genes (R/utils.R:100-115); 0-based indices here."""
    rs = np.asarray(x.sum(1)).ravel()
    if pred_genes is not None:
        pg = np.asarray(pred_genes, dtype=np.int64)
        if pg.min() < 0 or pg.max() >= ngenes:
            raise SaverError("pred.genes must be row indices of x")
        return pg
    if npred is None:
        return np.where(rs != 0)[0]
    if npred < ngenes:
        npred = min(int((rs != 0).sum()), int(npred))
        return np.argsort(-rs / x.shape[1], kind="stable")[:npred]
    raise SaverError("npred must be less than number of rows in x")


def get_chunk(length, n):
    """get.chunk (R/utils.R:118-126)."""
    cs = np.stack([np.arange(100, 151), np.arange(99, 48, -1)], 1).ravel()
    r = np.zeros(cs.size)
    for i, c in enumerate(cs):
        r[i] = int(np.ceil(length / c)) % n
        if r[i] == 0:
            return int(c)
    return int(cs[int(np.argmax(r))])


def est_lambda(y, r, coefs):
    """est.lambda (R/utils.R:129-134)."""
    y = np.asarray(y, dtype=np.float64)
    n = y.size
    lmax = r * y.std(ddof=1) * np.sqrt((n - 1) / n)
    lmin = lmax * np.exp(-np.sqrt(max(0.0, coefs[0] + coefs[1] * r)))
    return np.array([lmax, lmin])


def _lm(y, x):
    """lm(y ~ x)$coefficients with NA / non-finite rows dropped (na.omit)."""
    ok = np.isfinite(x) & np.isfinite(y)
    A = np.stack([np.ones(ok.sum()), x[ok]], 1)
    return np.linalg.lstsq(A, y[ok], rcond=None)[0]


# ----------------------------------------------------------------------------------------------
# fold assignment = what cv.glmnet draws after set.seed(j) (R/expr_predict.R:35-43)
# ----------------------------------------------------------------------------------------------
def gene_order(perm, pred1, rest, ngenes):
    """The processing order ``ind`` of saver.fit (R/saver_fit.R:87-92): the predictable genes in
    random order, then the others.  perm: "identity"; an int (numpy stream); or ("R", seed) /
    ("R-3.5", seed) -- the order the reference produces when the user calls set.seed(seed) right
    before saver(): c(sample(pred.genes1, npred1), sample(others, ngenes - npred1)), or
    sample(1:ngenes, ngenes) when every gene is predictable, on R's own stream with the current
    ("Rejection", R >= 3.6) or the older ("Rounding") sampler."""
    if isinstance(perm, str) and perm == "identity":
        return np.concatenate([pred1, rest])
    if isinstance(perm, (tuple, list)) and perm[0] in ("R", "R-3.5"):
        from . import rrng
        r = rrng.RRng(int(perm[1]), "rejection" if perm[0] == "R" else "rounding")
        if pred1.size < ngenes:
            return np.concatenate([r.sample(pred1), r.sample(rest)])
        return r.sample(np.arange(ngenes))
    rng = np.random.default_rng(perm)
    return np.concatenate([rng.permutation(pred1), rng.permutation(rest)])


def draw_folds(seeds, N, nfolds=5, fold_seed=0, method="philox"):
    """foldid for each gene: a permutation of rep(1..nfolds, length = N).

    seeds: the per-gene ``j`` of R/calc_estimate.R:98 (1-based position inside the stage block).
    method "philox": counter-based stream keyed on (fold_seed, j) -- reproducible and independent
    of how genes are split over blocks or GPUs.  (An R host draws the folds itself with
    set.seed(j); sample(...), see INTEGRATION.md.)
    method "R" / "R-3.5": exactly the folds the reference draws -- set.seed(j) followed by
    cv.glmnet's sample(rep(seq(nfolds), length = N)) -- with R's current sampler (R >= 3.6,
    sample.kind "Rejection") or the one before it ("Rounding"); fold_seed is ignored
    (saver_b200/rrng.py)."""
    seeds = np.asarray(seeds, dtype=np.int64)
    base = (np.arange(N) % nfolds + 1).astype(np.int32)
    out = np.empty((seeds.size, N), dtype=np.int32)
    if method in ("R", "R-3.5"):
        from . import rrng
        kind = "rejection" if method == "R" else "rounding"
        for i, j in enumerate(seeds):
            out[i] = rrng.cv_folds(int(j), N, nfolds, kind)
        return out
    if method != "philox":
        raise SaverError("unknown fold method %r" % (method,))
    for i, j in enumerate(seeds):
        rng = np.random.Generator(np.random.Philox(key=[int(fold_seed) & (2**63 - 1), int(j)]))
        out[i] = base[np.argsort(rng.random(N), kind="stable")]
    return out


# ----------------------------------------------------------------------------------------------
# calc.estimate on the device
# ----------------------------------------------------------------------------------------------
class CudaBackend:
    """Runs calc.estimate blocks through libsaver_cuda on one device.

    ``rank`` / ``world``: genes of every stage are dealt round-robin over the ranks of a
    torch.distributed job (one process per GPU); results are gathered on every rank because the
    stage schedule needs maxcor / sd.cv / lambda of the finished stage (R/saver_fit.R:158,244)."""

    def __init__(self, handle=None, comm=None, block=4096):
        from ._lib import default_handle
        self.handle = handle or default_handle()
        self.comm = comm
        self.block = int(block)

    def set_design(self, xest):
        ops.set_design(xest, handle=self.handle)

    def load_counts(self, x):
        """Sparse input (R: dgCMatrix): upload once, keep resident (saver_cuda_set_counts_csc)."""
        ops.set_counts_csc(x, handle=self.handle)
        return DeviceCounts(x, self.handle)

    def load_dense(self, x):
        """Dense input (R: matrix): upload once; clean.data's threshold, colSums / rowMeans and the
        log-normalised design are then computed on the device (saver_cuda_set_counts_dense)."""
        ops.set_counts_dense(x, handle=self.handle)
        return DenseDeviceCounts(x, self.handle)

    def calc_estimate(self, x, sf, scale_sf, mode, cutoff=0.0, coefs=None, self_col=None,
                      is_pred=None, pred_mask=None, foldid=None, nfolds=5, calc_maxcor=False,
                      estimates_only=False, want_mu=False, dfmax=300, nlambda=100, thresh=1e-7):
        """One block: x (B, n) raw counts.  Returns the 8-tuple of R/calc_estimate.R:159-160 as a
        dict (est, se, maxcor, lambda_max, lambda_min, sd_cv, ct, vt) plus status (and mu)."""
        return ops.calc_estimate(x, sf, scale_sf, mode, cutoff=cutoff, coefs=coefs,
                                 self_col=self_col, is_pred=is_pred, pred_mask=pred_mask,
                                 foldid=foldid, nfolds=nfolds, calc_maxcor=calc_maxcor,
                                 want_se=not estimates_only, want_mu=want_mu, dfmax=dfmax,
                                 nlambda=nlambda, thresh=thresh, handle=self.handle)


class _Sums:
    """Row / column sums of a DeviceCounts behind the ``x.sum(axis)`` the utils use."""

    def __init__(self, dc):
        self.dc, self.shape = dc, dc.shape

    def sum(self, axis):
        return self.dc.col_sums if axis == 0 else self.dc.row_sums


class SaverResult:
    """The ``saver`` S3 object (R/saver.R:173-175; fields R/saver_fit.R:66-70)."""

    def __init__(self, estimate, se, info, mu=None, gene_names=None, cell_names=None):
        self.estimate = estimate
        self.se = se
        self.info = info
        self.mu = mu
        self.gene_names = gene_names
        self.cell_names = cell_names


class LegacySaverResult:
    """A ``saver`` object in the layout of early SAVER releases, ``list(estimate, alpha, beta,
    info)``: the posterior of a value is Gamma(shape = alpha, rate = beta).  The reference still
    accepts such objects in sample.saver, cor.genes, cor.cells and combine.saver
    (R/sample_saver.R:34-36, R/cor_adjust.R:27-29,49-51, R/combine_saver.R:26-28 dispatching to
    R/utils.R:137-232); so do the functions below, on the same device kernels."""

    def __init__(self, estimate, alpha, beta, info, gene_names=None, cell_names=None):
        self.estimate = estimate
        self.alpha = alpha
        self.beta = beta
        self.info = info
        self.gene_names = gene_names
        self.cell_names = cell_names

    def moments(self, efficiency_known=False):
        """(mean, sd) panels that make the current kernels draw / shrink exactly as the legacy
        code does.  Gamma mode (rgamma(alpha, beta), R/utils.R:155-156) and the correlation
        adjustment (mean(alpha/beta^2), R/utils.R:189-191): mean = alpha/beta, sd =
        sqrt(alpha)/beta, so shape = mean^2/sd^2 = alpha and rate = mean/sd^2 = beta.  NB mode
        (rnbinom(mu = estimate, size = alpha), R/utils.R:152-153): mean = estimate, sd =
        estimate/sqrt(alpha), so size = mean^2/sd^2 = alpha."""
        a = np.asarray(self.alpha, np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            if efficiency_known:
                est = np.asarray(self.estimate, np.float64)
                return est, est / np.sqrt(a)
            b = np.asarray(self.beta, np.float64)
            return a / b, np.sqrt(a) / b


def _is_legacy(x):
    return getattr(x, "alpha", None) is not None


class DeviceCounts:
    """A sparse count matrix resident on the device (csrc/sparse.cu).  Replaces the Matrix-package
    calls of the reference's sparse path: colSums/rowSums (R/utils.R:72,100-115, R/saver.R:156),
    as.matrix(x[block, ]) (R/calc_estimate.R:69) and the dense x.est (R/saver.R:157)."""

    def __init__(self, x, handle):
        self.shape = x.shape
        self.handle = handle
        self.col_sums, self.row_sums = ops.csc_sums(handle=handle)

    def rows(self, rows):
        return ops.csc_gather(rows, handle=self.handle)

    def log_rows(self, rows, sf):
        """log((x[rows, ] + 1) / sf), left on the device for saver_cuda_set_design."""
        return ops.csc_gather(rows, sf=sf, transform=True, device=True, handle=self.handle)


class DenseDeviceCounts:
    """A dense count matrix resident on the device (csrc/sparse.cu, dense part): the whole-matrix
    passes of saver() -- x[x < 0.001] <- 0, colSums, rowMeans, x.est (R/utils.R:11-19,72,
    R/saver.R:156-157) -- without a host copy of the matrix or of the design."""

    def __init__(self, x, handle):
        self.shape = x.shape
        self.handle = handle
        self.col_sums, self.row_sums = ops.dense_sums(handle=handle)

    def log_rows(self, rows, sf):
        return ops.dense_gather(rows, sf=sf, transform=True, device=True, handle=self.handle)

    def release(self):
        ops.release_counts(handle=self.handle)


def _dense_rows(x, rows, dc=None):
    if dc is not None:
        return dc.rows(rows)
    if hasattr(x, "tocsc"):
        return np.asarray(x.tocsr()[rows].todense(), dtype=np.float64)
    xb = np.array(x[rows], dtype=np.float64)   # a copy: clean.data's threshold (a no-op on a cleaned x)
    xb[xb < 0.001] = 0
    return xb


class _Fit:
    """State shared by the stages of saver.fit (est / se / info assembly, R/saver_fit.R:60-70)."""

    def __init__(self, x, sf, scale_sf, ngenes, ncells, estimates_only, backend, keep_mu, comm,
                 dc=None):
        self.x, self.sf, self.scale_sf, self.dc = x, sf, scale_sf, dc
        self.backend, self.comm = backend, comm
        self.estimates_only = estimates_only
        self.est = np.zeros((ngenes, ncells))
        self.se = None if estimates_only else np.zeros((ngenes, ncells))
        self.mu = np.zeros((ngenes, ncells)) if keep_mu else None
        z = lambda: np.zeros(ngenes)  # noqa: E731
        self.info = {"size.factor": scale_sf * sf, "maxcor": z(), "lambda.max": z(),
                     "lambda.min": z(), "sd.cv": z(), "pred.time": z(), "var.time": z(),
                     "cutoff": 0.0, "lambda.coefs": 0.0, "total.time": 0.0}
        self.status = np.zeros(ngenes, dtype=np.int32)
        self.owner = np.zeros(ngenes, dtype=np.int32)
        self.seed_of = np.zeros(ngenes, dtype=np.int64)   # set.seed(j) of each cross-validated gene

    def gather(self):
        """Every rank receives the rows the other ranks computed (a no-op on one process)."""
        if self.comm and self.comm.world > 1:
            self.est = self.comm.gather_rows(self.est, self.owner)
            if self.se is not None:
                self.se = self.comm.gather_rows(self.se, self.owner)
            if self.mu is not None:
                self.mu = self.comm.gather_rows(self.mu, self.owner)

    def run(self, ind, mode, cutoff, coefs, calc_maxcor, self_col_of, is_pred_of, pred_mask,
            npred_cells, nfolds, fold_seed, fold_method, seeds=None):
        """calc.estimate(x[ind, ], ...) for one stage; genes are dealt over the ranks.  seeds: the
        cv seed j of every gene of ind (default: its 1-based position, R/calc_estimate.R:98)."""
        ind = np.asarray(ind, dtype=np.int64)
        seeds = np.arange(1, ind.size + 1) if seeds is None else np.asarray(seeds, dtype=np.int64)
        if ind.size == 0:
            return
        rank, world = (self.comm.rank, self.comm.world) if self.comm else (0, 1)
        pos = np.arange(ind.size)
        mine = pos[rank::world]
        self.owner[ind] = pos % world
        keys = ("maxcor", "lambda.max", "lambda.min", "sd.cv", "pred.time", "var.time")
        loc = {k: np.zeros(mine.size) for k in keys}
        loc_status = np.zeros(mine.size, dtype=np.int32)
        blk = self.backend.block
        for b0 in range(0, mine.size, blk):
            p = mine[b0:b0 + blk]
            g = ind[p]
            xb = _dense_rows(self.x, g, self.dc)
            fold = None
            if mode == 1:
                # seed j = 1-based position of the gene inside the stage block (R/calc_estimate.R:98)
                fold = np.zeros((g.size, xb.shape[1]), dtype=np.int32)
                pc = np.where(pred_mask)[0] if pred_mask is not None else np.arange(xb.shape[1])
                fold[:, pc] = draw_folds(seeds[p], npred_cells, nfolds, fold_seed, fold_method)
                self.seed_of[g] = seeds[p]
            r = self.backend.calc_estimate(
                xb, self.sf, self.scale_sf, mode, cutoff=cutoff, coefs=coefs,
                self_col=self_col_of[g], is_pred=is_pred_of[g], pred_mask=pred_mask, foldid=fold,
                nfolds=nfolds, calc_maxcor=calc_maxcor, estimates_only=self.estimates_only,
                want_mu=self.mu is not None)
            self.est[g] = r["est"]
            if self.se is not None:
                self.se[g] = r["se"]
            if self.mu is not None:
                self.mu[g] = r["mu"]
            sl = slice(b0, b0 + p.size)
            loc["maxcor"][sl] = r["maxcor"]
            loc["lambda.max"][sl] = r["lambda_max"]
            loc["lambda.min"][sl] = r["lambda_min"]
            loc["sd.cv"][sl] = r["sd_cv"]
            loc["pred.time"][sl] = r["ct"]
            loc["var.time"][sl] = r["vt"]
            loc_status[sl] = r["status"]
        if self.comm and world > 1:
            # per-gene scalars of the stage are needed by every rank (lm fits of the schedule)
            packed = np.stack([loc[k] for k in keys] + [loc_status.astype(np.float64)], 0)
            parts = self.comm.allgather_ragged(packed)
            for r_, part in enumerate(parts):
                g = ind[pos[r_::world]]
                for i, k in enumerate(keys):
                    self.info[k][g] = part[i]
                self.status[g] = part[len(keys)].astype(np.int32)
        else:
            g = ind[mine]
            for k in keys:
                self.info[k][g] = loc[k]
            self.status[g] = loc_status


def saver_fit(x, x_est, do_fast, ncores, sf, scale_sf, pred_genes, pred_cells, null_model,
              ngenes=None, ncells=None, gene_names=None, cell_names=None, estimates_only=False,
              *, self_col=None, backend=None, comm=None, perm=None, fold_seed=0,
              fold_method="philox", nfolds=5, keep_mu=False, messages=None, gather=True,
              counts=None, root_design=False, row_sums=None):
    """saver.fit (R/saver_fit.R:56-463): the stage schedule.

    x (genes x cells counts), x_est ((p, n) = t(x.est) of R/saver.R:157, gene-major),
    pred_genes / pred_cells 0-based.  self_col[g]: the x_est column carrying gene g's own name or
    -1 (R/calc_estimate.R:103 compares names; indices here).
    perm: the reference always permutes the genes with the ambient RNG (R/saver_fit.R:87-92); the
    order decides which genes feed the two schedule regressions of do.fast and the per-gene seeds.
    None (default) = a random permutation seeded with fold_seed; an int = that seed; ("R", seed) =
    the order R itself draws after set.seed(seed); "identity" = matrix order, an explicit opt-in for
    tests and benchmarks only (real matrices are often sorted, which would bias the schedule)."""
    if perm is None:
        perm = int(fold_seed)
    msg = messages or (lambda *a: None)
    ngenes = x.shape[0] if ngenes is None else ngenes
    ncells = x.shape[1] if ncells is None else ncells
    backend = backend or CudaBackend(comm=comm)
    nworkers = max(int(ncores), 1)
    fit = _Fit(x, sf, scale_sf, ngenes, ncells, estimates_only, backend, keep_mu, comm, counts)
    info = fit.info
    msg("Running SAVER with %d worker(s)" % nworkers)
    pred_genes = np.asarray(pred_genes, dtype=np.int64)
    if counts is not None:
        rs = counts.row_sums[pred_genes]
    elif row_sums is not None:
        rs = np.asarray(row_sums)[pred_genes]
    else:
        rs = np.asarray(x[pred_genes].sum(1)).ravel() if pred_genes.size else np.zeros(0)
    pred1 = pred_genes[rs > 0]
    npred = pred_genes.size
    if not null_model:
        if x_est is not None:
            msg("Calculating predictions for %d genes using %d genes and %d cells..."
                % (npred, x_est.shape[0], x_est.shape[1]))
        if comm is not None and comm.world > 1 and (root_design or comm.backend == "nccl"):
            # the design is uploaded once by rank 0 and broadcast (NCCL over NVLink on a GPU box);
            # the other ranks never derive it (x_est is None there in a root-only call)
            backend.set_design(comm.broadcast_array(x_est if comm.rank == 0 else None,
                                                    keep_on_device=comm.backend == "nccl"))
        else:
            backend.set_design(x_est)
    else:
        msg("Using means as predictions.")
    st = time.time()
    rest = np.setdiff1d(np.arange(ngenes), pred1, assume_unique=False)
    ind = gene_order(perm, pred1, rest, ngenes)
    self_col_of = np.full(ngenes, -1, dtype=np.int32)
    if self_col is not None:
        self_col_of[:] = np.asarray(self_col, dtype=np.int32)
    is_pred_of = np.zeros(ngenes, dtype=np.int32)
    is_pred_of[pred_genes] = 1
    pred_mask = None
    pc = np.asarray(pred_cells, dtype=np.int64)
    if pc.size != ncells or (pc != np.arange(ncells)).any():
        pred_mask = np.zeros(ncells, dtype=np.int32)
        pred_mask[pc] = 1
    common = dict(self_col_of=self_col_of, is_pred_of=is_pred_of, pred_mask=pred_mask,
                  npred_cells=np.unique(pc).size, nfolds=nfolds, fold_seed=fold_seed,
                  fold_method=fold_method)

    def finish(done):
        if done != ngenes:
            ind6 = ind[npred:ngenes]
            msg("Estimating remaining %d genes." % ind6.size)
            fit.run(ind6, 0, 0.0, None, False, **common)
        if gather:
            fit.gather()
        info["total.time"] = time.time() - st
        return dict(estimate=fit.est, se=fit.se, info=info, mu=fit.mu, status=fit.status,
                    seed_of=fit.seed_of, owner=fit.owner)

    if npred == 0:
        return finish(0)
    if do_fast and not null_model:
        n1 = min(max(8, nworkers), npred)
        msg("Estimating finish time...")
        fit.run(ind[:n1], 1, 0.0, None, True, **common)
        n2 = min(max(100, nworkers), npred)
        if n1 == npred:
            return finish(n1)
        msg("Calculating max cor cutoff...")
        fit.run(ind[n1:n2], 1, 0.0, None, True, **common)
        b = _lm(np.sqrt(info["sd.cv"][ind[:n2]]), info["maxcor"][ind[:n2]])
        cutoff = float((0.5 - b[0]) / b[1])
        info["cutoff"] = cutoff
        perc = float(np.mean(info["maxcor"][ind[:n2]] > cutoff))
        n3 = npred if perc <= 0 else min(max(int(np.ceil(n2 / perc)), nworkers) + n2, npred)
        if n2 == npred:
            return finish(n2)
        msg("Calculating lambda coefficients...")
        fit.run(ind[n2:n3], 1, cutoff, None, True, **common)
        if n3 == npred:
            return finish(n3)
        pred = np.where(info["maxcor"] > cutoff)[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            z = np.log(info["lambda.max"][pred] / info["lambda.min"][pred]) ** 2
        coefs = _lm(z, info["maxcor"][pred])
        info["lambda.coefs"] = coefs
        cutoff2 = max(cutoff, float(-coefs[0] / coefs[1]))
        # The reference predicts a quarter of the rest first only to print a finish-time estimate
        # (R/saver_fit.R:262-330); the fixed-lambda fits carry no seed and do not depend on the
        # stage they run in, so both parts go down as one block: one pass over the device, and the
        # few genes that crawl to glmnet's maxit hold one stage instead of two.
        msg("Predicting remaining genes...")
        fit.run(ind[n3:npred], 2, cutoff2, coefs, True, **common)
        return finish(npred)
    mode = 0 if null_model else 1
    # The slow schedule: n1, a quarter of the rest, the rest -- all cross-validated, maxcor = 2
    # (R/saver_fit.R:333-462).  The reference runs them as three calc.estimate calls only to print a
    # finish-time estimate in between; a gene's result depends on its stage through nothing but its
    # cv seed j = position inside the stage (R/calc_estimate.R:98).  So the three stages go down as
    # ONE device pass with exactly those seeds: same folds, same results, and the 8-gene first stage
    # does not cost a whole latency-bound pass of its own.
    n1 = min(max(8, nworkers), npred)
    n2 = min(int(np.ceil((npred - n1) / 4)) + n1, npred)
    seeds = np.concatenate([np.arange(1, b - a + 1) for a, b in ((0, n1), (n1, n2), (n2, npred))])
    msg("Predicting %d genes (stages of %d, %d, %d)..." % (npred, n1, n2 - n1, npred - n2))
    fit.run(ind[:npred], mode, 0.0, None, False, seeds=seeds, **common)
    return finish(npred)


def _fit_given_mu(x, sf, scale_sf, mu_rows, estimates_only, backend, comm, block, dc=None):
    """Shared body of saver.fit.mean / saver.fit.null: calc.post on a supplied prediction."""
    ngenes, ncells = x.shape
    est = np.zeros((ngenes, ncells))
    se = None if estimates_only else np.zeros((ngenes, ncells))
    z = lambda: np.zeros(ngenes)  # noqa: E731
    info = {"size.factor": scale_sf * sf, "maxcor": z(), "lambda.max": z(), "lambda.min": z(),
            "sd.cv": z(), "pred.time": z(), "var.time": z(), "cutoff": 0.0, "lambda.coefs": 0.0,
            "total.time": 0.0}
    st = time.time()
    h = backend.handle
    for b0 in range(0, ngenes, block):
        rows = np.arange(b0, min(ngenes, b0 + block))
        xb = _dense_rows(x, rows, dc)
        t0 = time.time()
        r = ops.calc_post(xb, mu_rows(rows, xb), sf, scale_sf, want_se=not estimates_only,
                          want_info=False, handle=h)
        est[rows] = r["est"]
        if se is not None:
            se[rows] = r["se"]
        info["var.time"][rows] = (time.time() - t0) / rows.size
    info["total.time"] = time.time() - st
    return dict(estimate=est, se=se, info=info, mu=None, status=np.zeros(ngenes, dtype=np.int32))


def saver_fit_mean(x, ncores, sf, scale_sf, mu, estimates_only=False, backend=None, comm=None,
                   counts=None):
    """saver.fit.mean (R/saver_fit.R:467-538) + calc.estimate.mean (R/calc_estimate.R:166-221):
    the user's mu, rescaled per gene to the mean of y, zero rows patched with the gene mean."""
    backend = backend or CudaBackend(comm=comm)

    def mu_rows(rows, xb):
        y = xb / sf
        imu = np.array(mu[rows], dtype=np.float64, copy=True)
        gm, mm = y.mean(1), imu.mean(1)
        imu[mm == 0] = gm[mm == 0][:, None]
        return imu * (y.mean(1) / imu.mean(1))[:, None]

    return _fit_given_mu(x, sf, scale_sf, mu_rows, estimates_only, backend, comm, backend.block,
                         counts)


def saver_fit_null(x, ncores, sf, scale_sf, estimates_only=False, backend=None, comm=None,
                   counts=None):
    """saver.fit.null (R/saver_fit.R:542-628) + calc.estimate.null (R/calc_estimate.R:226-277)."""
    backend = backend or CudaBackend(comm=comm)
    return _fit_given_mu(x, sf, scale_sf, lambda rows, xb: (xb / sf).mean(1), estimates_only,
                         backend, comm, backend.block, counts)


def saver(x, do_fast=True, ncores=1, size_factor=None, npred=None, pred_cells=None,
          pred_genes=None, pred_genes_only=False, null_model=False, mu=None,
          estimates_only=False, *, gene_names=None, cell_names=None, backend=None, comm=None,
          perm=None, fold_seed=0, fold_method="philox", nfolds=5, return_mu=False,
          verbose=False, gather=True, device_prep=True):
    """saver() (R/saver.R:98-179).  x: genes x cells counts (numpy array or scipy.sparse).

    pred_cells / pred_genes are 0-based.  ``ncores`` is accepted for signature compatibility;
    device parallelism (one process per GPU, ``comm``) replaces the PSOCK workers.  Returns a
    SaverResult (or the estimate matrix when ``estimates_only``), exactly as the reference.

    Multi-GPU (``comm``): every stage's genes are dealt over the ranks.  Either every rank passes
    the same x, or only rank 0 does and the others pass ``x=None`` (root-only: rank 0 prepares the
    design, one broadcast ships it).  ``gather=False`` leaves each rank with the rows it computed
    (``result.owner`` tells which) instead of all-gathering estimate / se to every rank -- the
    sink is host memory, so no collective is needed for the results."""
    msg = (lambda s: print(s, flush=True)) if verbose else (lambda s: None)
    world = comm.world if comm is not None else 1
    # Root-only call (one process per GPU, only rank 0 holds the matrix -- the reference's own
    # parallel contract: one x, chunks exported to the workers, R/calc_estimate.R:69-74): rank 0
    # does the host preparation once, the design goes out with one broadcast, the others receive
    # the plan and the counts of the genes to process.  x = None on a rank selects it.
    root_only = world > 1 and comm.max_scalar(1.0 if x is None else 0.0) > 0
    if root_only and comm.rank != 0:
        plan = comm.bcast_obj(None)
        if plan.get("error"):
            raise SaverError(plan["error"])
        x = comm.broadcast_array(None)
        backend = backend or CudaBackend(comm=comm)
        out = saver_fit(x, None, do_fast, ncores, plan["sf"], plan["scale_sf"], plan["pg"],
                        plan["pc"], null_model, estimates_only=estimates_only,
                        self_col=plan["self_col"], backend=backend, comm=comm, perm=perm,
                        fold_seed=fold_seed, fold_method=fold_method, nfolds=nfolds,
                        keep_mu=return_mu, messages=msg, gather=gather, root_design=True)
        gene_names, cell_names = plan["gene_names"], plan["cell_names"]
        sf, scale_sf = plan["sf"], plan["scale_sf"]
        return _saver_result(out, x, sf, scale_sf, gene_names, cell_names, estimates_only,
                             return_mu, gather)
    try:
        if root_only and (mu is not None or hasattr(x, "tocsc")):
            raise SaverError("a root-only saver() call (x = None on the other ranks) takes a dense "
                             "matrix and no mu; pass x on every rank instead")
        if mu is not None:
            mu = check_mu(x, mu)
        backend = backend or CudaBackend(comm=comm)
        # Dense matrices of benchmark size: the whole-matrix passes (threshold, colSums, rowMeans, the
        # log-normalised design) run on the device from one upload of x; the design never exists on
        # the host.  clean.data also drops all-zero cells -- if there is one, the host path below runs.
        ddc = None
        if (mu is None and not null_model and not hasattr(x, "tocsc") and hasattr(backend, "load_dense")
                and device_prep and np.ndim(x) == 2 and np.size(x) >= (1 << 20)):
            xa = np.asarray(x)
            if xa.shape[1] <= 1:
                raise SaverError("x should be a matrix with 2 or more columns")
            if xa.dtype != np.int32:
                xa = np.ascontiguousarray(xa, dtype=np.float64)
            ddc = backend.load_dense(np.ascontiguousarray(xa))
            if (ddc.col_sums == 0).any():
                ddc.release()
                ddc = None
        if ddc is None:
            x, kept = clean_data(x)
        else:
            x, kept = xa, np.ones(xa.shape[1], dtype=bool)
    except SaverError as e:
        if root_only:
            comm.bcast_obj({"error": str(e)})
        raise
    if cell_names is not None:
        cell_names = np.asarray(cell_names)[kept]
    ngenes, ncells = x.shape
    if gene_names is None:
        gene_names = np.arange(1, ngenes + 1)
    msg("%d genes, %d cells" % (ngenes, ncells))
    # sparse input stays sparse on the device; blocks are densified there (csrc/sparse.cu)
    dc = backend.load_counts(x) if hasattr(x, "tocsc") and hasattr(backend, "load_counts") else None
    sums = _Sums(dc) if dc is not None else _Sums(ddc) if ddc is not None else x
    sf, scale_sf = calc_size_factor(sums, size_factor, ncells)
    if mu is not None:
        out = saver_fit_mean(x, ncores, sf, scale_sf, mu, estimates_only, backend, comm, dc)
    elif null_model and not root_only:
        out = saver_fit_null(x, ncores, sf, scale_sf, estimates_only, backend, comm, dc)
    else:
        pc = get_pred_cells(pred_cells, ncells)
        pg = get_pred_genes(sums, pred_genes, npred, ngenes)
        row_sums = None if sums is x else sums.dc.row_sums
        rm = np.asarray(x.mean(1)).ravel() if row_sums is None else row_sums / ncells
        good = np.where(rm >= 0.1)[0]
        # x.est = t(log((x[good, ] + 1) / sf))  (R/saver.R:156-157), kept gene-major: (p, n)
        if null_model:
            x_est = None
        elif dc is not None:
            x_est = dc.log_rows(good, sf)
        elif ddc is not None:
            x_est = ddc.log_rows(good, sf)   # a device panel: goes into set_design / the broadcast as it is
            ddc.release()
        else:
            x_est = np.log((_dense_rows(x, good) + 1.0) / sf)
        # the x.est column carrying each gene's own name (R/calc_estimate.R:103), -1 if none
        self_col = np.full(ngenes, -1, dtype=np.int32)
        self_col[good] = np.arange(good.size, dtype=np.int32)
        if pred_genes_only:
            if dc is not None:       # row subset of a resident sparse matrix: re-upload the slice
                x = x.tocsr()[pg].tocsc()
                dc = backend.load_counts(x)
            else:
                x = x[pg]
            if row_sums is not None:
                row_sums = row_sums[pg]
            gene_names = np.asarray(gene_names)[pg]
            self_col = self_col[pg]
            pg = np.arange(x.shape[0])
        if root_only:
            comm.bcast_obj(dict(sf=sf, scale_sf=scale_sf, pg=pg, pc=pc, self_col=self_col,
                                gene_names=gene_names, cell_names=cell_names))
            comm.broadcast_array(x)   # counts of the genes to process (the whole x, or x[pred.genes, ])
        out = saver_fit(x, x_est, do_fast, ncores, sf, scale_sf, pg, pc, null_model,
                        estimates_only=estimates_only, self_col=self_col, backend=backend,
                        comm=comm, perm=perm, fold_seed=fold_seed, fold_method=fold_method,
                        nfolds=nfolds, keep_mu=return_mu, messages=msg, counts=dc, gather=gather,
                        root_design=root_only, row_sums=row_sums)
    msg("Done!")
    return _saver_result(out, x, sf, scale_sf, gene_names, cell_names, estimates_only, return_mu,
                         gather)


def _saver_result(out, x, sf, scale_sf, gene_names, cell_names, estimates_only, return_mu, gather):
    res = SaverResult(out["estimate"], out["se"], out["info"], out.get("mu"), gene_names,
                      cell_names)
    res.status = out.get("status")
    res.seed_of = out.get("seed_of")
    res.owner = out.get("owner")   # rank that computed each row (rows of other ranks are only
    res.x, res.sf, res.scale_sf = x, sf, scale_sf   # filled in when gather=True)
    if estimates_only and not return_mu and gather:
        return res.estimate
    return res


# ----------------------------------------------------------------------------------------------
# consumers of a saver object
# ----------------------------------------------------------------------------------------------
def _row_blocks(count, world):
    """Contiguous near-equal blocks [b0, b1) of `count` rows, one per rank, and the owner vector."""
    edges = [(count * r) // world for r in range(world + 1)]
    owner = np.zeros(count, dtype=np.int64)
    for r in range(world):
        owner[edges[r]:edges[r + 1]] = r
    return edges, owner


def sample_saver(x, rep=1, efficiency_known=False, seed=None, handle=None, block=4096, comm=None):
    """sample.saver (R/sample_saver.R:32-79).  x: SaverResult.  Returns one (G, n) array, or a
    list of ``rep`` arrays.  seed=None draws a fresh seed (R: the ambient RNG state).

    With ``comm`` (one process per GPU) every rank draws a contiguous block of genes and the rows
    are gathered; the generator is counter based on (seed, rep, gene, cell), so the dataset does
    not depend on the number of ranks or on ``block``."""
    rep = int(rep)
    if rep <= 0:
        raise SaverError("rep must be a positive integer.")
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(1)[0])
    world = comm.world if comm is not None else 1
    if world > 1:
        seed = int(comm.max_scalar(seed))  # all ranks agree on one seed
    if _is_legacy(x):  # sample.saver.old (R/utils.R:137-178)
        est, se = x.moments(efficiency_known)
    else:
        est, se = np.asarray(x.estimate), np.asarray(x.se)
    G = est.shape[0]
    edges, owner = _row_blocks(G, world)
    g0, g1 = (edges[comm.rank], edges[comm.rank + 1]) if world > 1 else (0, G)
    outs = []
    for j in range(rep):
        out = np.zeros_like(est) if world > 1 else np.empty_like(est)
        for b0 in range(g0, g1, block):
            sl = slice(b0, min(b0 + block, g1))
            out[sl] = ops.sample(est[sl], se[sl], nb_mode=efficiency_known, seed=seed, rep_idx=j,
                                 gene0=b0, handle=handle)
        if world > 1:
            out = comm.gather_rows(out, owner)
        outs.append(out)
    return outs[0] if rep == 1 else outs


def _posterior_sd(x):
    """se of a current object; sqrt(alpha)/beta for a legacy one, whose adjustment uses
    mean(alpha/beta^2) in place of mean(se^2) (cor.genes.old / cor.cells.old, R/utils.R:181-214)."""
    return x.moments(False)[1] if _is_legacy(x) else np.asarray(x.se)


def _cor_adjust(x, by_cells, cor_mat, handle, comm, gather):
    est, sd = np.asarray(x.estimate), _posterior_sd(x)
    world = comm.world if comm is not None else 1
    if world == 1:
        return ops.cor_adjust(est, sd, by_cells=by_cells, cor_in=cor_mat, handle=handle)
    # every rank holds est / se and computes a block of columns of the symmetric result
    # (cor.genes shards by column panel; the inputs are replicated)
    M = est.shape[1] if by_cells else est.shape[0]
    edges, owner = _row_blocks(M, world)
    c0, c1 = edges[comm.rank], edges[comm.rank + 1]
    cin = None if cor_mat is None else np.ascontiguousarray(np.asarray(cor_mat)[c0:c1])
    blk = (ops.cor_adjust(est, sd, by_cells=by_cells, cor_in=cin, col0=c0, ncols=c1 - c0,
                          handle=handle) if c1 > c0 else np.zeros((0, M)))
    if not gather:
        return blk, (c0, c1)
    full = np.zeros((M, M))
    full[c0:c1] = blk
    return comm.gather_rows(full, owner)


def cor_genes(x, cor_mat=None, handle=None, comm=None, gather=True):
    """cor.genes (R/cor_adjust.R:26-44): gene-gene correlation of the estimates, shrunk by the
    posterior uncertainty.  The diagonal is adj^2, not 1, exactly as in the reference.

    With ``comm`` each rank computes a contiguous block of columns; ``gather=False`` returns
    ``(block, (c0, c1))`` -- rows c0..c1 of the symmetric matrix -- so that every rank can write
    its own slice to the sink instead of materialising G x G everywhere."""
    return _cor_adjust(x, False, cor_mat, handle, comm, gather)


def cyclic_window(rank, world):
    """Row blocks rank `rank` computes of its column panel of a symmetric matrix split into
    world x world blocks: its own block and the next world // 2 in cyclic order (for an even world
    the farthest block only on the lower half of the ranks) -- every unordered pair of blocks is
    computed exactly once and every rank does the same amount of work to within one block."""
    half = world // 2
    blocks = [(rank + d) % world for d in range(half + 1)]
    if world % 2 == 0 and world > 1 and rank >= half:
        blocks = blocks[:-1]
    return blocks


def cor_genes_sharded(est, se, comm, handle=None, symmetric=True):
    """cor.genes (R/cor_adjust.R:26-44) on a gene-sharded saver object: rank r holds the rows
    (est, se: (B_r, n) CUDA tensors or arrays) of ITS genes, blocks stacked in rank order.

    Every rank standardises its own rows, ONE all-gather ships the standardised panel and adj, and
    the rank fills its column panel of the adjusted correlation -- with ``symmetric`` only the row
    blocks of cyclic_window(rank, world), the rest being the transposes of blocks other ranks own.
    Returns (panel, (c0, c1), blocks): panel (B_r, G) CUDA tensor = rows c0..c1 of the matrix,
    blocks = [(r0, r1)] column ranges of the panel that were computed."""
    import torch
    dev = comm.device
    on_gpu = dev.type == "cuda"
    h = (handle or ops.default_handle()) if on_gpu else handle

    def sync():   # the library runs on its own stream, the collectives on torch's
        if on_gpu:
            h.synchronize()
            torch.cuda.synchronize()

    est = est if hasattr(est, "is_cuda") else torch.from_numpy(np.ascontiguousarray(est))
    se = se if hasattr(se, "is_cuda") else torch.from_numpy(np.ascontiguousarray(se))
    est, se = est.to(dev).contiguous(), se.to(dev).contiguous()
    sizes = torch.zeros(comm.world, dtype=torch.int64, device=dev)
    sizes[comm.rank] = est.shape[0]
    comm.dist.all_reduce(sizes)
    sizes = [int(v) for v in sizes.cpu().numpy()]
    edges = np.concatenate([[0], np.cumsum(sizes)])
    G = int(edges[-1])
    Z, adj = ops.cor_prep(est, se, handle=h)
    sync()
    Zall = comm.allgather_blocks(Z, sizes)
    adj_all = comm.allgather_blocks(adj, sizes)
    Zall = torch.cat([Zall, torch.zeros((128, Zall.shape[1]), dtype=Zall.dtype, device=dev)], 0)
    c0, c1 = int(edges[comm.rank]), int(edges[comm.rank + 1])
    panel = torch.zeros((c1 - c0, G), dtype=torch.float64, device=dev)
    todo = cyclic_window(comm.rank, comm.world) if symmetric else list(range(comm.world))
    blocks = []
    sync()
    for b in todo:
        r0, r1 = int(edges[b]), int(edges[b + 1])
        if r1 > r0 and c1 > c0:
            ops.cor_block(Zall, adj_all, G, r0, r1 - r0, c0, c1 - c0, panel, handle=h)
            blocks.append((r0, r1))
    sync()
    return panel, (c0, c1), blocks


def cor_cells(x, cor_mat=None, handle=None, comm=None, gather=True):
    """cor.cells (R/cor_adjust.R:48-66); ``comm`` / ``gather`` as in cor_genes."""
    return _cor_adjust(x, True, cor_mat, handle, comm, gather)


def get_mu(x, saver_obj):
    """get.mu (R/get_mu.R:22-25): the prior mean behind a saver object (host arithmetic)."""
    est, se = np.asarray(saver_obj.estimate), np.asarray(saver_obj.se)
    with np.errstate(divide="ignore", invalid="ignore"):
        mu = (est ** 2 / se ** 2 - np.asarray(x)) / (est / se ** 2 - saver_obj.info["size.factor"])
    return np.round(mu, 3)


def combine_saver(saver_list):
    """combine.saver (R/combine_saver.R:25-44): stack objects fitted on gene subsets."""
    if _is_legacy(saver_list[0]):
        # combine.saver.old (R/utils.R:217-232): rbind estimate/alpha/beta; info[[1]] (size.factor)
        # from the first object, info[[2]] and info[[3]] concatenated, the rest from the first
        first = saver_list[0].info
        info = {}
        for i, k in enumerate(first):
            info[k] = (np.concatenate([np.atleast_1d(s.info[k]) for s in saver_list])
                       if i in (1, 2) else first[k])
        return LegacySaverResult(np.concatenate([s.estimate for s in saver_list], 0),
                                 np.concatenate([s.alpha for s in saver_list], 0),
                                 np.concatenate([s.beta for s in saver_list], 0), info,
                                 gene_names=_cat_names([getattr(s, "gene_names", None) for s in saver_list]),
                                 cell_names=getattr(saver_list[0], "cell_names", None))
    est = np.concatenate([s.estimate for s in saver_list], 0)
    se = np.concatenate([s.se for s in saver_list], 0)
    info = {"size.factor": saver_list[0].info["size.factor"]}
    for k in ("maxcor", "lambda.max", "lambda.min", "sd.cv", "pred.time", "var.time", "cutoff",
              "lambda.coefs"):
        info[k] = np.concatenate([np.atleast_1d(s.info[k]) for s in saver_list])
    info["total.time"] = float(sum(s.info["total.time"] for s in saver_list))
    return SaverResult(est, se, info, gene_names=_cat_names([getattr(s, "gene_names", None) for s in saver_list]),
                       cell_names=getattr(saver_list[0], "cell_names", None))


def _cat_names(parts):
    """rbind keeps the row names of its arguments (None if any part has none)."""
    if any(p is None for p in parts):
        return None
    return [n for p in parts for n in p]
