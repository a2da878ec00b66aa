/*
 * rcall_shim.c -- .Call glue between R and libsaver_cuda (include/saver_cuda.h).
 *
 * Compiled only where R exists (R CMD SHLIB, needs Rinternals.h); it cannot be built in the
 * image this repo is developed in (no R).  Pure marshalling: every numerical decision lives behind
 * the header-free C ABI.  R matrices are column-major; a block of genes travels as t(x[block, ])
 * (n x B, one gene per column), which is the panel layout of the ABI.
 * Never longjmp (Rf_error) across a live CUDA frame: errors are turned into R conditions only
 * after the library call has returned.
 */
#include <R.h>
#include <Rinternals.h>

#include "saver_cuda.h"

static void handle_finalizer(SEXP ptr)
{
    saver_cuda_handle h = (saver_cuda_handle)R_ExternalPtrAddr(ptr);
    if (h) saver_cuda_destroy(h);
    R_ClearExternalPtr(ptr);
}

static saver_cuda_handle get_handle(SEXP ptr)
{
    saver_cuda_handle h = (saver_cuda_handle)R_ExternalPtrAddr(ptr);
    if (!h) Rf_error("libsaver_cuda handle is closed");
    return h;
}

static void check(saver_cuda_handle h, int rc)
{
    if (rc != SAVER_OK) Rf_error("libsaver_cuda error %d: %s", rc, saver_cuda_last_error(h));
}

/* Argument checks run BEFORE the library call: REAL()/INTEGER() on a value of another type, or a
 * vector shorter than the library will read, must become an R error, never an out-of-bounds read
 * (len < 0: any length; NULL is accepted only where the caller tests Rf_isNull first). */
static double *real_arg(SEXP x, R_xlen_t len, const char *what)
{
    if (!Rf_isReal(x)) Rf_error("libsaver_cuda: %s must be a double vector / matrix (use storage.mode<-)", what);
    if (len >= 0 && (R_xlen_t)Rf_length(x) != len)
        Rf_error("libsaver_cuda: %s has length %d, expected %ld", what, Rf_length(x), (long)len);
    return REAL(x);
}

static int *int_arg(SEXP x, R_xlen_t len, const char *what)
{
    if (!Rf_isInteger(x)) Rf_error("libsaver_cuda: %s must be an integer vector", what);
    if (len >= 0 && (R_xlen_t)Rf_length(x) != len)
        Rf_error("libsaver_cuda: %s has length %d, expected %ld", what, Rf_length(x), (long)len);
    return INTEGER(x);
}

/* device: one CUDA ordinal, or several -- then ONE R process drives all of them the way ncores =
 * drives PSOCK workers (R/saver.R:120-138): saver_cuda_create_multi */
SEXP C_saver_cuda_create(SEXP device)
{
    saver_cuda_handle h = NULL;
    int rc;
    if (Rf_isInteger(device) && Rf_length(device) > 1)
        rc = saver_cuda_create_multi(Rf_length(device), INTEGER(device), &h);
    else
        rc = saver_cuda_create(Rf_asInteger(device), &h);
    if (rc != SAVER_OK) Rf_error("no usable CUDA device (libsaver_cuda has no CPU fallback), code %d", rc);
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, handle_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* x.est: n x p numeric matrix (R/saver.R:157) */
SEXP C_saver_cuda_set_design(SEXP hp, SEXP xest)
{
    saver_cuda_handle h = get_handle(hp);
    check(h, saver_cuda_set_design(h, real_arg(xest, (R_xlen_t)Rf_nrows(xest) * Rf_ncols(xest), "x.est"),
                                   Rf_nrows(xest), Rf_ncols(xest), 0));
    return R_NilValue;
}

/* the foreach body of calc.estimate (R/calc_estimate.R:75-141) for one block.
 * xt: n x B = t(x[block, ]); foldid: n x B integer or NULL; returns the 8-element list. */
SEXP C_saver_cuda_calc_estimate(SEXP hp, SEXP xt, SEXP sf, SEXP scale_sf, SEXP self_col,
                                SEXP is_pred, SEXP pred_mask, SEXP foldid, SEXP mode,
                                SEXP calc_maxcor, SEXP cutoff, SEXP coefs, SEXP estimates_only)
{
    saver_cuda_handle h = get_handle(hp);
    const int n = Rf_nrows(xt), B = Rf_ncols(xt), eo = Rf_asLogical(estimates_only);
    const R_xlen_t nb = (R_xlen_t)n * B;
    double *px = real_arg(xt, nb, "t(x)"), *psf = real_arg(sf, n, "sf");
    int *pself = int_arg(self_col, B, "self.col"), *ppred = int_arg(is_pred, B, "is.pred");
    int *pmask = Rf_isNull(pred_mask) ? NULL : int_arg(pred_mask, n, "pred.cells mask");
    int *pfold = Rf_isNull(foldid) ? NULL : int_arg(foldid, nb, "foldid");
    double *pcoefs = Rf_isNull(coefs) ? NULL : real_arg(coefs, 2, "coefs");
    SEXP est = PROTECT(Rf_allocMatrix(REALSXP, n, B));
    SEXP se = PROTECT(eo ? Rf_ScalarLogical(NA_LOGICAL) : Rf_allocMatrix(REALSXP, n, B));
    SEXP info = PROTECT(Rf_allocMatrix(REALSXP, B, 6));
    int rc = saver_cuda_calc_estimate(
        h, px, n, B, psf, Rf_asReal(scale_sf), pself, ppred, pmask, pfold,
        5, Rf_asInteger(mode), Rf_asLogical(calc_maxcor), Rf_asReal(cutoff),
        pcoefs, 300, 100, 1e-7, REAL(est), eo ? NULL : REAL(se),
        REAL(info), NULL, NULL, 0);
    check(h, rc);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, est);   /* n x B: the R wrapper transposes into B x n */
    SET_VECTOR_ELT(out, 1, se);
    SET_VECTOR_ELT(out, 2, info);  /* columns: maxcor, lambda.max, lambda.min, sd.cv, ct, vt */
    UNPROTECT(4);
    return out;
}

/* calc.post(y, mu, sf, scale.sf) for one gene (R/calc_posterior.R:26) */
SEXP C_saver_cuda_calc_post(SEXP hp, SEXP y, SEXP mu, SEXP sf, SEXP scale_sf)
{
    saver_cuda_handle h = get_handle(hp);
    const int n = Rf_length(y);
    double *py = real_arg(y, n, "y"), *pmu = real_arg(mu, Rf_length(mu) == 1 ? 1 : n, "mu"),
           *psf = real_arg(sf, n, "sf");
    SEXP est = PROTECT(Rf_allocVector(REALSXP, n)), se = PROTECT(Rf_allocVector(REALSXP, n));
    check(h, saver_cuda_calc_post(h, py, pmu, Rf_length(mu) == 1, psf, n, 1,
                                  Rf_asReal(scale_sf), REAL(est), REAL(se), NULL, NULL, NULL, NULL, 0));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, est);
    SET_VECTOR_ELT(out, 1, se);
    UNPROTECT(3);
    return out;
}

/* calc.a / calc.b / calc.k and calc.loglik.* (R/optimize_variance.R, R/calc_loglik.R) */
SEXP C_saver_cuda_calc_var(SEXP hp, SEXP model, SEXP y, SEXP mu, SEXP sf)
{
    saver_cuda_handle h = get_handle(hp);
    const int n = Rf_length(y);
    double *py = real_arg(y, n, "y"), *pmu = real_arg(mu, Rf_length(mu) == 1 ? 1 : n, "mu"),
           *psf = real_arg(sf, Rf_length(sf) == 1 ? 1 : n, "sf");
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 2));
    int flags = 0;
    check(h, saver_cuda_calc_var(h, Rf_asInteger(model), py, pmu, Rf_length(mu), psf,
                                 Rf_length(sf), n, REAL(out), &flags, NULL));
    UNPROTECT(1);
    return out;
}

/* calc.a / calc.b / calc.k with the 100-point grid of the fallback branch returned as well, so the
 * R side can do the reference's own draw, mean(sample(samp, 10000, replace = TRUE, prob = .))
 * (R/optimize_variance.R:46-53), on R's RNG stream: list(c(param, nll), flags, grid[200]).
 * flags bit 0 = the grid branch was taken; grid = 100 points then their 100 log-likelihoods. */
SEXP C_saver_cuda_calc_var_grid(SEXP hp, SEXP model, SEXP y, SEXP mu, SEXP sf)
{
    saver_cuda_handle h = get_handle(hp);
    const int n = Rf_length(y);
    double *py = real_arg(y, n, "y"), *pmu = real_arg(mu, Rf_length(mu) == 1 ? 1 : n, "mu"),
           *psf = real_arg(sf, Rf_length(sf) == 1 ? 1 : n, "sf");
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP par = PROTECT(Rf_allocVector(REALSXP, 2));
    SEXP fl = PROTECT(Rf_allocVector(INTSXP, 1));
    SEXP grid = PROTECT(Rf_allocVector(REALSXP, 200));
    int flags = 0;
    check(h, saver_cuda_calc_var(h, Rf_asInteger(model), py, pmu, Rf_length(mu), psf,
                                 Rf_length(sf), n, REAL(par), &flags, REAL(grid)));
    INTEGER(fl)[0] = flags;
    SET_VECTOR_ELT(out, 0, par);
    SET_VECTOR_ELT(out, 1, fl);
    SET_VECTOR_ELT(out, 2, grid);
    UNPROTECT(4);
    return out;
}

SEXP C_saver_cuda_loglik(SEXP hp, SEXP model, SEXP param, SEXP y, SEXP mu, SEXP sf)
{
    saver_cuda_handle h = get_handle(hp);
    double v = 0;
    const int n = Rf_length(y);
    double *py = real_arg(y, n, "y"), *pmu = real_arg(mu, Rf_length(mu) == 1 ? 1 : n, "mu"),
           *psf = real_arg(sf, Rf_length(sf) == 1 ? 1 : n, "sf");
    check(h, saver_cuda_loglik(h, Rf_asInteger(model), Rf_asReal(param), py, pmu,
                               Rf_length(mu), psf, Rf_length(sf), n, &v));
    return Rf_ScalarReal(v);
}

/* calc.maxcor(x.est, t(y)) with the design already installed */
SEXP C_saver_cuda_maxcor(SEXP hp, SEXP yt, SEXP self_col)
{
    saver_cuda_handle h = get_handle(hp);
    const int B = Rf_ncols(yt);
    double *py = real_arg(yt, (R_xlen_t)Rf_nrows(yt) * B, "x2");
    int *pself = int_arg(self_col, B, "self.col");
    SEXP out = PROTECT(Rf_allocVector(REALSXP, B));
    check(h, saver_cuda_maxcor(h, py, B, pself, REAL(out), 0));
    UNPROTECT(1);
    return out;
}

/* sample.saver / cor.genes / cor.cells on t(estimate), t(se) (n x G) */
SEXP C_saver_cuda_sample(SEXP hp, SEXP est_t, SEXP se_t, SEXP nb, SEXP seed, SEXP rep)
{
    saver_cuda_handle h = get_handle(hp);
    const int n = Rf_nrows(est_t), G = Rf_ncols(est_t);
    double *pe = real_arg(est_t, (R_xlen_t)n * G, "t(estimate)"), *ps = real_arg(se_t, (R_xlen_t)n * G, "t(se)");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, G));
    check(h, saver_cuda_sample(h, pe, ps, n, G, 0, Rf_asLogical(nb),
                               (unsigned long long)Rf_asReal(seed), Rf_asInteger(rep), REAL(out), 0));
    UNPROTECT(1);
    return out;
}

SEXP C_saver_cuda_cor_adjust(SEXP hp, SEXP est_t, SEXP se_t, SEXP by_cells, SEXP cor_mat)
{
    saver_cuda_handle h = get_handle(hp);
    const int n = Rf_nrows(est_t), G = Rf_ncols(est_t), bc = Rf_asLogical(by_cells);
    const int M = bc ? n : G;
    double *pe = real_arg(est_t, (R_xlen_t)n * G, "t(estimate)"), *ps = real_arg(se_t, (R_xlen_t)n * G, "t(se)");
    double *pc = Rf_isNull(cor_mat) ? NULL : real_arg(cor_mat, (R_xlen_t)M * M, "cor.mat");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, M, M));
    check(h, saver_cuda_cor_adjust(h, pe, ps, n, G, bc, pc, 0, M, REAL(out), 0));
    UNPROTECT(1);
    return out;
}

/* dgCMatrix slots (x@p, x@i, x@x) -> resident sparse counts; replaces the Matrix-package work of
 * R/saver.R:156-157 and as.matrix(x[block, ]) of R/calc_estimate.R:69 */
SEXP C_saver_cuda_set_counts_csc(SEXP hp, SEXP p, SEXP i, SEXP xv, SEXP dim)
{
    saver_cuda_handle h = get_handle(hp);
    int *pd = int_arg(dim, 2, "dim"), *pp = int_arg(p, (R_xlen_t)pd[1] + 1, "x@p");
    const R_xlen_t nnz = pp[pd[1]];
    check(h, saver_cuda_set_counts_csc(h, pp, int_arg(i, nnz, "x@i"), real_arg(xv, nnz, "x@x"), pd[0], pd[1]));
    return R_NilValue;
}

SEXP C_saver_cuda_csc_sums(SEXP hp, SEXP dim)
{
    saver_cuda_handle h = get_handle(hp);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    int *pd = int_arg(dim, 2, "dim");
    SEXP cs = PROTECT(Rf_allocVector(REALSXP, pd[1]));
    SEXP rs = PROTECT(Rf_allocVector(REALSXP, pd[0]));
    check(h, saver_cuda_csc_sums(h, REAL(cs), REAL(rs)));
    SET_VECTOR_ELT(out, 0, cs);
    SET_VECTOR_ELT(out, 1, rs);
    UNPROTECT(3);
    return out;
}

/* t(as.matrix(x[genes, ])) (n x B), or t(log((x[genes, ] + 1) / sf)) when sf is not NULL */
SEXP C_saver_cuda_csc_gather(SEXP hp, SEXP genes0, SEXP sf, SEXP ncells)
{
    saver_cuda_handle h = get_handle(hp);
    const int B = Rf_length(genes0), n = Rf_asInteger(ncells);
    int *pg = int_arg(genes0, B, "genes");
    double *psf = Rf_isNull(sf) ? NULL : real_arg(sf, n, "sf");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, B));
    check(h, saver_cuda_csc_gather(h, pg, B, psf, psf != NULL, REAL(out), 0));
    UNPROTECT(1);
    return out;
}

/* dense counts: x is the R matrix t(x) (cells x genes, so one gene is one contiguous column); double
 * or integer storage.  Replaces the whole-matrix passes of R/saver.R:105-157 (see saver_cuda.R). */
SEXP C_saver_cuda_set_counts_dense(SEXP hp, SEXP xt)
{
    saver_cuda_handle h = get_handle(hp);
    const int n = Rf_nrows(xt), G = Rf_ncols(xt);
    if (Rf_isInteger(xt))
        check(h, saver_cuda_set_counts_dense(h, int_arg(xt, (R_xlen_t)n * G, "t(x)"), 1, G, n));
    else
        check(h, saver_cuda_set_counts_dense(h, real_arg(xt, (R_xlen_t)n * G, "t(x)"), 0, G, n));
    return R_NilValue;
}

SEXP C_saver_cuda_dense_sums(SEXP hp, SEXP dim)
{
    saver_cuda_handle h = get_handle(hp);
    int *pd = int_arg(dim, 2, "dim");
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP cs = PROTECT(Rf_allocVector(REALSXP, pd[1]));
    SEXP rs = PROTECT(Rf_allocVector(REALSXP, pd[0]));
    check(h, saver_cuda_dense_sums(h, REAL(cs), REAL(rs)));
    SET_VECTOR_ELT(out, 0, cs);
    SET_VECTOR_ELT(out, 1, rs);
    UNPROTECT(3);
    return out;
}

/* t(log((x[genes, ] + 1) / sf)) (n x B) from the resident dense counts */
SEXP C_saver_cuda_dense_gather(SEXP hp, SEXP genes0, SEXP sf, SEXP ncells)
{
    saver_cuda_handle h = get_handle(hp);
    const int B = Rf_length(genes0), n = Rf_asInteger(ncells);
    int *pg = int_arg(genes0, B, "genes");
    double *psf = Rf_isNull(sf) ? NULL : real_arg(sf, n, "sf");
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, B));
    check(h, saver_cuda_dense_gather(h, pg, B, psf, psf != NULL, REAL(out), 0));
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"C_saver_cuda_create", (DL_FUNC)&C_saver_cuda_create, 1},
    {"C_saver_cuda_set_design", (DL_FUNC)&C_saver_cuda_set_design, 2},
    {"C_saver_cuda_calc_estimate", (DL_FUNC)&C_saver_cuda_calc_estimate, 13},
    {"C_saver_cuda_calc_post", (DL_FUNC)&C_saver_cuda_calc_post, 5},
    {"C_saver_cuda_calc_var", (DL_FUNC)&C_saver_cuda_calc_var, 5},
    {"C_saver_cuda_calc_var_grid", (DL_FUNC)&C_saver_cuda_calc_var_grid, 5},
    {"C_saver_cuda_loglik", (DL_FUNC)&C_saver_cuda_loglik, 6},
    {"C_saver_cuda_maxcor", (DL_FUNC)&C_saver_cuda_maxcor, 3},
    {"C_saver_cuda_sample", (DL_FUNC)&C_saver_cuda_sample, 6},
    {"C_saver_cuda_cor_adjust", (DL_FUNC)&C_saver_cuda_cor_adjust, 5},
    {"C_saver_cuda_set_counts_csc", (DL_FUNC)&C_saver_cuda_set_counts_csc, 5},
    {"C_saver_cuda_csc_sums", (DL_FUNC)&C_saver_cuda_csc_sums, 2},
    {"C_saver_cuda_csc_gather", (DL_FUNC)&C_saver_cuda_csc_gather, 4},
    {"C_saver_cuda_set_counts_dense", (DL_FUNC)&C_saver_cuda_set_counts_dense, 2},
    {"C_saver_cuda_dense_sums", (DL_FUNC)&C_saver_cuda_dense_sums, 2},
    {"C_saver_cuda_dense_gather", (DL_FUNC)&C_saver_cuda_dense_gather, 4},
    {NULL, NULL, 0}};

void R_init_SAVERcuda(DllInfo* dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
