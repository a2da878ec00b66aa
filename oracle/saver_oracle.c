/*
 * saver_oracle.c -- CPU restatement of the SAVER expression-recovery hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under saver_b200/ links, loads or calls
 * this file.  Allowed users: tests/, __graft_entry__.smoke(), and bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * What it restates (reference = /root/reference, an R package with no native
 * code of its own):
 *   R/calc_loglik.R:26-75        calc.loglik.a/b/k            -> orc_loglik
 *   R/optimize_variance.R:24-134 calc.a/b/k                   -> orc_calc_var
 *   R/calc_posterior.R:26-68     calc.post                    -> orc_calc_post
 *   R/calc_maxcor.R:26-42        calc.maxcor                  -> orc_maxcor
 *   R/utils.R:129-134            est.lambda                   -> orc_est_lambda
 *   R/expr_predict.R:33-85       expr.predict (both variants) -> orc_predict_cv,
 *                                                                orc_predict_path
 * and the out-of-tree arithmetic those call (NOT under /root/reference; the
 * dependency versions are unpinned there -- DESCRIPTION:20-22):
 *   base R stats::optimize  (Brent_fmin, tol = DBL_EPSILON^0.25)  -> brent_fmin
 *   base R stats::uniroot   (R_zeroin2,  tol = DBL_EPSILON^0.25)  -> zeroin2
 *   base R var/mean/sum     (long double accumulation, 2-pass mean)
 *   glmnet::glmnet(family="poisson") = Fortran fishnet/fishnet1   -> fishnet
 *   glmnet::cv.glmnet(nfolds=5) deviance CV glue                  -> orc_predict_cv
 * Those are restated from their published algorithms.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against the
 * reference's own bundled golden object (data/linnarsson_saver.rda, committed
 * as tests/golden/linnarsson.npz): maxcor, lambda.max, est.lambda, the
 * Brent + calc.b + posterior chain (digit-exact on the deterministic genes)
 * and the fixed-lambda lasso chain (to the 3-decimal quantiser).  NOT pinned:
 * the CV fold draw (R's RNG stream is not recoverable from the object) and the
 * stochastic fallback draw of optimize_variance.R:53 -- see DESIGN.md.
 *
 * Deliberate, documented deviation: R/optimize_variance.R:53 replaces the
 * parameter by mean(sample(samp, 10000, replace=TRUE, prob=p)), a Monte-Carlo
 * estimate of sum(samp*p).  Without R's RNG stream the draw is not
 * reproducible; this file (and the CUDA path) use the expectation sum(samp*p).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/*
 * Host threads INSIDE one gene (bench.py's CPU arm at the 20k x 10k shape, where one gene is
 * minutes of scalar work): the loops over predictor columns that dominate glmnet's cost (per-fit
 * standardisation, the full-gradient / KKT passes, the curvature pass over the strong set) are
 * split over `orc_inner_threads` threads, and the nfolds fold fits of cv.glmnet run on
 * `orc_outer_threads` teams at once.  Every column is still reduced by ONE thread in the original
 * order, so results are bit-identical to the single-threaded run (tests/test_oracle_golden.py).
 * Default 1 x 1: the callers that map genes over a thread pool keep one gene per thread.
 */
static int orc_outer_threads = 1, orc_inner_threads = 1;
static _Thread_local int tl_threads = 0; /* threads of the fit running on this thread (0: inner) */
static int fit_threads(void) { return tl_threads > 0 ? tl_threads : orc_inner_threads; }
ORC_API void orc_set_threads(int outer, int inner)
{
    orc_outer_threads = outer > 1 ? outer : 1;
    orc_inner_threads = inner > 1 ? inner : 1;
#ifdef _OPENMP
    omp_set_max_active_levels(2);
#endif
}

typedef long double ld;

/* ------------------------------------------------------------------ */
/* base-R style reductions (long double accumulators, two-pass mean)   */
/* ------------------------------------------------------------------ */
static double r_mean(const double *x, int n)
{
    ld s = 0;
    for (int i = 0; i < n; ++i) s += x[i];
    s /= n;
    if (isfinite((double)s)) {
        ld t = 0;
        for (int i = 0; i < n; ++i) t += (x[i] - s);
        s += t / n;
    }
    return (double)s;
}

static double r_var(const double *x, int n)
{
    /* stats::var -> cov.c: refined mean, then sum of squares / (n-1) */
    ld s = 0;
    for (int i = 0; i < n; ++i) s += x[i];
    ld m = s / n;
    if (isfinite((double)m)) {
        ld t = 0;
        for (int i = 0; i < n; ++i) t += (x[i] - m);
        m += t / n;
    }
    ld ss = 0;
    for (int i = 0; i < n; ++i) {
        ld d = x[i] - m;
        ss += d * d;
    }
    return (double)(ss / (n - 1));
}

/* ------------------------------------------------------------------ */
/* calc.loglik.a / .b / .k   (R/calc_loglik.R:26-75)                   */
/* model: 0 = a (constant CV), 1 = b (constant Fano), 2 = k (constant  */
/* variance).  y raw counts, mu prediction, sf size factors.           */
/* ------------------------------------------------------------------ */
typedef struct {
    int model, n;
    const double *y, *mu, *sf;
    int nevals;
} ll_ctx;

static double loglik_a(double a, const ll_ctx *c)
{
    /* R/calc_loglik.R:34-39 */
    const int n = c->n;
    double ia = 1 / a;
    double f1 = n / a * log(ia);
    ld s2 = 0, s4 = 0, s5 = 0;
    for (int i = 0; i < n; ++i) {
        s2 += ia * log(c->mu[i]);
        s4 += lgamma(c->y[i] + ia);
        s5 += (c->y[i] + ia) * log(c->sf[i] + 1 / (a * c->mu[i]));
    }
    double f2 = -(double)s2, f3 = -n * lgamma(ia), f4 = (double)s4, f5 = -(double)s5;
    ld tot = 0;
    tot += f1; tot += f2; tot += f3; tot += f4; tot += f5;
    return -(double)tot;
}

static double loglik_b(double b, const ll_ctx *c)
{
    /* R/calc_loglik.R:52-56 */
    const int n = c->n;
    double lib = log(1 / b), ib = 1 / b;
    ld s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    for (int i = 0; i < n; ++i) {
        double mb = c->mu[i] / b;
        s1 += mb * lib;
        s2 += lgamma(mb);
        s3 += lgamma(c->y[i] + mb);
        s4 += (c->y[i] + mb) * log(c->sf[i] + ib);
    }
    ld tot = 0;
    tot += (double)s1; tot += -(double)s2; tot += (double)s3; tot += -(double)s4;
    return -(double)tot;
}

static double loglik_k(double k, const ll_ctx *c)
{
    /* R/calc_loglik.R:69-74 */
    const int n = c->n;
    double lk = log(k);
    ld s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    for (int i = 0; i < n; ++i) {
        double m = c->mu[i], m2 = m * m;
        s3 += m2 * log(m) / k;
        s4 += m2 * lk / k;
        s5 += lgamma(m2 / k);
        s6 += lgamma(c->y[i] + m2 / k);
        s7 += (c->y[i] + m2 / k) * log(c->sf[i] + m / k);
    }
    ld tot = 0;
    tot += (double)s3; tot += -(double)s4; tot += -(double)s5; tot += (double)s6; tot += -(double)s7;
    return -(double)tot;
}

static double loglik_eval(double p, ll_ctx *c)
{
    c->nevals++;
    if (c->model == 0) return loglik_a(p, c);
    if (c->model == 1) return loglik_b(p, c);
    return loglik_k(p, c);
}

ORC_API double orc_loglik(int model, double param, const double *y, const double *mu,
                          const double *sf, int n)
{
    ll_ctx c = {model, n, y, mu, sf, 0};
    return loglik_eval(param, &c);
}

/* ------------------------------------------------------------------ */
/* stats::optimize -> Brent_fmin (golden section + parabolic steps).   */
/* Non-finite objective values are replaced by DBL_MAX as do_fmin does.*/
/* ------------------------------------------------------------------ */
typedef double (*fn1d)(double, void *);

static double fmin_guard(fn1d f, double x, void *ctx)
{
    double v = f(x, ctx);
    return isfinite(v) ? v : DBL_MAX;
}

static double brent_fmin(double lo, double hi, fn1d f, void *ctx, double tol)
{
    const double gold = (3.0 - sqrt(5.0)) * 0.5;
    const double sqeps = sqrt(DBL_EPSILON);
    double a = lo, b = hi;
    double x = a + gold * (b - a), w = x, v = x;
    double step = 0.0, prev = 0.0;
    double fx = fmin_guard(f, x, ctx), fw = fx, fv = fx;
    const double tol3 = tol / 3.0;
    for (;;) {
        double mid = (a + b) * 0.5;
        double tol1 = sqeps * fabs(x) + tol3;
        double tol2 = tol1 * 2.0;
        if (fabs(x - mid) <= tol2 - (b - a) * 0.5) break;
        double p = 0.0, q = 0.0, r = 0.0;
        if (fabs(prev) > tol1) {
            r = (x - w) * (fx - fv);
            q = (x - v) * (fx - fw);
            p = (x - v) * q - (x - w) * r;
            q = (q - r) * 2.0;
            if (q > 0.0) p = -p; else q = -q;
            r = prev;
            prev = step;
        }
        double u;
        if (fabs(p) >= fabs(q * 0.5 * r) || p <= q * (a - x) || p >= q * (b - x)) {
            prev = (x < mid) ? b - x : a - x;
            step = gold * prev;
        } else {
            step = p / q;
            u = x + step;
            if (u - a < tol2 || b - u < tol2) {
                step = tol1;
                if (x >= mid) step = -step;
            }
        }
        if (fabs(step) >= tol1) u = x + step;
        else if (step > 0.0) u = x + tol1;
        else u = x - tol1;
        double fu = fmin_guard(f, u, ctx);
        if (fu <= fx) {
            if (u < x) b = x; else a = x;
            v = w; fv = fw;
            w = x; fw = fx;
            x = u; fx = fu;
        } else {
            if (u < x) a = u; else b = u;
            if (fu <= fw || w == x) {
                v = w; fv = fw;
                w = u; fw = fu;
            } else if (fu <= fv || v == x || v == w) {
                v = u; fv = fu;
            }
        }
    }
    return x;
}

/* stats::uniroot -> R_zeroin2.  fa, fb are the end-point values. */
static double zeroin2(double ax, double bx, double fa, double fb, fn1d f, void *ctx,
                      double tol, int maxit)
{
    double a = ax, b = bx, c = a, fc = fa;
    if (fa == 0.0) return a;
    if (fb == 0.0) return b;
    maxit += 1;
    while (maxit--) {
        double prev_step = b - a;
        if (fabs(fc) < fabs(fb)) {
            a = b; b = c; c = a;
            fa = fb; fb = fc; fc = fa;
        }
        double tol_act = 2 * DBL_EPSILON * fabs(b) + tol / 2;
        double new_step = (c - b) / 2;
        if (fabs(new_step) <= tol_act || fb == 0.0) return b;
        if (fabs(prev_step) >= tol_act && fabs(fa) > fabs(fb)) {
            double p, q, cb = c - b;
            if (a == c) {
                double t1 = fb / fa;
                p = cb * t1;
                q = 1.0 - t1;
            } else {
                double t1, t2;
                q = fa / fc; t1 = fb / fc; t2 = fb / fa;
                p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0));
                q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0);
            }
            if (p > 0.0) q = -q; else p = -p;
            if (p < (0.75 * cb * q - fabs(tol_act * q) / 2) && p < fabs(prev_step * q / 2))
                new_step = p / q;
        }
        if (fabs(new_step) < tol_act) new_step = (new_step > 0.0) ? tol_act : -tol_act;
        a = b; fa = fb;
        b += new_step;
        fb = f(b, ctx);
        if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) {
            c = a; fc = fa;
        }
    }
    return b;
}

/* ------------------------------------------------------------------ */
/* calc.a / calc.b / calc.k   (R/optimize_variance.R:24-134)           */
/* out[0] = 1/a | 1/b | k, out[1] = negative log-lik there.            */
/* Returns 0 ok, 1 = the R code would have raised an error (caller     */
/* mirrors tryCatch -> c(0, Inf), R/calc_posterior.R:41-49).           */
/* flags: bit0 fallback branch taken, bit1 uniroot used.               */
/* grid_ll (optional, 100): log-lik on the fallback grid; grid_x (100) */
/* ------------------------------------------------------------------ */
static double ll_obj(double p, void *ctx) { return loglik_eval(p, (ll_ctx *)ctx); }

typedef struct { ll_ctx *c; double shift; } root_ctx;
static double ll_root(double p, void *ctx)
{
    root_ctx *r = (root_ctx *)ctx;
    return loglik_eval(p, r->c) + r->shift;
}

#define BRENT_TOL 1.220703125e-4 /* .Machine$double.eps^0.25 */

static int calc_var_core(int model, const double *y, const double *mu, const double *sf, int n,
                         double *out, int *flags, int *nevals, double *grid_x, double *grid_ll)
{
    ll_ctx c = {model, n, y, mu, sf, 0};
    *flags = 0;
    if (model == 1) { /* R/optimize_variance.R:69-71 */
        ld sm = 0;
        for (int i = 0; i < n; ++i) sm += mu[i];
        if ((double)sm == 0) { out[0] = 0; out[1] = 0; if (nevals) *nevals = 0; return 0; }
    }
    double *ysf = (double *)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) ysf[i] = y[i] / sf[i];
    double vr = r_var(ysf, n), mn = r_mean(ysf, n);
    free(ysf);
    double upper = (model == 0) ? vr / (mn * mn) : (model == 1) ? vr / mn : vr;

    double par = brent_fmin(0.0, upper, ll_obj, &c, BRENT_TOL);
    double nll = loglik_eval(par, &c);
    double ll_min = -loglik_eval(1e-05, &c);
    double ll_mle = -loglik_eval(par, &c);
    double ll_max = -loglik_eval(upper, &c);
    int rc = 0;
    if (isnan(ll_mle - ll_min)) { rc = 1; goto done; } /* `if (NA)` is an R error */
    if (ll_mle - ll_min < 0.5) {
        *flags |= 1;
        double gmax;
        if (isnan(ll_mle - ll_max)) { rc = 1; goto done; }
        if (ll_mle - ll_max > 10) {
            root_ctx rcx = {&c, ll_mle - 10};
            double flo = ll_root(1e-05, &rcx), fhi = ll_root(upper, &rcx);
            if (isnan(flo) || isnan(fhi) || flo * fhi > 0) { rc = 1; goto done; }
            gmax = zeroin2(1e-05, upper, flo, fhi, ll_root, &rcx, BRENT_TOL, 1000);
            *flags |= 2;
        } else {
            gmax = upper;
        }
        double samp[100], ll[100], llmin = INFINITY;
        for (int i = 0; i < 100; ++i) {
            double pp = (i + 1 - 0.5) / 100.0; /* ppoints(100) */
            samp[i] = (exp((exp(pp) - 1) / 2) - 1) * gmax;
            ll[i] = -loglik_eval(samp[i], &c);
            if (ll[i] < llmin) llmin = ll[i];
            if (isnan(ll[i])) { rc = 1; }
        }
        if (grid_x) memcpy(grid_x, samp, sizeof samp);
        if (grid_ll) memcpy(grid_ll, ll, sizeof ll);
        if (rc) goto done; /* sample(): "NA in probability vector" */
        ld tot = 0, acc = 0;
        double wgt[100];
        for (int i = 0; i < 100; ++i) { wgt[i] = exp(ll[i] - llmin); tot += wgt[i]; }
        for (int i = 0; i < 100; ++i) acc += (ld)samp[i] * (wgt[i] / (double)tot);
        if (!isfinite((double)acc)) { rc = 1; goto done; }
        par = (double)acc; /* E[mean(sample(samp, 10000, TRUE, prob))], see header */
        nll = loglik_eval(par, &c);
    }
    out[0] = (model == 2) ? par : 1 / par;
    out[1] = nll;
done:
    if (nevals) *nevals = c.nevals;
    if (rc) { out[0] = 0; out[1] = INFINITY; }
    return rc;
}

ORC_API int orc_calc_var(int model, const double *y, const double *mu, const double *sf, int n,
                         double *out2, int *flags, int *nevals, double *grid_x, double *grid_ll)
{
    return calc_var_core(model, y, mu, sf, n, out2, flags, nevals, grid_x, grid_ll);
}

/* ------------------------------------------------------------------ */
/* calc.post   (R/calc_posterior.R:26-68)                              */
/* est_raw/se_raw: posterior mean / sd before the 3-decimal ceiling;   */
/* est/se: ceiling(v*1000*scale_sf)/1000.  Any output may be NULL.     */
/* info[0]=chosen model (0 a,1 b,2 k; 3 = constant-mu branch; -1 zero  */
/* gene), info[1]=prior parameter, info[2..4]=nll a,b,k,               */
/* info[5]=fallback bitmask (bit m: model m took the grid branch).     */
/* returns 0 ok, 2 if R would have stopped (error outside tryCatch).   */
/* ------------------------------------------------------------------ */
ORC_API int orc_calc_post(const double *y, const double *mu, const double *sf, int n,
                          double scale_sf, double *est, double *se, double *est_raw,
                          double *se_raw, double *info)
{
    ld sy = 0;
    for (int i = 0; i < n; ++i) sy += y[i];
    double inf_[6] = {-1, 0, NAN, NAN, NAN, 0};
    if ((double)sy == 0) { /* :34-36 */
        for (int i = 0; i < n; ++i) {
            if (est) est[i] = 0; if (se) se[i] = 0;
            if (est_raw) est_raw[i] = 0; if (se_raw) se_raw[i] = 0;
        }
        if (info) memcpy(info, inf_, sizeof inf_);
        return 0;
    }
    int method, flags, fb = 0;
    double res[3][2];
    if (r_var(mu, n) == 0) { /* :37-39 */
        if (calc_var_core(1, y, mu, sf, n, res[1], &flags, NULL, NULL, NULL)) return 2;
        method = 3;
        fb = (flags & 1) << 1;
        inf_[3] = res[1][1];
    } else {
        for (int m = 0; m < 3; ++m) { /* :41-49 */
            calc_var_core(m, y, mu, sf, n, res[m], &flags, NULL, NULL, NULL);
            fb |= (flags & 1) << m;
            inf_[2 + m] = res[m][1];
        }
        /* which.min skips NaN, first minimum wins; all-NaN cannot index -> error */
        method = -1;
        for (int m = 0; m < 3; ++m)
            if (!isnan(res[m][1]) && (method < 0 || res[m][1] < res[method][1])) method = m;
        if (method < 0) return 2;
    }
    inf_[0] = method;
    inf_[1] = res[method == 3 ? 1 : method][0];
    inf_[5] = fb;
    for (int i = 0; i < n; ++i) {
        double al, be, m = mu[i];
        if (method == 0) { al = res[0][0]; be = res[0][0] / m; }
        else if (method == 1 || method == 3) { be = res[1][0]; al = m * res[1][0]; }
        else { al = m * m / res[2][0]; be = m / res[2][0]; }
        double pa = al + y[i], pb = be + sf[i];
        double lam = pa / pb, s = sqrt(pa / (pb * pb));
        if (est_raw) est_raw[i] = lam;
        if (se_raw) se_raw[i] = s;
        if (est) est[i] = ceil(lam * 1000 * scale_sf) / 1000;
        if (se) se[i] = ceil(s * 1000 * scale_sf) / 1000;
    }
    if (info) memcpy(info, inf_, sizeof inf_);
    return 0;
}

/* ------------------------------------------------------------------ */
/* calc.maxcor   (R/calc_maxcor.R:26-42) : Pearson cor of every column */
/* of x (n x p, column-major) with y; NA -> 0; the pair whose names    */
/* match (self, -1 = none) zeroed; max |r| returned.                   */
/* stats::cor is cov.c: refined long-double means, then sums.          */
/* ------------------------------------------------------------------ */
static void col_center_stats(const double *v, int n, double *mean, double *ss)
{
    ld s = 0;
    for (int i = 0; i < n; ++i) s += v[i];
    ld m = s / n, t = 0;
    for (int i = 0; i < n; ++i) t += (v[i] - m);
    m += t / n;
    ld q = 0;
    for (int i = 0; i < n; ++i) { ld d = v[i] - m; q += d * d; }
    *mean = (double)m;
    *ss = (double)q;
}

ORC_API double orc_maxcor(const double *x, int n, int p, const double *y, int self)
{
    double ym, yss;
    col_center_stats(y, n, &ym, &yss);
    double ysd = sqrt(yss / (n - 1));
    double best = 0;
    for (int j = 0; j < p; ++j) {
        if (j == self) continue;
        const double *xj = x + (size_t)j * n;
        double xm, xss;
        col_center_stats(xj, n, &xm, &xss);
        ld c = 0;
        for (int i = 0; i < n; ++i) c += ((ld)xj[i] - xm) * ((ld)y[i] - ym);
        double xsd = sqrt(xss / (n - 1));
        double r = (double)(c / (n - 1)) / (xsd * ysd);
        if (isnan(r)) r = 0;
        if (r > 1) r = 1;
        if (r < -1) r = -1;
        if (fabs(r) > best) best = fabs(r);
    }
    return best;
}

/* est.lambda   (R/utils.R:129-134).  y = all cells, r = maxcor */
ORC_API void orc_est_lambda(const double *y, int n, double r, const double *coefs, double *out2)
{
    double sdv = sqrt(r_var(y, n));
    double lmax = r * sdv * sqrt((n - 1.0) / n);
    double arg = coefs[0] + coefs[1] * r;
    if (arg < 0) arg = 0;
    out2[0] = lmax;
    out2[1] = lmax * exp(-sqrt(arg));
}

/* ------------------------------------------------------------------ */
/* glmnet::glmnet(x, y, family = "poisson", dfmax) -- dense Poisson    */
/* lasso path (Fortran fishnet -> fishnet1), as called from            */
/* R/expr_predict.R:41-43,65-67.  alpha = 1, standardize = TRUE,       */
/* intercept = TRUE, unit observation weights, no offset.              */
/* ------------------------------------------------------------------ */
typedef struct {
    int lmu;        /* number of solutions kept */
    int jerr;       /* 0 ok; <0 glmnet warnings (path truncated); >0 fatal */
    int nlp;        /* CD sweeps */
    int ngrad;      /* full-gradient (KKT) passes */
    double nulldev; /* dev0 in fishnet1's normalisation */
} fish_info;

/*
 * xs  : standardised design, m rows x p columns, column-major (own copy).
 * ju  : 1 = column eligible.
 * y   : m responses; q = 1/m each.
 * nlam, flmin, ulam: flmin < 1 -> automatic geometric sequence; else ulam.
 * ne (dfmax), nx (pmax), thr, maxit.
 * Outputs: a0[nlam], ca[nx*nlam] (compressed coefs on standardised scale),
 * ia[nx] (0-based column of compressed slot), kin[nlam], dev[nlam], alm[nlam].
 */
/* diagnostics (single-threaded use only): per-lambda sweep / gradient-pass counters of the last fit */
static int dbg_nlp[128], dbg_ngrad[128], dbg_nin[128];
ORC_API void orc_debug_last_fit(int *nlp, int *ngrad, int *nin)
{
    memcpy(nlp, dbg_nlp, sizeof dbg_nlp); memcpy(ngrad, dbg_ngrad, sizeof dbg_ngrad);
    memcpy(nin, dbg_nin, sizeof dbg_nin);
}

static void fishnet1(int m, int p, const double *xs, const int *ju, const double *y, int nlam,
                     double flmin, const double *ulam, int ne, int nx, double thr, int maxit,
                     double *a0, double *ca, int *ia, int *kin, double *dev, double *alm,
                     fish_info *fi)
{
    const double sml = 1.0e-5 * 10.0, eps = 1.0e-6, big = 9.9e35, devmax = 0.999;
    const int mnlam = 5;
    const double fmax = log(DBL_MAX * 0.1);
    double *a = (double *)calloc(p, sizeof(double)), *as = (double *)calloc(p, sizeof(double));
    double *ga = (double *)calloc(p, sizeof(double)), *v = (double *)calloc(p, sizeof(double));
    int *mm = (int *)calloc(p, sizeof(int)), *ixx = (int *)calloc(p, sizeof(int));
    double *t = (double *)malloc(sizeof(double) * m), *w = (double *)malloc(sizeof(double) * m);
    double *wr = (double *)malloc(sizeof(double) * m), *f = (double *)malloc(sizeof(double) * m);
    const double q = 1.0 / m;
    double yb = 0;
    for (int i = 0; i < m; ++i) { t[i] = q * y[i]; yb += t[i]; }
    if (!(yb > 0.0)) { /* fishnet: "no positive observations" -> fatal jerr, R raises an error */
        fi->jerr = 9999; fi->lmu = 0; fi->nlp = 0; fi->ngrad = 0; fi->nulldev = 0;
        free(a); free(as); free(ga); free(v); free(mm); free(ixx);
        free(t); free(w); free(wr); free(f);
        return;
    }
    double az = log(yb), dv0 = yb * (az - 1.0);
    for (int i = 0; i < m; ++i) { w[i] = q * yb; f[i] = az; wr[i] = t[i] - w[i]; }
    double v0 = yb, dvr = -yb;
    for (int i = 0; i < m; ++i) if (t[i] > 0.0) dvr += t[i] * log(y[i]);
    dvr -= dv0;
    fi->nulldev = dvr;
    fi->jerr = 0; fi->lmu = 0; fi->nlp = 0; fi->ngrad = 1;
    double alf = 1.0;
    if (flmin < 1.0) {
        double eqs = flmin > eps ? flmin : eps;
        /* Fortran: alf=eqs**(1.0/(nlam-1)) -- the exponent is a REAL*4 expression.  Pinned by
         * the golden lambda.min values (they sit on this grid to 1e-15, and 3e-9 off the
         * double-precision-exponent grid). */
        alf = pow(eqs, (double)(1.0f / (float)(nlam - 1)));
    }
    int nin = 0, nlp = 0, mnl = mnlam < nlam ? mnlam : nlam;
    const double shr = thr * dvr;
    double al = 0.0;
    const int nth = fit_threads();
    (void)nth;
#pragma omp parallel for schedule(static) num_threads(nth) if (nth > 1 && (size_t)p * m > 4000000)
    for (int j = 0; j < p; ++j) {
        if (!ju[j]) continue;
        const double *xj = xs + (size_t)j * m;
        double s = 0;
        for (int i = 0; i < m; ++i) s += wr[i] * xj[i];
        ga[j] = fabs(s);
    }
    for (int ilm = 0; ilm < nlam; ++ilm) {
        double al0 = al;
        if (flmin >= 1.0) al = ulam[ilm];
        else if (ilm > 1) al *= alf;
        else if (ilm == 0) al = big;
        else {
            al0 = 0.0;
            for (int j = 0; j < p; ++j) if (ju[j] && ga[j] > al0) al0 = ga[j];
            al = alf * al0;
        }
        double tlam = 2.0 * al - al0;
        for (int k = 0; k < p; ++k)
            if (!ixx[k] && ju[k] && ga[k] > tlam) ixx[k] = 1;
        int overflow = 0;
        for (;;) { /* IRLS outer loop */
            double az0 = az;
            for (int l = 0; l < nin; ++l) as[ia[l]] = a[ia[l]];
#pragma omp parallel for schedule(dynamic, 64) num_threads(nth) if (nth > 1 && (size_t)p * m > 4000000)
            for (int j = 0; j < p; ++j) {
                if (!ixx[j]) continue;
                const double *xj = xs + (size_t)j * m;
                double s = 0;
                for (int i = 0; i < m; ++i) s += w[i] * xj[i] * xj[i];
                v[j] = s;
            }
            for (;;) { /* sweep over the strong set */
                ++nlp;
                double dlx = 0.0;
                for (int k = 0; k < p; ++k) {
                    if (!ixx[k]) continue;
                    const double *xk = xs + (size_t)k * m;
                    double ak = a[k], s = 0;
                    for (int i = 0; i < m; ++i) s += wr[i] * xk[i];
                    double u = s + v[k] * ak, au = fabs(u) - al;
                    a[k] = (au > 0.0) ? copysign(au, u) / v[k] : 0.0;
                    if (a[k] == ak) continue;
                    double d = a[k] - ak;
                    if (v[k] * d * d > dlx) dlx = v[k] * d * d;
                    for (int i = 0; i < m; ++i) { wr[i] -= d * w[i] * xk[i]; f[i] += d * xk[i]; }
                    if (mm[k] == 0) {
                        ++nin;
                        if (nin > nx) break;
                        mm[k] = nin;
                        ia[nin - 1] = k;
                    }
                }
                if (nin > nx) break;
                {
                    double s = 0;
                    for (int i = 0; i < m; ++i) s += wr[i];
                    double d = s / v0;
                    az += d;
                    if (v0 * d * d > dlx) dlx = v0 * d * d;
                    for (int i = 0; i < m; ++i) { wr[i] -= d * w[i]; f[i] += d; }
                }
                if (dlx < shr) break;
                if (nlp > maxit) { fi->jerr = -(ilm + 1); goto finish; }
                for (;;) { /* sweeps over the ever-active list */
                    ++nlp;
                    dlx = 0.0;
                    for (int l = 0; l < nin; ++l) {
                        int k = ia[l];
                        const double *xk = xs + (size_t)k * m;
                        double ak = a[k], s = 0;
                        for (int i = 0; i < m; ++i) s += wr[i] * xk[i];
                        double u = s + v[k] * ak, au = fabs(u) - al;
                        a[k] = (au > 0.0) ? copysign(au, u) / v[k] : 0.0;
                        if (a[k] == ak) continue;
                        double d = a[k] - ak;
                        if (v[k] * d * d > dlx) dlx = v[k] * d * d;
                        for (int i = 0; i < m; ++i) { wr[i] -= d * w[i] * xk[i]; f[i] += d * xk[i]; }
                    }
                    double s = 0;
                    for (int i = 0; i < m; ++i) s += wr[i];
                    double d = s / v0;
                    az += d;
                    if (v0 * d * d > dlx) dlx = v0 * d * d;
                    for (int i = 0; i < m; ++i) { wr[i] -= d * w[i]; f[i] += d; }
                    if (dlx < shr) break;
                    if (nlp > maxit) { fi->jerr = -(ilm + 1); goto finish; }
                }
            }
            if (nin > nx) { overflow = 1; break; }
            v0 = 0;
            for (int i = 0; i < m; ++i) {
                double fc = fabs(f[i]) < fmax ? f[i] : copysign(fmax, f[i]);
                w[i] = q * exp(fc);
                v0 += w[i];
                wr[i] = t[i] - w[i];
            }
            if (v0 * (az - az0) * (az - az0) >= shr) continue;
            int changed = 0;
            for (int l = 0; l < nin; ++l) {
                int k = ia[l];
                double d = a[k] - as[k];
                if (v[k] * d * d >= shr) { changed = 1; break; }
            }
            if (changed) continue;
            fi->ngrad++;
            int viol = 0;
#pragma omp parallel for schedule(static) num_threads(nth) reduction(| : viol) if (nth > 1 && (size_t)p * m > 4000000)
            for (int k = 0; k < p; ++k) {
                if (ixx[k] || !ju[k]) continue;
                const double *xk = xs + (size_t)k * m;
                double s = 0;
                for (int i = 0; i < m; ++i) s += wr[i] * xk[i];
                ga[k] = fabs(s);
                if (ga[k] > al) { ixx[k] = 1; viol = 1; }
            }
            if (!viol) break;
        }
        if (overflow) { fi->jerr = -10000 - (ilm + 1); break; }
        for (int l = 0; l < nin; ++l) ca[(size_t)ilm * nx + l] = a[ia[l]];
        kin[ilm] = nin;
        if (ilm < 128) { dbg_nlp[ilm] = nlp; dbg_ngrad[ilm] = fi->ngrad; dbg_nin[ilm] = nin; }
        a0[ilm] = az;
        alm[ilm] = al;
        fi->lmu = ilm + 1;
        {
            double s = 0;
            for (int i = 0; i < m; ++i) s += t[i] * f[i];
            dev[ilm] = (s - v0 - dv0) / dvr;
        }
        if (ilm + 1 < mnl || flmin >= 1.0) continue;
        int me = 0;
        for (int l = 0; l < nin; ++l) if (ca[(size_t)ilm * nx + l] != 0.0) ++me;
        if (me > ne) break;
        if ((dev[ilm] - dev[ilm - mnl + 1]) / dev[ilm] < sml) break;
        if (dev[ilm] > devmax) break;
    }
finish:
    fi->nlp = nlp;
    free(a); free(as); free(ga); free(v); free(mm); free(ixx);
    free(t); free(w); free(wr); free(f);
}

/*
 * The R-level glmnet() + Fortran fishnet wrapper: row subset, chkvars,
 * weighted standardisation (1/m variance), path, un-standardise, fix.lam.
 *   x      : n_all x p column-major raw design; rows[m] selects observations.
 *   excl   : column dropped before the fit (R/calc_estimate.R:103-110), -1 none.
 *   lambda : NULL -> automatic sequence (nlambda = 100,
 *            lambda.min.ratio = nobs < nvars ? 0.01 : 1e-4); else nlam user
 *            values (sorted decreasing here, as glmnet does).
 * Outputs (caller-allocated, nlam_max = 100 or nlam): lam[], a0[], beta
 * (dense p x nlam column-major, ORIGINAL scale; excluded column = 0), dev[].
 */
typedef struct {
    int lmu, jerr, nlp, ngrad, nx;
} glm_info;

static int cmp_desc(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x < y) - (x > y);
}

static void glmnet_poisson(const double *x, int n_all, int p, const int *rows, int m,
                           const double *y_all, int excl, const double *lambda, int nlam_in,
                           int dfmax, double thr, int maxit, double *lam, double *a0,
                           double *beta, double *dev, glm_info *gi)
{
    int nvars = p - (excl >= 0 ? 1 : 0);
    int nlam = lambda ? nlam_in : 100;
    int nx = dfmax * 2 + 20;
    if (nx > nvars) nx = nvars;
    double flmin = lambda ? 1.0 : (m < nvars ? 0.01 : 1e-4);
    double *ulam = NULL;
    if (lambda) {
        ulam = (double *)malloc(sizeof(double) * nlam);
        memcpy(ulam, lambda, sizeof(double) * nlam);
        qsort(ulam, nlam, sizeof(double), cmp_desc);
    }
    double *xs = (double *)malloc(sizeof(double) * (size_t)m * p);
    double *xm = (double *)malloc(sizeof(double) * p), *xsd = (double *)malloc(sizeof(double) * p);
    int *ju = (int *)malloc(sizeof(int) * p);
    double *y = (double *)malloc(sizeof(double) * m);
    for (int i = 0; i < m; ++i) y[i] = y_all[rows[i]];
    const double q = 1.0 / m;
    const int nth = fit_threads();
    (void)nth;
#pragma omp parallel for schedule(static) num_threads(nth) if (nth > 1 && (size_t)p * m > 4000000)
    for (int j = 0; j < p; ++j) {
        const double *xj = x + (size_t)j * n_all;
        double *sj = xs + (size_t)j * m;
        ju[j] = 0;
        xm[j] = 0; xsd[j] = 1;
        if (j == excl) { memset(sj, 0, sizeof(double) * m); continue; }
        for (int i = 0; i < m; ++i) sj[i] = xj[rows[i]];
        for (int i = 1; i < m; ++i) if (sj[i] != sj[0]) { ju[j] = 1; break; } /* chkvars */
        if (!ju[j]) continue;
        double s = 0;
        for (int i = 0; i < m; ++i) s += q * sj[i];
        xm[j] = s;
        for (int i = 0; i < m; ++i) sj[i] -= s;
        double s2 = 0;
        for (int i = 0; i < m; ++i) s2 += q * sj[i] * sj[i];
        xsd[j] = sqrt(s2);
        for (int i = 0; i < m; ++i) sj[i] /= xsd[j];
    }
    double *ca = (double *)calloc((size_t)nx * nlam, sizeof(double));
    int *ia = (int *)calloc(nx > 0 ? nx : 1, sizeof(int)), *kin = (int *)calloc(nlam, sizeof(int));
    fish_info fi;
    fishnet1(m, p, xs, ju, y, nlam, flmin, ulam, dfmax, nx, thr, maxit, a0, ca, ia, kin, dev, lam,
             &fi);
    memset(beta, 0, sizeof(double) * (size_t)p * nlam);
    for (int l = 0; l < fi.lmu; ++l) {
        double *bl = beta + (size_t)l * p;
        double shift = 0;
        for (int k = 0; k < kin[l]; ++k) {
            int j = ia[k];
            double b = ca[(size_t)l * nx + k] / xsd[j];
            bl[j] = b;
            shift += b * xm[j];
        }
        a0[l] -= shift;
    }
    if (!lambda && fi.lmu > 2) lam[0] = exp(2 * log(lam[1]) - log(lam[2])); /* fix.lam */
    gi->lmu = fi.lmu; gi->jerr = fi.jerr; gi->nlp = fi.nlp; gi->ngrad = fi.ngrad; gi->nx = nx;
    free(ulam); free(xs); free(xm); free(xsd); free(ju); free(y); free(ca); free(ia); free(kin);
}

/* low-level export for tests: one glmnet(family="poisson") fit */
ORC_API int orc_glmnet_poisson(const double *x, int n_all, int p, const int *rows, int m,
                               const double *y_all, int excl, const double *lambda, int nlam,
                               int dfmax, double thr, double *lam, double *a0, double *beta,
                               double *dev, int *stats /* lmu, jerr, nlp, ngrad */)
{
    glm_info gi;
    glmnet_poisson(x, n_all, p, rows, m, y_all, excl, lambda, nlam, dfmax, thr, 100000, lam, a0,
                   beta, dev, &gi);
    stats[0] = gi.lmu; stats[1] = gi.jerr; stats[2] = gi.nlp; stats[3] = gi.ngrad;
    return gi.jerr > 0;
}

static double sd_all(const double *y, int n) { return sqrt(r_var(y, n)); }

static void fill_mean(const double *y, const int *rows, int m, int n, double *mu)
{
    ld s = 0;
    for (int i = 0; i < m; ++i) s += y[rows[i]];
    double mn = (double)(s / m);
    for (int i = 0; i < n; ++i) mu[i] = mn;
}

static void predict_all(const double *x, int n, int p, double a0, const double *b, double *mu)
{
    for (int i = 0; i < n; ++i) mu[i] = a0;
    for (int j = 0; j < p; ++j) {
        if (b[j] == 0.0) continue;
        const double *xj = x + (size_t)j * n;
        for (int i = 0; i < n; ++i) mu[i] += xj[i] * b[j];
    }
    for (int i = 0; i < n; ++i) mu[i] = exp(mu[i]);
}

/*
 * expr.predict, CV variant (R/expr_predict.R:39-60) = cv.glmnet(nfolds) with
 * the fold assignment passed in (foldid[m], values 1..nfolds, aligned with
 * rows[]); type.measure = Poisson deviance, grouped = TRUE.
 * out4 = lambda.max, lambda.min, sd.cv, index of lambda.min (0-based).
 * cv_out (optional, 3*100): lambda, cvm, cvsd.  stats: lmu, jerr, sweeps, grads
 * returns status: 0 ok, 1 = sd(y)==0 shortcut, 2 = fit failed -> mean fallback.
 */
ORC_API int orc_predict_cv(const double *x, int n, int p, const double *y, const int *rows, int m,
                           const int *foldid, int nfolds, int excl, int dfmax, double thr,
                           double *mu, double *out4, double *cv_out, int *stats)
{
    out4[0] = out4[1] = out4[2] = 0; out4[3] = -1;
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (sd_all(y, n) == 0) { fill_mean(y, rows, m, n, mu); return 1; }
    enum { NL = 100 };
    double lam[NL], a0[NL], dev[NL];
    double *beta = (double *)malloc(sizeof(double) * (size_t)p * NL);
    glm_info gi;
    tl_threads = orc_outer_threads * orc_inner_threads; /* the master fit runs alone: all threads */
    glmnet_poisson(x, n, p, rows, m, y, excl, NULL, 0, dfmax, thr, 100000, lam, a0, beta, dev, &gi);
    tl_threads = 0;
    int L = gi.lmu, tot_nlp = gi.nlp, tot_gr = gi.ngrad;
    if (gi.jerr > 0 || L < 1) { free(beta); fill_mean(y, rows, m, n, mu); return 2; }
    /* fold fits on the master lambda sequence */
    double *cvraw = (double *)calloc((size_t)nfolds * L, sizeof(double)); /* per-fold mean deviance */
    double *nf = (double *)calloc(nfolds, sizeof(double));
    int bad = 0;
    const int nouter = orc_outer_threads < nfolds ? orc_outer_threads : nfolds;
    (void)nouter;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nouter) if (nouter > 1)
    for (int fo = 1; fo <= nfolds; ++fo) {
        double *fbeta = (double *)malloc(sizeof(double) * (size_t)p * L);
        int *tr = (int *)malloc(sizeof(int) * m), *te = (int *)malloc(sizeof(int) * m);
        double flam[NL], fa0[NL], fdev[NL];
        int ntr = 0, nte = 0;
        for (int i = 0; i < m; ++i) {
            if (foldid[i] == fo) te[nte++] = rows[i]; else tr[ntr++] = rows[i];
        }
        glm_info fg;
        tl_threads = nouter > 1 ? orc_inner_threads : orc_outer_threads * orc_inner_threads;
        glmnet_poisson(x, n, p, tr, ntr, y, excl, lam, L, dfmax, thr, 100000, flam, fa0, fbeta, fdev,
                       &fg);
        tl_threads = 0;
#pragma omp atomic
        tot_nlp += fg.nlp;
#pragma omp atomic
        tot_gr += fg.ngrad;
        if (fg.jerr > 0 || fg.lmu < 1) {
#pragma omp atomic write
            bad = 1;
        } else {
            nf[fo - 1] = nte;
            const int nth = nouter > 1 ? orc_inner_threads : orc_outer_threads * orc_inner_threads;
            (void)nth;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nth) if (nth > 1 && (size_t)p * m > 4000000)
            for (int l = 0; l < L; ++l) {
                int ls = l < fg.lmu ? l : fg.lmu - 1; /* short fold path: last column repeated */
                const double *bl = fbeta + (size_t)ls * p;
                ld acc = 0;
                for (int i = 0; i < nte; ++i) {
                    int r = te[i];
                    double eta = fa0[ls];
                    for (int j = 0; j < p; ++j) if (bl[j] != 0.0) eta += x[(size_t)j * n + r] * bl[j];
                    double yy = y[r];
                    double deveta = yy * eta - exp(eta);
                    double devy = yy == 0 ? 0 : yy * log(yy) - yy;
                    acc += 2 * (devy - deveta);
                }
                cvraw[(size_t)(fo - 1) * L + l] = (double)(acc / nte);
            }
        }
        free(fbeta); free(tr); free(te);
    }
    int rc = 0;
    if (bad) { fill_mean(y, rows, m, n, mu); rc = 2; }
    else {
        double cvm[NL], cvsd[NL];
        double wsum = 0;
        for (int f = 0; f < nfolds; ++f) wsum += nf[f];
        for (int l = 0; l < L; ++l) {
            double s = 0;
            for (int f = 0; f < nfolds; ++f) s += nf[f] * cvraw[(size_t)f * L + l];
            cvm[l] = s / wsum;
            double s2 = 0;
            for (int f = 0; f < nfolds; ++f) {
                double d = cvraw[(size_t)f * L + l] - cvm[l];
                s2 += nf[f] * d * d;
            }
            cvsd[l] = sqrt(s2 / wsum / (nfolds - 1));
        }
        double cvmin = INFINITY;
        for (int l = 0; l < L; ++l) if (cvm[l] < cvmin) cvmin = cvm[l];
        int imin = -1;
        for (int l = 0; l < L; ++l) if (cvm[l] <= cvmin && (imin < 0 || lam[l] > lam[imin])) imin = l;
        predict_all(x, n, p, a0[imin], beta + (size_t)imin * p, mu);
        out4[0] = lam[0];
        out4[1] = lam[imin];
        out4[2] = (cvm[0] - cvm[imin]) / cvsd[imin];
        out4[3] = imin;
        if (cv_out) {
            for (int l = 0; l < NL; ++l) {
                cv_out[l] = l < L ? lam[l] : NAN;
                cv_out[NL + l] = l < L ? cvm[l] : NAN;
                cv_out[2 * NL + l] = l < L ? cvsd[l] : NAN;
            }
        }
    }
    if (stats) { stats[0] = L; stats[1] = gi.jerr; stats[2] = tot_nlp; stats[3] = tot_gr; }
    free(beta); free(cvraw); free(nf);
    return rc;
}

/*
 * expr.predict, fixed-lambda variant (R/expr_predict.R:61-82):
 * lambda.seq = c(exp(seq(log(lmax), log(lmin), by = -0.2)), lmin); one path;
 * mu = exp(eta) at s = lmin (predict.glmnet clamps s into the fitted range).
 * returns 0 ok, 1 sd(y)==0, 2 fit failed.
 */
ORC_API int orc_predict_path(const double *x, int n, int p, const double *y, const int *rows, int m,
                             int excl, double lmax, double lmin, int dfmax, double thr, double *mu,
                             int *stats)
{
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (sd_all(y, n) == 0) { fill_mean(y, rows, m, n, mu); return 1; }
    /* seq(from, to, by): n = floor((to-from)/by + 1e-10) steps */
    double from = log(lmax), to = log(lmin);
    int nstep = (int)floor((to - from) / -0.2 + 1e-10);
    if (nstep < 0) nstep = 0;
    int nl = nstep + 2;
    double *ls = (double *)malloc(sizeof(double) * nl);
    for (int i = 0; i <= nstep; ++i) ls[i] = exp(from + i * -0.2);
    ls[nl - 1] = lmin;
    double *lam = (double *)malloc(sizeof(double) * nl), *a0 = (double *)malloc(sizeof(double) * nl);
    double *dev = (double *)malloc(sizeof(double) * nl);
    double *beta = (double *)malloc(sizeof(double) * (size_t)p * nl);
    glm_info gi;
    glmnet_poisson(x, n, p, rows, m, y, excl, ls, nl, dfmax, thr, 100000, lam, a0, beta, dev, &gi);
    int rc = 0;
    if (gi.jerr > 0 || gi.lmu < 1) { fill_mean(y, rows, m, n, mu); rc = 2; }
    else {
        /* lambda.interp(lambda, s): s clamped to [min, max] of the fitted lambdas; exact
         * hits return that column; otherwise linear interpolation in lambda. */
        int L = gi.lmu;
        double s = lmin;
        if (s > lam[0]) s = lam[0];
        if (s < lam[L - 1]) s = lam[L - 1];
        int left = L - 1, right = L - 1;
        double frac = 1.0;
        if (L > 1 && lam[0] > lam[L - 1]) {
            double sf = (lam[0] - s) / (lam[0] - lam[L - 1]);
            /* coordinates of each lambda on [0,1]; find bracketing pair */
            left = 0; right = L - 1;
            for (int l = 0; l < L; ++l) {
                double c = (lam[0] - lam[l]) / (lam[0] - lam[L - 1]);
                if (c <= sf) left = l;
            }
            for (int l = L - 1; l >= 0; --l) {
                double c = (lam[0] - lam[l]) / (lam[0] - lam[L - 1]);
                if (c >= sf) right = l;
            }
            if (lam[left] == lam[right]) frac = 1.0;
            else frac = (s - lam[right]) / (lam[left] - lam[right]);
        }
        double *b = (double *)malloc(sizeof(double) * p);
        for (int j = 0; j < p; ++j)
            b[j] = frac * beta[(size_t)left * p + j] + (1 - frac) * beta[(size_t)right * p + j];
        double ai = frac * a0[left] + (1 - frac) * a0[right];
        predict_all(x, n, p, ai, b, mu);
        free(b);
    }
    if (stats) { stats[0] = gi.lmu; stats[1] = gi.jerr; stats[2] = gi.nlp; stats[3] = gi.ngrad; }
    free(ls); free(lam); free(a0); free(dev); free(beta);
    return rc;
}
