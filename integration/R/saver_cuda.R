# saver_cuda.R -- the R side of the drop-in: the bodies a maintainer of mohuangx/SAVER swaps in.
# Signatures are the reference's (NAMESPACE:3-23); everything numerical is one .Call into
# libsaver_cuda (src/rcall_shim.c).  No R in the build image (DESIGN.md 1b): the .Call sites below are
# checked against the shim's registration table, and the shim itself runs on a fake R runtime in
# tests/test_r_shim.py.

.saver.cuda <- new.env()
saver.cuda.handle <- function() {
  if (is.null(.saver.cuda$h))
    # options(saver.cuda.devices = 0:7): one R process, eight GPUs (blocks of calc.estimate are split
    # over the devices inside the library); default: the single device saver.cuda.device
    .saver.cuda$h <- .Call(C_saver_cuda_create,
                           as.integer(getOption("saver.cuda.devices", getOption("saver.cuda.device", 0L))))
  .saver.cuda$h
}

# cv.glmnet's fold draw, done by the host so folds are identical to the reference's
# (R/expr_predict.R:35-36: set.seed(seed); cv.glmnet then calls sample(rep(seq(nfolds), length = N)))
.draw.folds <- function(seeds, pred.cells, ncells, nfolds = 5L) {
  fold <- matrix(0L, ncells, length(seeds))
  for (i in seq_along(seeds)) {
    set.seed(seeds[i])
    fold[pred.cells, i] <- sample(rep(seq(nfolds), length = length(pred.cells)))
  }
  fold
}

# replaces R/calc_estimate.R:65-161 (same arguments, same 8-element result)
calc.estimate <- function(x, x.est, cutoff = 0, coefs = NULL, sf, scale.sf, pred.gene.names,
                          pred.cells, null.model, nworkers, calc.maxcor, estimates.only) {
  h <- saver.cuda.handle()
  x <- as.matrix(x)
  storage.mode(x) <- "double"                       # integer count matrices are common
  storage.mode(x.est) <- "double"
  if (is.null(rownames(x))) rownames(x) <- seq_len(nrow(x))   # clean.data's default names
  if (!identical(.saver.cuda$design, x.est)) {      # x.est is installed once per saver() call
    .Call(C_saver_cuda_set_design, h, x.est)
    .saver.cuda$design <- x.est
  }
  ncells <- ncol(x)
  self.col <- match(rownames(x), colnames(x.est), nomatch = 0L) - 1L   # 0-based, -1 = none
  is.pred <- as.integer(rownames(x) %in% pred.gene.names)
  mask <- integer(ncells); mask[pred.cells] <- 1L
  mode <- if (null.model) 0L else if (is.null(coefs)) 1L else 2L
  foldid <- if (mode == 1L) .draw.folds(seq_len(nrow(x)), pred.cells, ncells) else NULL
  out <- .Call(C_saver_cuda_calc_estimate, h, t(x), as.numeric(sf), as.numeric(scale.sf),
               as.integer(self.col), is.pred, mask, foldid, mode, as.logical(calc.maxcor),
               as.numeric(cutoff), if (is.null(coefs)) NULL else as.numeric(coefs),
               as.logical(estimates.only))
  info <- out[[3]]
  list(est = t(out[[1]]), se = if (estimates.only) NA else t(out[[2]]), maxcor = info[, 1],
       lambda.max = info[, 2], lambda.min = info[, 3], sd.cv = info[, 4], ct = info[, 5],
       vt = info[, 6])
}

# replaces R/calc_posterior.R:26-68, R/optimize_variance.R:24-134, R/calc_loglik.R:26-75
calc.post <- function(y, mu, sf, scale.sf) {
  out <- .Call(C_saver_cuda_calc_post, saver.cuda.handle(), as.numeric(y), as.numeric(mu),
               rep_len(as.numeric(sf), length(y)), as.numeric(scale.sf))
  list(estimate = out[[1]], se = out[[2]])
}
# calc.a / calc.b / calc.k.  Default: the library's deterministic value (the expectation of the
# reference's Monte-Carlo mean).  options(saver.cuda.r.sample = TRUE) reproduces the reference's own
# stochastic branch instead (R/optimize_variance.R:46-55,86-95,123-132): the device returns the 100
# grid points and their log-likelihoods, sample() runs here on R's RNG stream, and the likelihood
# at the drawn value is one more device evaluation.
.calc.var <- function(model, loglik.fn, y, mu, sf) {
  y <- as.numeric(y); mu <- as.numeric(mu); sf <- as.numeric(sf)
  if (!isTRUE(getOption("saver.cuda.r.sample", FALSE)))
    return(.Call(C_saver_cuda_calc_var, saver.cuda.handle(), model, y, mu, sf))
  out <- .Call(C_saver_cuda_calc_var_grid, saver.cuda.handle(), model, y, mu, sf)
  if (bitwAnd(out[[2]], 1L) == 0L || bitwAnd(out[[2]], 4L) != 0L) return(out[[1]])
  samp <- out[[3]][1:100]; ll <- out[[3]][101:200]     # ll = log-likelihood = -calc.loglik.*()
  w <- exp(ll - min(ll)); w <- w / sum(w)
  p <- mean(sample(samp, 10000, replace = TRUE, prob = w))
  c(if (model == 2L) p else 1 / p, loglik.fn(p, y, mu, sf))
}
calc.a <- function(y, mu, sf) .calc.var(0L, calc.loglik.a, y, mu, sf)
calc.b <- function(y, mu, sf) .calc.var(1L, calc.loglik.b, y, mu, sf)
calc.k <- function(y, mu, sf) .calc.var(2L, calc.loglik.k, y, mu, sf)
calc.loglik.a <- function(a, y, mu, sf) .Call(C_saver_cuda_loglik, saver.cuda.handle(), 0L, a, as.numeric(y), as.numeric(mu), as.numeric(sf))
calc.loglik.b <- function(b, y, mu, sf) .Call(C_saver_cuda_loglik, saver.cuda.handle(), 1L, b, as.numeric(y), as.numeric(mu), as.numeric(sf))
calc.loglik.k <- function(k, y, mu, sf) .Call(C_saver_cuda_loglik, saver.cuda.handle(), 2L, k, as.numeric(y), as.numeric(mu), as.numeric(sf))

# replaces R/calc_maxcor.R:26-42
calc.maxcor <- function(x1, x2) {
  h <- saver.cuda.handle()
  storage.mode(x1) <- "double"
  storage.mode(x2) <- "double"
  .Call(C_saver_cuda_set_design, h, x1)
  .saver.cuda$design <- x1                          # the resident design changed: keep the cache honest
  self.col <- if (is.null(colnames(x2)) || is.null(colnames(x1))) rep(-1L, ncol(x2))
              else match(colnames(x2), colnames(x1), nomatch = 0L) - 1L
  .Call(C_saver_cuda_maxcor, h, x2, as.integer(self.col))
}

# Legacy objects list(estimate, alpha, beta, info) (R/utils.R:137-232: sample.saver.old,
# cor.genes.old, cor.cells.old) run on the same kernels through equivalent (mean, sd) panels:
# gamma draws and the correlation adjustment use mean = alpha/beta, sd = sqrt(alpha)/beta
# (shape = mean^2/sd^2 = alpha, rate = mean/sd^2 = beta, sd^2 = alpha/beta^2); NB draws use
# mean = estimate, sd = estimate/sqrt(alpha) (size = alpha).
.saver.moments <- function(x, efficiency.known = FALSE) {
  if (is.null(x$alpha)) return(list(est = x$estimate, se = x$se))
  if (efficiency.known) return(list(est = x$estimate, se = x$estimate / sqrt(x$alpha)))
  list(est = x$alpha / x$beta, se = sqrt(x$alpha) / x$beta)
}

# replaces R/sample_saver.R:32-79 and R/cor_adjust.R:26-66
sample.saver <- function(x, rep = 1, efficiency.known = FALSE, seed = NULL) {
  rep <- as.integer(rep)
  if (rep <= 0) stop("rep must be a positive integer.")
  if (is.null(seed)) seed <- sample.int(.Machine$integer.max, 1)
  m <- .saver.moments(x, efficiency.known)
  draw <- function(j) {
    s <- t(.Call(C_saver_cuda_sample, saver.cuda.handle(), t(m$est), t(m$se),
                 as.logical(efficiency.known), as.numeric(seed), as.integer(j - 1L)))
    dimnames(s) <- dimnames(x$estimate)
    s
  }
  if (rep == 1) draw(1L) else lapply(seq_len(rep), draw)
}
cor.genes <- function(x, cor.mat = NULL)
  .Call(C_saver_cuda_cor_adjust, saver.cuda.handle(), t(x$estimate), t(.saver.moments(x)$se), FALSE, cor.mat)
cor.cells <- function(x, cor.mat = NULL)
  .Call(C_saver_cuda_cor_adjust, saver.cuda.handle(), t(x$estimate), t(.saver.moments(x)$se), TRUE, cor.mat)

# sparse input (dgCMatrix): upload the slots once; row blocks and x.est are densified on the device.
# In saver() (R/saver.R:156-157) and calc.estimate (R/calc_estimate.R:69) replace
#   x.est <- t(as.matrix(log(sweep(x[good.genes, ] + 1, 2, sf, "/"))))   by   saver.cuda.log.rows(good.genes, sf)
#   as.matrix(x[block, ])                                                by   t(saver.cuda.rows(block))
saver.cuda.set.counts <- function(x) {
  stopifnot(methods::is(x, "dgCMatrix"))
  .Call(C_saver_cuda_set_counts_csc, saver.cuda.handle(), x@p, x@i, as.numeric(x@x), dim(x))
  sums <- .Call(C_saver_cuda_csc_sums, saver.cuda.handle(), dim(x))
  list(col.sums = sums[[1]], row.sums = sums[[2]], ncells = ncol(x))
}
saver.cuda.rows <- function(genes, ncells)
  .Call(C_saver_cuda_csc_gather, saver.cuda.handle(), as.integer(genes - 1L), NULL, ncells)
saver.cuda.log.rows <- function(genes, sf)
  .Call(C_saver_cuda_csc_gather, saver.cuda.handle(), as.integer(genes - 1L), as.numeric(sf),
        length(sf))

# dense input (matrix): one upload of t(x); colSums / rowMeans / x.est come back from the device, so
# saver() (R/saver.R:105-157) never makes its three host passes over the whole matrix:
#   d <- saver.cuda.set.dense(x); sf <- d$col.sums / mean(d$col.sums)
#   good.genes <- which(d$row.sums / ncol(x) >= 0.1); x.est <- saver.cuda.dense.log.rows(good.genes, sf)
saver.cuda.set.dense <- function(x) {
  .Call(C_saver_cuda_set_counts_dense, saver.cuda.handle(), t(x))
  sums <- .Call(C_saver_cuda_dense_sums, saver.cuda.handle(), dim(x))
  list(col.sums = sums[[1]], row.sums = sums[[2]], ncells = ncol(x))
}
saver.cuda.dense.log.rows <- function(genes, sf)
  .Call(C_saver_cuda_dense_gather, saver.cuda.handle(), as.integer(genes - 1L), as.numeric(sf),
        length(sf))

# saver(), saver.fit(), saver.fit.mean/.null, get.mu, combine.saver and R/utils.R stay as they are
# in the reference (they only call the functions above); saver() forces ncores to one process:
#   if (ncores > 1) message("libsaver_cuda: device parallelism replaces ", ncores, " PSOCK workers")
