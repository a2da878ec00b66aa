"""Synthetic negative-binomial count matrices for the benchmark configs.

K = 12 latent factors, gene loadings N(0, 0.5^2) with 30 % of genes loading-free, log base
mean N(-1, 1.5^2), cell depth LogNormal(0, 0.35^2), dispersion Gamma(2, 1) clipped to
[0.1, 50]; counts ~ NB(mean = depth * rate, size = dispersion), int32, genes x cells.
"""
import numpy as np


def nb_counts(n_genes, n_cells, seed, k=12, gene_stream=0):
    """gene_stream > 0 draws a different block of genes over the SAME cells (same latent factors
    and depths), which is how bench.py gives every GPU its own block against one shared design."""
    cell_rng = np.random.default_rng([seed, 0])
    Z = cell_rng.standard_normal((n_cells, k))
    depth = cell_rng.lognormal(0.0, 0.35, n_cells)
    rng = np.random.default_rng([seed, 1 + gene_stream])
    W = rng.normal(0.0, 0.5, (n_genes, k))
    W[rng.random(n_genes) < 0.3] = 0.0
    b = rng.normal(-1.0, 1.5, n_genes)
    phi = np.clip(rng.gamma(2.0, 1.0, n_genes), 0.1, 50.0)
    out = np.empty((n_genes, n_cells), dtype=np.int32)
    step = max(1, (1 << 24) // max(n_cells, 1))
    for g0 in range(0, n_genes, step):
        g1 = min(n_genes, g0 + step)
        rate = np.exp(b[g0:g1, None] + W[g0:g1] @ Z.T) * depth[None, :]
        lam = rng.gamma(phi[g0:g1, None], rate / phi[g0:g1, None])
        out[g0:g1] = rng.poisson(lam)
    return out


def fold_ids(n_genes, m, nfolds, seed):
    """Per-gene CV fold assignment: a permutation of rep(1..nfolds, length = m) per gene
    (what cv.glmnet draws after set.seed(j), R/expr_predict.R:35-43), from a Philox stream."""
    rng = np.random.Generator(np.random.Philox(seed))
    base = (np.arange(m) % nfolds + 1).astype(np.int32)
    out = np.empty((n_genes, m), dtype=np.int32)
    for g in range(n_genes):
        out[g] = rng.permutation(base)
    return out
