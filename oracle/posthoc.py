"""numpy restatement of the reference's post-hoc functions (TEST INFRASTRUCTURE ONLY).

    cor.genes / cor.cells   /root/reference/R/cor_adjust.R:26-66
    get.mu                  /root/reference/R/get_mu.R:22-25
    sample.saver moments    /root/reference/R/sample_saver.R:50-57 (distribution parameters only:
                            R's Mersenne-Twister stream is not reproducible outside R)
"""
import numpy as np


def cor_adjust(est, se, by_cells=False):
    """est, se: (G, n).  cor(t(est)) [or cor(est)] * outer(adj, adj)."""
    e = est.T if by_cells else est          # rows = the vectors being correlated
    s = se.T if by_cells else se
    cor = np.corrcoef(e)
    var = e.var(1, ddof=1)
    adj = np.sqrt(var / (var + (s ** 2).mean(1)))
    return cor * np.outer(adj, adj)


def get_mu(x, est, se, size_factor):
    return np.round((est ** 2 / se ** 2 - x) / (est / se ** 2 - size_factor), 3)


def sample_params(est, se):
    """(shape, rate) of the gamma draw and (mu, size) of the negative-binomial draw."""
    shape = est ** 2 / se ** 2
    return shape, est / se ** 2
