"""saver_b200 -- B200-native implementation of SAVER's expression-recovery hot path.

Host-side mirror of the reference's R function API over the C ABI of libsaver_cuda
(include/saver_cuda.h).  No CPU fallback: every numerical entry point runs hand-written
sm_100a CUDA and raises if the library or a CUDA device is missing.
"""
from . import host, ops, rda  # noqa: F401
from .host import (LegacySaverResult, SaverError, SaverResult, combine_saver, cor_cells, cor_genes, cor_genes_sharded, get_mu,  # noqa: F401
                   sample_saver, saver, saver_fit, saver_fit_mean, saver_fit_null)
from ._lib import Handle, SaverCudaError, default_handle  # noqa: F401

__version__ = "0.1.0"
