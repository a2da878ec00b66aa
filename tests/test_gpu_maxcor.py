"""GPU parity: design standardisation + FP64 DMMA contraction + calc.maxcor vs golden/oracle."""
import numpy as np
import pytest

import oracle
from saver_b200 import ops

pytestmark = pytest.mark.gpu


def test_maxcor_matches_golden(golden, handle):
    """T1: info$maxcor of the reference's stored object, all 3,529 genes."""
    ops.set_design(golden["xest"], handle=handle)
    G = golden["x"].shape[0]
    mc = ops.maxcor(golden["y"], self_col=np.arange(G), handle=handle)
    assert np.abs(mc - golden["maxcor"]).max() < 1e-12


def test_maxcor_ragged_shapes_and_constant_columns(handle):
    """n, p not multiples of the GEMM tile; constant design column and constant gene -> 0."""
    rng = np.random.default_rng(5)
    n, p, B = 333, 201, 70
    x = rng.standard_normal((p, n)) + rng.standard_normal((p, 1))
    x[17] = 2.5                       # constant predictor: cor is NA -> 0
    Y = np.exp(rng.standard_normal((B, n)))
    Y[3] = 1.0                        # sd(y) == 0
    Y[4] = x[40] * 2 + 1              # perfectly correlated with column 40
    self_col = np.full(B, -1, dtype=np.int32)
    self_col[5] = 12
    ops.set_design(x, handle=handle)
    got = ops.maxcor(Y, self_col=self_col, handle=handle)
    ref = oracle.maxcor(x, Y, self_col=self_col)
    assert np.abs(got - ref).max() < 1e-12
    assert got[3] == 0.0 and abs(got[4] - 1.0) < 1e-12
    # no self exclusion when self_col is None
    got2 = ops.maxcor(Y[:9], handle=handle)
    assert np.abs(got2 - oracle.maxcor(x, Y[:9])).max() < 1e-12
