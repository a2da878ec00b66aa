"""GPU property tests: edge branches of calc.post / calc.loglik on random inputs,
checked against the CPU oracle.  Branches: sum(y) == 0 (R/calc_posterior.R:34-36), var(mu) == 0
(:37-39), single non-zero cell, huge counts, scalar recycling (R/calc_loglik.R:28-33)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle
from saver_b200 import ops

pytestmark = pytest.mark.gpu
SET = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(8, 300), kind=st.sampled_from(["poisson", "sparse", "huge", "one"]),
       const_mu=st.booleans())
def test_calc_post_matches_oracle_on_random_genes(handle, seed, n, kind, const_mu):
    rng = np.random.default_rng(seed)
    sf = np.exp(0.4 * rng.standard_normal(n))
    sf /= sf.mean()
    base = float(np.exp(rng.uniform(-3, 4)))
    mu = np.full(n, base) if const_mu else base * np.exp(0.5 * rng.standard_normal(n))
    if kind == "poisson":
        y = rng.poisson(mu * sf).astype(np.float64)
    elif kind == "sparse":
        y = rng.poisson(mu * sf * 0.05).astype(np.float64)
    elif kind == "huge":
        y = rng.poisson(mu * sf * 3e4).astype(np.float64)
    else:
        y = np.zeros(n)
        y[rng.integers(n)] = float(rng.integers(1, 50))
    got = ops.calc_post(y[None], mu[None], sf, 1.0, want_raw=True, handle=handle)
    ref = oracle.calc_post(y, mu, sf, 1.0)
    if ref["rc"] != 0:
        return  # R itself would stop() here (no finite model): nothing to compare
    assert got["info"][0, 0] == ref["info"][0], "branch / chosen model"
    if y.sum() == 0:
        assert (got["est"] == 0).all() and (got["se"] == 0).all()
        return
    assert int(got["info"][0, 5]) == int(ref["info"][5]), "fallback branch mask"
    assert _rel(got["est_raw"][0], ref["est_raw"]).max() < 1e-6
    assert _rel(got["se_raw"][0], ref["se_raw"]).max() < 1e-6


@settings(**SET)
@given(seed=st.integers(0, 2**31 - 1), n=st.integers(2, 200), model=st.integers(0, 2),
       logp=st.floats(-4, 4))
def test_loglik_matches_oracle_and_recycles_scalars(handle, seed, n, model, logp):
    rng = np.random.default_rng(seed)
    y = rng.poisson(3.0, n).astype(np.float64)
    mu = np.exp(rng.uniform(-1, 2, n))
    sf = np.exp(0.2 * rng.standard_normal(n))
    prm = float(np.exp(logp))
    got = ops.loglik(model, prm, y, mu, sf, handle=handle)
    ref = oracle.loglik(model, prm, y, mu, sf)
    assert abs(got - ref) <= 1e-10 * max(1.0, abs(ref))
    # scalar mu / sf are recycled
    got_s = ops.loglik(model, prm, y, 1.7, 1.0, handle=handle)
    ref_s = oracle.loglik(model, prm, y, np.full(n, 1.7), np.ones(n))
    assert abs(got_s - ref_s) <= 1e-10 * max(1.0, abs(ref_s))
