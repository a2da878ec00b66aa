"""R's default random number stream, restated: ``set.seed(s)`` + ``sample()``.

The reference draws two things from R's RNG: the gene order of saver.fit
(``sample(1:npred, npred)``, R/saver_fit.R:87-92) and, inside ``cv.glmnet``, the fold assignment
``sample(rep(seq(nfolds), length = N))`` after ``set.seed(j)`` with ``j`` the gene's position in
its stage block (R/expr_predict.R:35-36, R/calc_estimate.R:98).  An R host does both itself
(INTEGRATION.md); this module lets the Python host draw *the same* folds, so a run here can be
compared with -- and the cross-validation glue pinned against -- results stored by R
(``tests/test_oracle_golden.py``: with these folds the CPU oracle reproduces ``lambda.min`` and
``sd.cv`` of the reference's stored linnarsson object).

Restated from R's published sources (src/main/RNG.c, src/main/random.c):

* ``set.seed(seed)`` for the default Mersenne-Twister: scramble ``seed = 69069 * seed + 1``
  (mod 2^32) fifty times, then fill 625 words with the same recurrence; word 0 is the position
  ``mti`` and is forced to 624, so the first draw regenerates the whole state.
* ``unif_rand()`` = MT19937 output * 2.3283064365386963e-10, clamped into (0, 1).
* ``sample.int(n)`` without replacement: ``for i: j = R_unif_index(n); y[i] = x[j]; x[j] = x[--n]``.
  ``R_unif_index(dn)`` is ``floor(dn * unif_rand())`` up to R 3.5.x (``sample_kind="rounding"``) and
  rejection sampling on ``ceil(log2(dn))`` random bits, assembled from 16-bit chunks, from R 3.6.0
  (``sample_kind="rejection"``).
"""
import numpy as np

_I2_32M1 = 2.328306437080797e-10  # R's fixup() bounds
_SCALE = 2.3283064365386963e-10   # 2^-32


class RRng:
    def __init__(self, seed, sample_kind="rejection"):
        if sample_kind not in ("rejection", "rounding"):
            raise ValueError("sample_kind must be 'rejection' (R >= 3.6) or 'rounding' (R < 3.6)")
        self.sample_kind = sample_kind
        s = int(seed) & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        words = np.empty(625, dtype=np.uint32)
        for j in range(625):
            s = (69069 * s + 1) & 0xFFFFFFFF
            words[j] = s
        self._bg = np.random.MT19937()
        st = self._bg.state
        st["state"]["key"] = words[1:].copy()
        st["state"]["pos"] = 624
        self._bg.state = st

    def unif_rand(self, size=None):
        raw = self._bg.random_raw(size)
        u = np.asarray(raw, dtype=np.float64) * _SCALE
        u = np.where(u <= 0.0, 0.5 * _I2_32M1, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * _I2_32M1, u)
        return float(u) if size is None else u

    def _rbits(self, bits):
        v = 0
        n = 0
        while n <= bits:
            v = 65536 * v + int(np.floor(self.unif_rand() * 65536))
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return v

    def unif_index(self, dn):
        if self.sample_kind == "rounding":
            return int(np.floor(dn * self.unif_rand()))
        if dn <= 0:
            return 0
        bits = int(np.ceil(np.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return dv

    def sample_int(self, n, size=None):
        """``sample.int(n, size)`` without replacement, 1-based like R."""
        k = n if size is None else int(size)
        x = np.arange(n, dtype=np.int64)
        y = np.empty(k, dtype=np.int64)
        m = n
        if self.sample_kind == "rounding":  # one uniform per draw: vectorise the stream
            u = self.unif_rand(k)
            for i in range(k):
                j = int(np.floor(m * u[i]))
                y[i] = x[j] + 1
                m -= 1
                x[j] = x[m]
            return y
        for i in range(k):
            j = self.unif_index(m)
            y[i] = x[j] + 1
            m -= 1
            x[j] = x[m]
        return y

    def sample(self, x):
        """``sample(x)``: a random permutation of the vector x."""
        x = np.asarray(x)
        return x[self.sample_int(x.size) - 1]


def cv_folds(seed, N, nfolds=5, sample_kind="rejection"):
    """cv.glmnet's ``foldid`` after ``set.seed(seed)``: sample(rep(seq(nfolds), length = N))."""
    base = (np.arange(N) % nfolds + 1).astype(np.int32)
    return RRng(seed, sample_kind).sample(base).astype(np.int32)
