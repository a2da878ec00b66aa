"""CPU tests of the R workspace-file exchange (saver_b200/rda.py) and of
the legacy alpha/beta object support (R/utils.R:137-232).

The writer is checked three ways: round trip through the product reader, decoding by the
independent test-side reader (oracle/rdx2.py, the one the golden fixture was made with), and --
when the reference checkout is present (this container, not the GPU box) -- the product reader
on the reference's own ``data/*.rda`` against the committed golden arrays."""
import os
import struct

import numpy as np
import pytest

from oracle import rdx2
from saver_b200 import host, rda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DATA = "/root/reference/data"


def _result(G=7, n=5, seed=0, names=True):
    rng = np.random.default_rng(seed)
    est = np.round(rng.gamma(2.0, 1.0, (G, n)), 3)
    se = np.round(rng.gamma(1.0, 0.3, (G, n)), 3) + 0.001
    info = {"size.factor": rng.gamma(5.0, 0.2, n), "maxcor": rng.random(G),
            "lambda.max": rng.random(G), "lambda.min": rng.random(G) * 0.1, "sd.cv": rng.random(G),
            "pred.time": rng.random(G), "var.time": rng.random(G), "cutoff": 0.25,
            "lambda.coefs": np.array([-3.9744, 13.0896]), "total.time": 12.5}
    gn = ["gene-%d" % i for i in range(G)] if names else None
    cn = ["cell_%d" % i for i in range(n)] if names else None
    return host.SaverResult(est, se, info, gene_names=gn, cell_names=cn)


@pytest.mark.parametrize("compress", ["gzip", "bzip2", "xz", None])
@pytest.mark.parametrize("ext", [".rda", ".rds"])
def test_saver_object_round_trip(tmp_path, compress, ext):
    res = _result()
    path = str(tmp_path / ("obj" + ext))
    rda.save_saver(path, res, compress=compress)
    back = rda.load_saver(path)
    assert isinstance(back, host.SaverResult)
    assert np.array_equal(back.estimate, res.estimate) and back.estimate.flags.c_contiguous
    assert np.array_equal(back.se, res.se)
    assert back.gene_names == res.gene_names and back.cell_names == res.cell_names
    assert list(back.info) == list(res.info)  # R's field order (R/saver_fit.R:66-70)
    for k, v in res.info.items():
        assert np.array_equal(np.atleast_1d(back.info[k]), np.atleast_1d(v)), k
    assert isinstance(back.info["cutoff"], float) and isinstance(back.info["total.time"], float)


def test_writer_is_decoded_by_the_independent_reader(tmp_path):
    res = _result(G=11, n=4, seed=3)
    path = str(tmp_path / "obj.rda")
    rda.save_saver(path, res, name="out", compress="bzip2")
    obj = rdx2.load_rda(path)["out"]
    assert obj["attr"]["class"]["value"] == ["saver"]
    fields = rdx2.as_list(obj)
    assert list(fields) == ["estimate", "se", "info"]
    est, (gn, cn) = rdx2.as_matrix(fields["estimate"])
    assert np.array_equal(est, res.estimate) and gn == res.gene_names and cn == res.cell_names
    info = rdx2.as_list(fields["info"])
    assert np.array_equal(info["maxcor"]["value"], res.info["maxcor"])
    assert info["total.time"]["attr"]["class"]["value"] == ["difftime"]
    assert info["total.time"]["attr"]["units"]["value"] == ["secs"]
    assert info["lambda.coefs"]["attr"]["names"]["value"] == ["(Intercept)", "x"]


def test_stream_layout_matches_r_serialize():
    """Byte-level known answers of R's XDR serialisation for a tiny object:
    save(x, file=..., compress=FALSE) with x <- c(a=1.5) in format version 2."""
    x = rda.RObject("double", np.array([1.5]), {"names": rda.RObject("character", ["a"])})
    buf = rda.write_rda(None, {"x": x}, compress=None)
    want = b"RDX2\nX\n" + struct.pack(">iii", 2, 0x030603, 0x020300)
    want += struct.pack(">I", 0x402)                                   # pairlist cell with a tag
    want += struct.pack(">I", 1) + struct.pack(">Ii", 0x40009, 1) + b"x"  # symbol x (ASCII CHARSXP)
    want += struct.pack(">Ii", 0x20E, 1) + struct.pack(">d", 1.5)      # REALSXP + attributes, length 1
    want += struct.pack(">I", 0x402)                                   # attribute pairlist
    want += struct.pack(">I", 1) + struct.pack(">Ii", 0x40009, 5) + b"names"
    want += struct.pack(">Ii", 0x10, 1) + struct.pack(">Ii", 0x40009, 1) + b"a"
    want += struct.pack(">I", 0xFE) * 2                                # end of attributes, end of list
    assert buf == want
    assert rda.read_rda(buf)["x"].names == ["a"]


def test_symbol_references_and_na_strings():
    a = rda.RObject("character", ["p", None, "é"], {"names": rda.RObject("character", ["u", "v", "w"])})
    b = rda.RObject("integer", np.array([1, -2 ** 31, 3], np.int32),
                    {"names": rda.RObject("character", ["u", "v", "w"])})
    buf = rda.write_rda(None, {"a": a, "b": b}, compress=None)
    assert buf.count(b"names") == 1  # the second use is a REFSXP
    back = rda.read_rda(buf)
    assert back["a"].value == ["p", None, "é"] and back["b"].names == ["u", "v", "w"]
    assert back["b"].value[1] == -2 ** 31  # NA_integer_ survives


def test_legacy_object_round_trip_and_moments(tmp_path):
    rng = np.random.default_rng(5)
    alpha = rng.gamma(3.0, 1.0, (6, 8)) + 0.01
    beta = rng.gamma(2.0, 1.0, (6, 8)) + 0.01
    est = alpha / beta
    leg = host.LegacySaverResult(est, alpha, beta, {"size.factor": np.ones(8), "a": np.arange(6.0),
                                                    "b": np.arange(6.0) * 2, "c": 1.0})
    path = str(tmp_path / "old.rda")
    rda.save_saver(path, leg)
    back = rda.load_saver(path)
    assert isinstance(back, host.LegacySaverResult)
    assert np.array_equal(back.alpha, alpha) and np.array_equal(back.beta, beta)
    # gamma mode: shape = mean^2/sd^2 = alpha, rate = mean/sd^2 = beta (what saver_cuda_sample forms)
    m, s = back.moments(False)
    assert np.allclose(m ** 2 / s ** 2, alpha, rtol=1e-13) and np.allclose(m / s ** 2, beta, rtol=1e-13)
    # correlation adjustment: mean(se^2) must be mean(alpha/beta^2)
    assert np.allclose((s ** 2).mean(1), (alpha / beta ** 2).mean(1), rtol=1e-13)
    # NB mode: mu = estimate, size = alpha
    m, s = back.moments(True)
    assert np.array_equal(m, est) and np.allclose(m ** 2 / s ** 2, alpha, rtol=1e-13)


def test_combine_saver_legacy_follows_the_old_rule():
    def part(g0, g1):
        a = np.full((g1 - g0, 3), 2.0) + np.arange(g0, g1)[:, None]
        info = {"size.factor": np.array([1.0, 2.0, 3.0]), "pred.time": np.arange(g0, g1, dtype=float),
                "var.time": np.arange(g0, g1, dtype=float) * 10, "cutoff": float(g0)}
        return host.LegacySaverResult(a / 2.0, a, np.full_like(a, 2.0), info,
                                      gene_names=["g%d" % i for i in range(g0, g1)])
    out = host.combine_saver([part(0, 2), part(2, 5)])
    assert isinstance(out, host.LegacySaverResult)
    assert out.estimate.shape == (5, 3) and out.alpha.shape == (5, 3) and out.beta.shape == (5, 3)
    assert np.array_equal(out.info["pred.time"], np.arange(5.0))       # info[[2]]: concatenated
    assert np.array_equal(out.info["var.time"], np.arange(5.0) * 10)   # info[[3]]: concatenated
    assert out.info["cutoff"] == 0.0                                   # info[[4..]]: first object's
    assert out.gene_names == ["g%d" % i for i in range(5)]


def test_combine_saver_keeps_names_of_current_objects():
    a, b = _result(G=3, seed=1), _result(G=4, seed=2)
    b.gene_names = ["other-%d" % i for i in range(4)]
    out = host.combine_saver([a, b])
    assert out.estimate.shape == (7, 5)
    assert out.gene_names == a.gene_names + b.gene_names and out.cell_names == a.cell_names
    assert out.info["total.time"] == 25.0  # summed (R/combine_saver.R:40)


def test_counts_dense_and_sparse(tmp_path):
    import scipy.sparse as sp
    rng = np.random.default_rng(9)
    x = rng.poisson(0.3, (12, 9)).astype(np.int32)
    gn, cn = ["g%d" % i for i in range(12)], ["c%d" % i for i in range(9)]
    p = str(tmp_path / "dense.rda")
    rda.save_counts(p, x, gn, cn)
    back, rn, cc = rda.load_counts(p)
    assert back.dtype == np.int32 and np.array_equal(back, x) and rn == gn and cc == cn
    p = str(tmp_path / "sparse.rds")
    rda.save_counts(p, sp.csr_matrix(x.astype(float)), gn, cn)
    m, rn, cc = rda.load_counts(p)
    assert sp.issparse(m) and m.format == "csc" and np.array_equal(m.toarray(), x) and rn == gn
    obj = rda.read_rds(p)
    assert obj.kind == "S4" and obj.rclass == ["dgCMatrix"]
    assert obj.attr["class"].attr["package"].value == ["Matrix"]
    assert np.array_equal(obj["p"].value, sp.csc_matrix(x).indptr)


def test_altrep_compact_sequences_and_wrappers():
    """Format-3 streams hold 1:n as ALTREP(compact_intseq, state = c(n, start, step))."""
    def sym(s):
        return struct.pack(">I", 1) + struct.pack(">Ii", 0x40009, len(s)) + s.encode()
    head = b"X\n" + struct.pack(">iii", 3, 0x040301, 0x030500) + struct.pack(">i", 5) + b"UTF-8"
    info = (struct.pack(">I", 2) + sym("compact_intseq") + struct.pack(">I", 2) + sym("base") +
            struct.pack(">I", 2) + struct.pack(">Iii", 13, 1, 13) + struct.pack(">I", 0xFE))
    state = struct.pack(">Ii", 14, 3) + struct.pack(">ddd", 5.0, 3.0, 1.0)
    buf = head + struct.pack(">I", 238) + info + state + struct.pack(">I", 0xFE)
    v = rda.read_rds(buf)
    assert v.kind == "integer" and np.array_equal(v.value, [3, 4, 5, 6, 7])
    bad = head + struct.pack(">I", 238) + info.replace(b"compact_intseq", b"mmap_integer00") + state + \
        struct.pack(">I", 0xFE)
    with pytest.raises(rda.RdaError, match="ALTREP"):
        rda.read_rds(bad)


def test_errors_are_loud(tmp_path):
    with pytest.raises(rda.RdaError, match="not an R workspace"):
        rda.read_rda(b"RDA2\nA\n1 2 3")
    with pytest.raises(rda.RdaError, match="XDR"):
        rda.read_rda(b"RDX2\nA\n")
    good = rda.write_rda(None, {"x": np.arange(10.0)}, compress=None)
    with pytest.raises(rda.RdaError, match="truncated"):
        rda.read_rda(good[:-20])
    with pytest.raises(rda.RdaError, match="trailing"):
        rda.read_rda(good + b"\0\0\0\0")
    with pytest.raises(rda.RdaError, match="name the"):
        rda.load_saver(rda.write_rda(None, {"a": 1.0, "b": 2.0}))
    with pytest.raises(rda.RdaError, match="not a saver object"):
        rda.load_saver(rda.write_rda(None, {"a": {"u": 1.0}}))
    closure = b"X\n" + struct.pack(">iii", 2, 0x030603, 0x020300) + struct.pack(">I", 3)
    with pytest.raises(rda.RdaError, match="unsupported SEXP type 3"):
        rda.read_rds(closure)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DATA, "linnarsson_saver.rda")),
                    reason="reference checkout not present (GPU box)")
def test_reads_the_reference_data_files():
    g = np.load(os.path.join(ROOT, "tests", "golden", "linnarsson.npz"))
    res = rda.load_saver(os.path.join(REF_DATA, "linnarsson_saver.rda"))
    assert isinstance(res, host.SaverResult)
    assert np.array_equal(np.rint(res.estimate * 1000).astype(np.int64), g["estimate_m"])
    assert np.array_equal(np.rint(res.se * 1000).astype(np.int64), g["se_m"])
    assert res.gene_names == list(g["gene_names"]) and res.cell_names == list(g["cell_names"])
    assert list(res.info) == list(rda._INFO_KEYS)
    for k, gk in (("size.factor", "size_factor"), ("maxcor", "maxcor"), ("lambda.max", "lambda_max"),
                  ("lambda.min", "lambda_min"), ("sd.cv", "sd_cv"), ("lambda.coefs", "lambda_coefs")):
        assert np.array_equal(res.info[k], g[gk], equal_nan=True), k
    assert res.info["cutoff"] == float(g["cutoff"][0])
    x, gn, cn = rda.load_counts(os.path.join(REF_DATA, "linnarsson.rda"))
    assert np.array_equal(x, g["counts"]) and gn == list(g["gene_names"]) and cn == list(g["cell_names"])
    # and the reference's object survives a write -> read cycle bit for bit
    back = rda.load_saver(rda.write_rda(None, {"o": rda.read_rda(
        os.path.join(REF_DATA, "linnarsson_saver.rda"))["linnarsson_saver"]}, compress=None))
    assert np.array_equal(back.estimate, res.estimate) and np.array_equal(back.se, res.se)
