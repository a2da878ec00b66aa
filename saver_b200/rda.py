"""On-disk exchange of ``saver`` objects with R sessions: ``save()`` / ``load()`` files (.rda).

SURVEY.md section 8(f) rank 4.  The reference keeps its data and results as R workspace files
(``data/linnarsson.rda``, ``data/linnarsson_saver.rda``, documented in R/data.R:10,21) and users
pass ``saver`` objects between sessions with ``save()`` / ``saveRDS()``.  This module reads and
writes that container without R, so a Python harness and a real R session can hand each other

* count matrices (dense ``matrix`` or Matrix's ``dgCMatrix``),
* ``saver`` objects in the current layout ``list(estimate, se, info)`` (R/saver.R:173-175, fields
  R/saver_fit.R:66-70), and
* ``saver`` objects in the legacy layout ``list(estimate, alpha, beta, info)`` that the
  reference still accepts in sample.saver / cor.genes / cor.cells / combine.saver
  (R/utils.R:137-232) -> ``host.LegacySaverResult``.

Container: optional gzip / bzip2 / xz compression around ``RDX2\\n`` (or ``RDX3\\n``) + ``X\\n``
+ version words, then one pairlist (symbol -> value); ``saveRDS`` files are the same stream
without the ``RDXn`` magic and with a single value.  Everything is XDR (big endian).  An item
starts with an int32 of flags: low byte = SEXP type, 0x100 = is an object, 0x200 = has
attributes, 0x400 = has a tag; CHARSXP carries its encoding in the ``gp`` bits (flags >> 12).
Only the vector types an expression matrix or a ``saver`` object is made of are supported
(NULL, symbol, pairlist, logical, integer, double, character, list, S4 instance); anything else
(closures, environments, ALTREP, ...) raises ``RdaError``.
"""
import bz2
import gzip
import io
import lzma
import struct

import numpy as np

__all__ = ["RdaError", "RObject", "read_rda", "read_rds", "write_rda", "write_rds", "to_python",
           "matrix", "named_list", "load_saver", "save_saver", "load_counts", "save_counts"]

_SYM, _LIST, _CHAR, _LGL, _INT, _REAL, _STR, _VEC, _S4 = 1, 2, 9, 10, 13, 14, 16, 19, 25
_NIL, _REF, _ALTREP = 254, 255, 238
_IS_OBJ, _HAS_ATTR, _HAS_TAG = 0x100, 0x200, 0x400
_ASCII, _UTF8, _LATIN1 = 64 << 12, 8 << 12, 4 << 12
_NA_INT = -2 ** 31
_WRITER_VERSION = (3 << 16) | (6 << 8) | 3     # the stream claims R 3.6.3 ...
_MIN_READER = (2 << 16) | (3 << 8) | 0         # ... and is readable from R 2.3.0 (format 2)


class RdaError(ValueError):
    pass


class RObject:
    """A decoded R value that is not a plain scalar container: ``kind`` is one of 'logical',
    'integer', 'double', 'character', 'list', 'S4'; ``value`` a numpy array (atomic kinds), a
    Python list (character: str or None for NA; list: decoded members) or None (S4); ``attr`` an
    ordered dict of attributes (names, dim, dimnames, class, slots of an S4 instance, ...)."""

    __slots__ = ("kind", "value", "attr")

    def __init__(self, kind, value, attr=None):
        self.kind = kind
        self.value = value
        self.attr = dict(attr or {})

    def __repr__(self):
        n = "" if self.value is None else " x%d" % len(self.value)
        return "<R %s%s %s>" % (self.kind, n, list(self.attr))

    # -- conveniences ------------------------------------------------------------------------
    @property
    def names(self):
        nm = self.attr.get("names")
        return None if nm is None else list(nm.value)

    @property
    def rclass(self):
        c = self.attr.get("class")
        return [] if c is None else list(c.value)

    def __getitem__(self, key):
        """``x$key`` for a named list / ``x@key`` for an S4 instance."""
        if self.kind == "S4":
            return self.attr[key]
        nm = self.names
        if self.kind != "list" or nm is None or key not in nm:
            raise KeyError(key)
        return self.value[nm.index(key)]

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default


_KIND = {_LGL: "logical", _INT: "integer", _REAL: "double", _STR: "character", _VEC: "list"}


# ----------------------------------------------------------------------------------------------
# reading
# ----------------------------------------------------------------------------------------------
def _decompress(raw):
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


class _Reader:
    def __init__(self, buf, pos):
        self.buf = buf
        self.pos = pos
        self.symbols = []

    def i32(self):
        if self.pos + 4 > len(self.buf):
            raise RdaError("truncated stream at byte %d" % self.pos)
        v = struct.unpack_from(">i", self.buf, self.pos)[0]
        self.pos += 4
        return v

    def length(self):
        n = self.i32()
        if n == -1:  # long vector: two more words
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        if n < 0:
            raise RdaError("negative vector length at byte %d" % self.pos)
        return n

    def array(self, dtype, n):
        size = np.dtype(dtype).itemsize * n
        if self.pos + size > len(self.buf):
            raise RdaError("truncated vector at byte %d" % self.pos)
        a = np.frombuffer(self.buf, dtype, n, self.pos)
        self.pos += size
        return a

    def header(self):
        fmt = self.i32()
        self.i32()  # writer version
        self.i32()  # minimal reader version
        if fmt == 3:  # native encoding name
            n = self.i32()
            self.pos += n
        elif fmt != 2:
            raise RdaError("unsupported serialisation format %d" % fmt)

    def item(self):
        return self.decode(self.i32())

    def attributes(self):
        pl = self.item()
        if pl is None:
            return {}
        return {tag: val for tag, val in pl}

    def decode(self, flags):
        t = flags & 0xFF
        if t == _NIL:
            return None
        if t == _REF:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.symbols[idx - 1]
        if t == _SYM:
            name = self.item()
            self.symbols.append(name)
            return name
        if t == _LIST:
            out = []
            while True:
                attr = self.attributes() if flags & _HAS_ATTR else None
                tag = self.item() if flags & _HAS_TAG else None
                out.append((tag, self.item()))
                del attr
                flags = self.i32()
                if (flags & 0xFF) == _NIL:
                    return out
                if (flags & 0xFF) != _LIST:
                    raise RdaError("pairlist continues with SEXP type %d" % (flags & 0xFF))
        if t == _CHAR:
            n = self.i32()
            if n == -1:
                return None  # NA_character_
            raw = bytes(self.buf[self.pos:self.pos + n])
            self.pos += n
            return raw.decode("latin1" if flags & _LATIN1 else "utf-8", errors="replace")
        if t == _S4:
            return RObject("S4", None, self.attributes() if flags & _HAS_ATTR else {})
        if t == _ALTREP:
            return self.altrep()
        if t not in _KIND:
            raise RdaError("unsupported SEXP type %d at byte %d" % (t, self.pos))
        n = self.length()
        if t in (_LGL, _INT):
            v = self.array(">i4", n).astype(np.int32)
        elif t == _REAL:
            v = self.array(">f8", n).astype(np.float64)
        else:
            v = [self.item() for _ in range(n)]
        attr = self.attributes() if flags & _HAS_ATTR else {}
        return RObject(_KIND[t], v, attr)


def _altrep(self):
    """The compact / wrapper ALTREP classes base R writes in format 3 (``1:n`` sequences and
    attribute wrappers); anything else has no portable state and is refused."""
    info, state, attr = self.item(), self.item(), self.item()
    cls = info[0][1] if info else None
    attr = {} if attr is None else {tag: val for tag, val in attr}
    if cls in ("compact_intseq", "compact_realseq") and isinstance(state, RObject):
        n, start, step = (float(v) for v in state.value[:3])
        seq = start + step * np.arange(int(n), dtype=np.float64)
        if cls == "compact_intseq":
            return RObject("integer", seq.astype(np.int32), attr)
        return RObject("double", seq, attr)
    if cls is not None and cls.startswith("wrap_") and isinstance(state, RObject) and state.value:
        inner = state.value[0]
        if isinstance(inner, RObject):
            merged = dict(inner.attr)
            merged.update(attr)
            return RObject(inner.kind, inner.value, merged)
    raise RdaError("ALTREP class %r is not supported (materialise the vector before saving)" % cls)


_Reader.altrep = _altrep


def _open(path_or_bytes):
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        return _decompress(bytes(path_or_bytes))
    with open(path_or_bytes, "rb") as fh:
        return _decompress(fh.read())


def read_rda(path_or_bytes):
    """``load(file)``: {object name: decoded value} of a ``save()`` file."""
    buf = _open(path_or_bytes)
    if buf[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise RdaError("not an R workspace file (magic %r); ascii and pre-2.0 formats are not "
                       "supported" % buf[:5])
    if buf[5:7] != b"X\n":
        raise RdaError("only the XDR binary serialisation is supported (got %r)" % buf[5:7])
    rd = _Reader(buf, 7)
    rd.header()
    pl = rd.item()
    if rd.pos != len(buf):
        raise RdaError("trailing bytes: consumed %d of %d" % (rd.pos, len(buf)))
    return {} if pl is None else {tag: val for tag, val in pl}


def read_rds(path_or_bytes):
    """``readRDS(file)``: the single decoded value of a ``saveRDS()`` file."""
    buf = _open(path_or_bytes)
    if buf[:2] != b"X\n":
        raise RdaError("only the XDR binary serialisation is supported (got %r)" % buf[:2])
    rd = _Reader(buf, 2)
    rd.header()
    val = rd.item()
    if rd.pos != len(buf):
        raise RdaError("trailing bytes: consumed %d of %d" % (rd.pos, len(buf)))
    return val


# ----------------------------------------------------------------------------------------------
# writing
# ----------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.out = io.BytesIO()
        self.symbols = {}

    def i32(self, v):
        self.out.write(struct.pack(">i", v))

    def flags(self, v):
        self.out.write(struct.pack(">I", v & 0xFFFFFFFF))

    def header(self):
        self.i32(2)
        self.i32(_WRITER_VERSION)
        self.i32(_MIN_READER)

    def charsxp(self, s):
        if s is None:
            self.flags(_CHAR)
            self.i32(-1)
            return
        try:
            raw = s.encode("ascii")
            enc = _ASCII
        except UnicodeEncodeError:
            raw = s.encode("utf-8")
            enc = _UTF8
        self.flags(_CHAR | enc)
        self.i32(len(raw))
        self.out.write(raw)

    def symbol(self, name):
        idx = self.symbols.get(name)
        if idx is not None:
            self.flags(_REF | (idx << 8))
            return
        self.symbols[name] = len(self.symbols) + 1
        self.flags(_SYM)
        self.charsxp(name)

    def pairlist(self, pairs):
        for tag, val in pairs:
            self.flags(_LIST | _HAS_TAG)
            self.symbol(tag)
            self.value(val)
        self.flags(_NIL)

    def length(self, n):
        if n > 2 ** 31 - 1:
            self.i32(-1)
            self.i32(n >> 32)
            self.flags(n & 0xFFFFFFFF)
        else:
            self.i32(n)

    def value(self, v):
        v = _as_robject(v)
        if v is None:
            self.flags(_NIL)
            return
        attr = [(k, a) for k, a in v.attr.items() if a is not None]
        fl = (_HAS_ATTR if attr else 0) | (_IS_OBJ if "class" in v.attr else 0)
        if v.kind == "S4":
            self.flags(_S4 | fl | _IS_OBJ | ((1 << 4) << 12))  # gp bit 4 = S4_OBJECT_MASK
            if attr:
                self.pairlist(attr)
            return
        code = {k: c for c, k in _KIND.items()}[v.kind]
        self.flags(code | fl)
        self.length(len(v.value))
        if v.kind in ("logical", "integer"):
            self.out.write(np.ascontiguousarray(v.value, dtype=">i4").tobytes())
        elif v.kind == "double":
            self.out.write(np.ascontiguousarray(v.value, dtype=">f8").tobytes())
        elif v.kind == "character":
            for s in v.value:
                self.charsxp(s)
        else:
            for m in v.value:
                self.value(m)
        if attr:
            self.pairlist(attr)


def _as_robject(v):
    """Python value -> RObject (None stays NULL)."""
    if v is None or isinstance(v, RObject):
        return v
    if isinstance(v, str):
        return RObject("character", [v])
    if isinstance(v, (bool, np.bool_)):
        return RObject("logical", np.array([int(v)], np.int32))
    if isinstance(v, (int, np.integer)):
        return RObject("integer", np.array([int(v)], np.int32))
    if isinstance(v, (float, np.floating)):
        return RObject("double", np.array([float(v)], np.float64))
    if isinstance(v, dict):
        return named_list(v)
    if isinstance(v, (list, tuple)):
        if len(v) and all(isinstance(s, str) or s is None for s in v):
            return RObject("character", list(v))
        return RObject("list", [_as_robject(m) for m in v])
    a = np.asarray(v)
    if a.ndim == 2:
        return matrix(a)
    if a.ndim != 1:
        raise RdaError("cannot serialise an array of %d dimensions" % a.ndim)
    if a.dtype == np.bool_:
        return RObject("logical", a.astype(np.int32))
    if np.issubdtype(a.dtype, np.integer):
        if a.size and (a.min() <= _NA_INT or a.max() > 2 ** 31 - 1):
            return RObject("double", a.astype(np.float64))
        return RObject("integer", a.astype(np.int32))
    if np.issubdtype(a.dtype, np.floating):
        return RObject("double", a.astype(np.float64))
    if a.dtype.kind in "US":
        return RObject("character", [str(s) for s in a])
    raise RdaError("cannot serialise dtype %s" % a.dtype)


def matrix(a, rownames=None, colnames=None):
    """(nrow, ncol) numpy array -> R ``matrix`` (column-major data + dim [+ dimnames])."""
    a = np.asarray(a)
    if a.ndim != 2:
        raise RdaError("matrix() needs a 2-d array")
    flat = _as_robject(np.ravel(a, order="F"))
    flat.attr["dim"] = RObject("integer", np.array(a.shape, np.int32))
    if rownames is not None or colnames is not None:
        if rownames is not None and len(rownames) != a.shape[0]:
            raise RdaError("rownames: %d names for %d rows" % (len(rownames), a.shape[0]))
        if colnames is not None and len(colnames) != a.shape[1]:
            raise RdaError("colnames: %d names for %d columns" % (len(colnames), a.shape[1]))
        flat.attr["dimnames"] = RObject("list", [
            None if rownames is None else RObject("character", [str(s) for s in rownames]),
            None if colnames is None else RObject("character", [str(s) for s in colnames])])
    return flat


def named_list(d, rclass=None):
    """dict -> named R list (insertion order kept), optionally with a class attribute."""
    out = RObject("list", [_as_robject(v) for v in d.values()])
    out.attr["names"] = RObject("character", [str(k) for k in d])
    if rclass is not None:
        out.attr["class"] = RObject("character", [rclass] if isinstance(rclass, str) else list(rclass))
    return out


def _compress(buf, compress):
    if compress in (None, False, "none"):
        return buf
    if compress in (True, "gzip"):
        return gzip.compress(buf, compresslevel=6, mtime=0)
    if compress == "bzip2":
        return bz2.compress(buf, 9)
    if compress == "xz":
        return lzma.compress(buf)
    raise RdaError("unknown compression %r" % (compress,))


def write_rda(path, objects, compress="gzip"):
    """``save(list = names(objects), file = path)``; ``path=None`` returns the bytes."""
    w = _Writer()
    w.out.write(b"RDX2\nX\n")
    w.header()
    if not objects:
        raise RdaError("nothing to save")
    w.pairlist(list(objects.items()))
    buf = _compress(w.out.getvalue(), compress)
    if path is None:
        return buf
    with open(path, "wb") as fh:
        fh.write(buf)
    return None


def write_rds(path, value, compress="gzip"):
    """``saveRDS(value, file = path)``; ``path=None`` returns the bytes."""
    w = _Writer()
    w.out.write(b"X\n")
    w.header()
    w.value(value)
    buf = _compress(w.out.getvalue(), compress)
    if path is None:
        return buf
    with open(path, "wb") as fh:
        fh.write(buf)
    return None


# ----------------------------------------------------------------------------------------------
# R value -> Python
# ----------------------------------------------------------------------------------------------
def _dimnames(obj):
    dn = obj.attr.get("dimnames")
    if dn is None:
        return None, None
    rn, cn = (None if d is None else list(d.value) for d in dn.value)
    return rn, cn


def to_python(obj):
    """Matrix -> C-ordered (nrow, ncol) numpy array, vector -> 1-d array (character -> list of
    str), named list -> dict (recursively), unnamed list -> list."""
    if obj is None or not isinstance(obj, RObject):
        return obj
    if obj.kind == "S4":
        return obj
    if obj.kind == "list":
        vals = [to_python(v) for v in obj.value]
        nm = obj.names
        return dict(zip(nm, vals)) if nm is not None else vals
    if obj.kind == "character":
        return list(obj.value)
    v = obj.value
    if obj.kind == "logical":
        v = np.where(v == _NA_INT, -1, v).astype(np.int8)  # -1 marks NA
    dim = obj.attr.get("dim")
    if dim is not None and len(dim.value) == 2:
        nrow, ncol = (int(d) for d in dim.value)
        return np.ascontiguousarray(v.reshape(ncol, nrow).T)
    return np.array(v)


# ----------------------------------------------------------------------------------------------
# saver objects and count matrices
# ----------------------------------------------------------------------------------------------
_INFO_KEYS = ("size.factor", "maxcor", "lambda.max", "lambda.min", "sd.cv", "pred.time", "var.time",
              "cutoff", "lambda.coefs", "total.time")   # R/saver_fit.R:66-70


def _saver_from_r(obj):
    from . import host
    if not isinstance(obj, RObject) or obj.kind != "list" or obj.names is None:
        raise RdaError("not a saver object (expected a named list)")
    nm = obj.names
    if "estimate" not in nm:
        raise RdaError("not a saver object: no 'estimate' (fields: %s)" % nm)
    est = to_python(obj["estimate"])
    gene_names, cell_names = _dimnames(obj["estimate"])
    info = to_python(obj["info"]) if "info" in nm else {}
    if isinstance(info, list):
        info = dict(enumerate(info))
    for k, v in list(info.items()):  # scalars come back as numbers (cutoff, total.time)
        if isinstance(v, np.ndarray) and v.shape == (1,) and k in ("cutoff", "total.time"):
            info[k] = float(v[0])
    if "alpha" in nm and "beta" in nm:  # legacy layout (R/utils.R:137-232)
        return host.LegacySaverResult(est, to_python(obj["alpha"]), to_python(obj["beta"]), info,
                                      gene_names=gene_names, cell_names=cell_names)
    if "se" not in nm:
        raise RdaError("not a saver object: neither 'se' nor 'alpha'/'beta' (fields: %s)" % nm)
    return host.SaverResult(est, to_python(obj["se"]), info, gene_names=gene_names,
                            cell_names=cell_names)


def _pick(objs, name, what):
    if name is not None:
        if name not in objs:
            raise RdaError("no object %r in the file (has: %s)" % (name, sorted(objs)))
        return objs[name]
    if len(objs) != 1:
        raise RdaError("the file holds %d objects (%s): name the %s" % (len(objs), sorted(objs), what))
    return next(iter(objs.values()))


def load_saver(path, name=None):
    """A ``saver`` object from a ``save()`` (.rda/.RData) or ``saveRDS()`` (.rds) file ->
    ``host.SaverResult`` (or ``host.LegacySaverResult`` for alpha/beta objects)."""
    buf = _open(path)
    obj = read_rds(buf) if buf[:2] == b"X\n" else _pick(read_rda(buf), name, "saver object")
    return _saver_from_r(obj)


def _info_to_r(info):
    out = {}
    keys = [k for k in _INFO_KEYS if k in info] + [k for k in info if k not in _INFO_KEYS]
    for k in keys:
        v = info[k]
        if k == "total.time":   # a difftime in seconds (R/saver_fit.R:440-441)
            r = RObject("double", np.array([float(v)], np.float64))
            r.attr["class"] = RObject("character", ["difftime"])
            r.attr["units"] = RObject("character", ["secs"])
        elif k == "lambda.coefs" and v is not None and np.size(v) == 2:
            r = RObject("double", np.asarray(v, np.float64).ravel())
            r.attr["names"] = RObject("character", ["(Intercept)", "x"])  # coef(lm(y ~ x))
        elif v is None:
            r = None
        else:
            r = RObject("double", np.atleast_1d(np.asarray(v, np.float64)).ravel())
        out[str(k)] = r
    return named_list(out)


def save_saver(path, obj, name="saver.out", compress="gzip"):
    """Write a ``host.SaverResult`` / ``host.LegacySaverResult`` as an R ``saver`` object
    (class "saver"; an .rds path writes ``saveRDS`` format, anything else ``save`` format)."""
    gn, cn = obj.gene_names, obj.cell_names
    fields = {"estimate": matrix(np.asarray(obj.estimate, np.float64), gn, cn)}
    if hasattr(obj, "alpha"):
        fields["alpha"] = matrix(np.asarray(obj.alpha, np.float64), gn, cn)
        fields["beta"] = matrix(np.asarray(obj.beta, np.float64), gn, cn)
    else:
        se = obj.se
        fields["se"] = None if se is None else matrix(np.asarray(se, np.float64), gn, cn)
    r = RObject("list", list(fields.values()) + [_info_to_r(obj.info)])
    r.attr["names"] = RObject("character", list(fields) + ["info"])
    r.attr["class"] = RObject("character", ["saver"])
    if str(path).endswith(".rds"):
        return write_rds(path, r, compress)
    return write_rda(path, {name: r}, compress)


def load_counts(path, name=None):
    """An expression matrix from a .rda/.rds file: dense ``matrix`` -> (array, gene names, cell
    names); ``dgCMatrix`` -> (scipy.sparse.csc_matrix genes x cells, gene names, cell names)."""
    buf = _open(path)
    obj = read_rds(buf) if buf[:2] == b"X\n" else _pick(read_rda(buf), name, "matrix")
    if isinstance(obj, RObject) and obj.kind == "S4":
        cls = obj.rclass
        if "dgCMatrix" not in cls:
            raise RdaError("unsupported S4 class %s (only dgCMatrix)" % cls)
        import scipy.sparse as sp
        nrow, ncol = (int(d) for d in obj["Dim"].value)
        m = sp.csc_matrix((obj["x"].value, obj["i"].value, obj["p"].value), shape=(nrow, ncol))
        dn = obj.attr.get("Dimnames")
        rn, cn = (None, None) if dn is None else (None if d is None else list(d.value) for d in dn.value)
        return m, rn, cn
    if not isinstance(obj, RObject) or "dim" not in obj.attr:
        raise RdaError("not a matrix")
    rn, cn = _dimnames(obj)
    return to_python(obj), rn, cn


def save_counts(path, x, gene_names=None, cell_names=None, name="x", compress="gzip"):
    """Write a genes x cells matrix: numpy -> R ``matrix``; scipy sparse -> ``dgCMatrix``."""
    if hasattr(x, "tocsc"):
        c = x.tocsc()
        c.sort_indices()
        cls = RObject("character", ["dgCMatrix"])
        cls.attr["package"] = RObject("character", ["Matrix"])
        r = RObject("S4", None, {
            "i": RObject("integer", c.indices.astype(np.int32)),
            "p": RObject("integer", c.indptr.astype(np.int32)),
            "Dim": RObject("integer", np.array(c.shape, np.int32)),
            "Dimnames": RObject("list", [
                None if gene_names is None else RObject("character", [str(s) for s in gene_names]),
                None if cell_names is None else RObject("character", [str(s) for s in cell_names])]),
            "x": RObject("double", c.data.astype(np.float64)),
            "factors": RObject("list", []),
            "class": cls,
        })
    else:
        r = matrix(np.asarray(x), gene_names, cell_names)
    if str(path).endswith(".rds"):
        return write_rds(path, r, compress)
    return write_rda(path, {name: r}, compress)
