"""Minimal reader for R's RDX2 / XDR serialisation (``save()`` .rda files).

TEST INFRASTRUCTURE ONLY -- nothing under ``saver_b200/`` imports this.

It decodes the two bundled data objects of the reference
(``/root/reference/data/linnarsson.rda`` and ``linnarsson_saver.rda``,
documented in ``/root/reference/R/data.R:10,21``) so that the golden vectors
can be committed as ``tests/golden/linnarsson.npz`` by
``tests/golden/make_golden.py``.  Only the SEXP types that occur in those two
files are handled.

Format facts: bzip2 stream -> ``RDX2\\nX\\n`` + 3 int32 version words -> one
pairlist (tag -> value); everything is XDR big-endian; item header is an int32
of flags whose low byte is the SEXP type, 0x100 = is-object, 0x200 = has
attributes, 0x400 = has tag.
"""
import bz2
import struct

import numpy as np

NILVALUE, REFSXP = 254, 255
SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 1, 2, 9, 10, 13, 14, 16, 19


class _Stream:
    def __init__(self, buf):
        self.buf = buf
        self.pos = 0
        self.refs = []

    def int32(self):
        v = struct.unpack_from(">i", self.buf, self.pos)[0]
        self.pos += 4
        return v

    def item(self):
        flags = self.int32()
        return self._item(flags)

    def _pairlist(self, flags):
        out = []
        while True:
            has_attr = bool(flags & 0x200)
            has_tag = bool(flags & 0x400)
            if has_attr:
                self.item()
            tag = self.item() if has_tag else None
            out.append((tag, self.item()))
            flags = self.int32()
            if (flags & 0xFF) == NILVALUE:
                return out
            if (flags & 0xFF) != LISTSXP:
                raise ValueError("pairlist cdr is not a pairlist: %d" % (flags & 0xFF))

    def _item(self, flags):
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        if t == NILVALUE:
            return None
        if t == REFSXP:
            return self.refs[(flags >> 8) - 1]
        if t == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if t == LISTSXP:
            return self._pairlist(flags)
        if t == CHARSXP:
            n = self.int32()
            if n == -1:
                return None
            s = self.buf[self.pos:self.pos + n].decode("latin1")
            self.pos += n
            return s
        n = self.int32()
        if t in (LGLSXP, INTSXP):
            v = np.frombuffer(self.buf, ">i4", n, self.pos).astype(np.int32)
            self.pos += 4 * n
        elif t == REALSXP:
            v = np.frombuffer(self.buf, ">f8", n, self.pos).astype(np.float64)
            self.pos += 8 * n
        elif t in (STRSXP, VECSXP):
            v = [self.item() for _ in range(n)]
        else:
            raise ValueError("unsupported SEXP type %d" % t)
        attrs = dict(self.item()) if has_attr else {}
        return {"type": t, "value": v, "attr": attrs}


def load_rda(path):
    """Return {object name: decoded value} for an RDX2 ``.rda`` file."""
    with open(path, "rb") as fh:
        raw = fh.read()
    buf = bz2.decompress(raw)
    if buf[:7] != b"RDX2\nX\n":
        raise ValueError("not an RDX2 XDR stream")
    st = _Stream(buf)
    st.pos = 7
    st.int32(), st.int32(), st.int32()  # format, writer and min-reader versions
    out = dict(st.item())
    if st.pos != len(buf):
        raise ValueError("trailing bytes: consumed %d of %d" % (st.pos, len(buf)))
    return out


def as_matrix(obj):
    """Column-major R matrix -> C-ordered numpy (nrow, ncol) + (rownames, colnames)."""
    nrow, ncol = (int(v) for v in obj["attr"]["dim"]["value"])
    mat = np.ascontiguousarray(obj["value"].reshape(ncol, nrow).T)
    names = (None, None)
    dn = obj["attr"].get("dimnames")
    if dn is not None:
        names = tuple(None if d is None else d["value"] for d in dn["value"])
    return mat, names


def as_list(obj):
    """Named R list -> dict."""
    return dict(zip(obj["attr"]["names"]["value"], obj["value"]))
