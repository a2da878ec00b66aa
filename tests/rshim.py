"""Drives integration/src/rcall_shim.c (the .Call glue a SAVER maintainer compiles with
R CMD SHLIB) outside R: the shim is compiled against a stand-in for R's headers
(tests/r_stub/Rinternals.h) and linked with a small fake runtime (tests/r_stub/r_stub_runtime.c)
and with libsaver_cuda.so.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "r_stub")
SHIM = os.path.join(ROOT, "integration", "src", "rcall_shim.c")
LIB = os.path.join(STUB, "librshim_test.so")
LGLSXP, INTSXP, REALSXP, VECSXP, NILSXP, EXTPTRSXP = 10, 13, 14, 19, 0, 22
NA_LOGICAL = -2 ** 31

STRICT = ["-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
          "-Werror=int-conversion", "-Wno-cast-function-type"]


def build(force=False):
    from saver_b200 import build as _b
    cuda_lib = _b.build_library()
    srcs = [SHIM, os.path.join(STUB, "r_stub_runtime.c")]
    deps = srcs + [os.path.join(STUB, "Rinternals.h"), os.path.join(STUB, "R.h"),
                   os.path.join(ROOT, "include", "saver_cuda.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        cmd = ["gcc", "-O1", "-std=gnu11", "-fPIC", "-shared", "-fvisibility=hidden"] + STRICT + \
              ["-I" + os.path.join(ROOT, "include"), "-I" + STUB] + srcs + \
              ["-o", LIB, "-L" + os.path.dirname(cuda_lib), "-lsaver_cuda",
               "-Wl,-rpath,$ORIGIN/../../saver_b200/csrc"]  # relative: the tree may be relocated
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("the .Call shim does not compile against include/saver_cuda.h:\n" + r.stderr)
    return LIB


class RError(RuntimeError):
    """An R condition raised by the shim (Rf_error)."""


class Session:
    """A fake R session: converts numpy values to SEXPs, runs .Call routines, converts back."""

    def __init__(self):
        path = build()
        from saver_b200 import build as _b
        C.CDLL(_b.LIB, mode=C.RTLD_GLOBAL)  # resolve libsaver_cuda wherever the tree lives
        self.lib = lib = C.CDLL(path)
        P = C.c_void_p
        lib.stub_last_error.restype = C.c_char_p
        lib.stub_routine_name.restype = C.c_char_p
        for f in ("stub_nil", "stub_vector", "stub_matrix", "stub_data", "stub_elt"):
            getattr(lib, f).restype = P
        lib.stub_vector.argtypes = [C.c_int, C.c_ssize_t, P]
        lib.stub_matrix.argtypes = [C.c_int, C.c_int, C.c_int, P]
        lib.stub_len.restype = C.c_longlong
        for f in ("stub_type", "stub_len", "stub_nrow", "stub_ncol", "stub_data"):
            getattr(lib, f).argtypes = [P]
        lib.stub_elt.argtypes = [P, C.c_int]
        lib.stub_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(P), C.POINTER(P)]
        assert lib.stub_init() == 1

    def routines(self):
        return {self.lib.stub_routine_name(k).decode(): self.lib.stub_routine_nargs(k)
                for k in range(self.lib.stub_routine_count())}

    # -- Python -> SEXP ---------------------------------------------------------------------
    def sexp(self, v):
        lib = self.lib
        if isinstance(v, _Sexp):
            return v.p
        if v is None:
            return lib.stub_nil()
        if isinstance(v, (bool, np.bool_)):
            a = np.array([int(v)], np.int32)
            return lib.stub_vector(LGLSXP, 1, a.ctypes.data)
        if isinstance(v, (int, np.integer)):
            a = np.array([int(v)], np.int32)
            return lib.stub_vector(INTSXP, 1, a.ctypes.data)
        if isinstance(v, (float, np.floating)):
            a = np.array([float(v)], np.float64)
            return lib.stub_vector(REALSXP, 1, a.ctypes.data)
        a = np.asarray(v)
        if np.issubdtype(a.dtype, np.integer):
            t, a = INTSXP, a.astype(np.int32)
        elif np.issubdtype(a.dtype, np.floating):
            t, a = REALSXP, a.astype(np.float64)
        else:
            raise TypeError("no SEXP for dtype %s" % a.dtype)
        if a.ndim == 1:
            a = np.ascontiguousarray(a)
            return lib.stub_vector(t, a.size, a.ctypes.data)
        if a.ndim == 2:  # an R matrix is column-major
            f = np.asfortranarray(a)
            return lib.stub_matrix(t, a.shape[0], a.shape[1], f.ctypes.data)
        raise TypeError("no SEXP for %d-d arrays" % a.ndim)

    # -- SEXP -> Python ---------------------------------------------------------------------
    def value(self, p):
        lib = self.lib
        t = lib.stub_type(p)
        if t == NILSXP:
            return None
        if t == EXTPTRSXP:
            return _Sexp(p)
        n = lib.stub_len(p)
        if t == VECSXP:
            return [self.value(lib.stub_elt(p, i)) for i in range(n)]
        dt = {REALSXP: np.float64, INTSXP: np.int32, LGLSXP: np.int32}[t]
        buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(lib.stub_data(p))
        a = np.frombuffer(buf, dt, n).copy()
        nr, nc = lib.stub_nrow(p), lib.stub_ncol(p)
        return a.reshape(nc, nr).T if nr >= 0 else a

    def call(self, name, *args):
        """.Call(name, ...)"""
        arr = (C.c_void_p * max(len(args), 1))(*[self.sexp(a) for a in args])
        out = C.c_void_p()
        rc = self.lib.stub_call(name.encode(), len(args), arr, C.byref(out))
        if rc != 0:
            raise RError(self.lib.stub_last_error().decode())
        return self.value(out.value)

    def close(self):
        self.lib.stub_reset()  # runs the external-pointer finalizers (saver_cuda_destroy)


class _Sexp:
    """An opaque R value held by the test (the library handle's external pointer)."""

    def __init__(self, p):
        self.p = p


def r_call_sites(path):
    """[(routine, number of arguments after the routine)] for every .Call( in an R source."""
    src = open(path).read()
    src = "\n".join(line.split("#", 1)[0] for line in src.splitlines())
    out, pos = [], 0
    while True:
        i = src.find(".Call(", pos)
        if i < 0:
            return out
        j, depth, parts, cur = i + 6, 1, [], ""
        while depth:
            ch = src[j]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            if ch == "," and depth == 1:
                parts.append(cur.strip())
                cur = ""
            else:
                cur += ch
            j += 1
        parts.append(cur.strip())
        out.append((parts[0], len(parts) - 1))
        pos = j
