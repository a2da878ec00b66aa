/* A tiny fake of the R runtime -- TEST INFRASTRUCTURE ONLY (see Rinternals.h in this directory).
 * Just enough of the C API for integration/src/rcall_shim.c to run outside R: vectors and
 * matrices of doubles / ints / logicals, generic lists, external pointers with finalizers, and
 * Rf_error as a longjmp back to the trampoline that made the call (R does the same to its
 * top-level context).  Values are never freed individually; stub_reset() drops them all. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"

#define EXTPTRSXP 22
#define NILSXP 0

struct SEXPREC {
    int type, nrow, ncol;   /* nrow = ncol = -1 for plain vectors */
    R_xlen_t len;
    void *data;             /* doubles / ints / SEXP elements / the external address */
    R_CFinalizer_t fin;
    struct SEXPREC *next;   /* allocation list */
};

static struct SEXPREC nil_value = {NILSXP, -1, -1, 0, NULL, NULL, NULL};
SEXP R_NilValue = &nil_value;
static SEXP all_values = NULL;
static jmp_buf jump;
static int jump_armed = 0;
static char last_error[1024];
static const R_CallMethodDef *registered = NULL;

#define API __attribute__((visibility("default")))

static SEXP new_value(int type, R_xlen_t len, size_t elt)
{
    SEXP s = (SEXP)calloc(1, sizeof(*s));
    s->type = type; s->len = len; s->nrow = s->ncol = -1;
    s->data = calloc(len > 0 ? (size_t)len : 1, elt);
    s->next = all_values; all_values = s;
    return s;
}

static size_t elt_size(unsigned int type)
{
    switch (type) {
    case REALSXP: return sizeof(double);
    case INTSXP: case LGLSXP: return sizeof(int);
    case VECSXP: return sizeof(SEXP);
    default: Rf_error("stub runtime: unsupported vector type %u", type);
    }
}

API SEXP Rf_allocVector(unsigned int type, R_xlen_t n)
{
    SEXP s = new_value((int)type, n, elt_size(type));
    if (type == VECSXP) for (R_xlen_t i = 0; i < n; ++i) ((SEXP *)s->data)[i] = R_NilValue;
    return s;
}
API SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol)
{
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->nrow = nrow; s->ncol = ncol;
    return s;
}
static void want(SEXP x, int type, const char *acc)
{
    if (x->type != type) Rf_error("stub runtime: %s() on a value of type %d", acc, x->type);
}
API double *REAL(SEXP x) { want(x, REALSXP, "REAL"); return (double *)x->data; }
API int *INTEGER(SEXP x) { want(x, INTSXP, "INTEGER"); return (int *)x->data; }
API int *LOGICAL(SEXP x) { want(x, LGLSXP, "LOGICAL"); return (int *)x->data; }
API SEXP PROTECT(SEXP x) { return x; }
API void UNPROTECT(int n) { (void)n; }
API SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v)
{
    want(x, VECSXP, "SET_VECTOR_ELT");
    if (i < 0 || i >= x->len) Rf_error("stub runtime: list index out of range");
    ((SEXP *)x->data)[i] = v;
    return v;
}
API SEXP VECTOR_ELT(SEXP x, R_xlen_t i)
{
    want(x, VECSXP, "VECTOR_ELT");
    if (i < 0 || i >= x->len) Rf_error("stub runtime: list index out of range");
    return ((SEXP *)x->data)[i];
}
API int Rf_length(SEXP x) { return (int)x->len; }
API int Rf_nrows(SEXP x) { if (x->nrow < 0) Rf_error("object is not a matrix"); return x->nrow; }
API int Rf_ncols(SEXP x) { if (x->ncol < 0) Rf_error("object is not a matrix"); return x->ncol; }
API Rboolean Rf_isNull(SEXP x) { return x->type == NILSXP ? TRUE : FALSE; }
API Rboolean Rf_isReal(SEXP x) { return x->type == REALSXP ? TRUE : FALSE; }
API Rboolean Rf_isInteger(SEXP x) { return x->type == INTSXP ? TRUE : FALSE; }
API double Rf_asReal(SEXP x)
{
    if (x->len < 1) Rf_error("stub runtime: asReal of an empty value");
    if (x->type == REALSXP) return ((double *)x->data)[0];
    if (x->type == INTSXP || x->type == LGLSXP) return (double)((int *)x->data)[0];
    Rf_error("stub runtime: asReal of type %d", x->type);
}
API int Rf_asInteger(SEXP x) { return (int)Rf_asReal(x); }
API int Rf_asLogical(SEXP x)
{
    if (x->len < 1) Rf_error("stub runtime: asLogical of an empty value");
    if (x->type == LGLSXP || x->type == INTSXP) return ((int *)x->data)[0];
    return Rf_asReal(x) != 0.0;
}
API SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
API SEXP Rf_ScalarLogical(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); LOGICAL(s)[0] = v; return s; }
API void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error, sizeof last_error, fmt, ap);
    va_end(ap);
    if (!jump_armed) { fprintf(stderr, "Rf_error outside stub_call: %s\n", last_error); abort(); }
    longjmp(jump, 1);
}
API void *R_ExternalPtrAddr(SEXP p) { want(p, EXTPTRSXP, "R_ExternalPtrAddr"); return *(void **)p->data; }
API SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
    (void)tag; (void)prot;
    SEXP s = new_value(EXTPTRSXP, 1, sizeof(void *));
    *(void **)s->data = p;
    return s;
}
API void R_ClearExternalPtr(SEXP p) { want(p, EXTPTRSXP, "R_ClearExternalPtr"); *(void **)p->data = NULL; }
API void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; s->fin = fun; }
API int R_registerRoutines(DllInfo *info, const R_CMethodDef *c, const R_CallMethodDef *call,
                           const R_FortranMethodDef *f, const R_ExternalMethodDef *ext)
{
    (void)info; (void)c; (void)f; (void)ext;
    registered = call;
    return 1;
}
API Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value) { (void)info; (void)value; return TRUE; }

/* ---- what the Python side of the test drives ---- */
extern void R_init_SAVERcuda(DllInfo *dll);

API int stub_init(void) { R_init_SAVERcuda(NULL); return registered != NULL; }
API int stub_routine_count(void)
{
    int k = 0;
    while (registered && registered[k].name) ++k;
    return k;
}
API const char *stub_routine_name(int k) { return registered[k].name; }
API int stub_routine_nargs(int k) { return registered[k].numArgs; }
API const char *stub_last_error(void) { return last_error; }
API SEXP stub_nil(void) { return R_NilValue; }
API SEXP stub_vector(int type, R_xlen_t n, const void *src)
{
    SEXP s = Rf_allocVector((unsigned)type, n);
    if (src && n > 0) memcpy(s->data, src, (size_t)n * elt_size((unsigned)type));
    return s;
}
API SEXP stub_matrix(int type, int nrow, int ncol, const void *src)
{
    SEXP s = Rf_allocMatrix((unsigned)type, nrow, ncol);
    if (src && s->len > 0) memcpy(s->data, src, (size_t)s->len * elt_size((unsigned)type));
    return s;
}
API int stub_type(SEXP x) { return x->type; }
API long long stub_len(SEXP x) { return (long long)x->len; }
API int stub_nrow(SEXP x) { return x->nrow; }
API int stub_ncol(SEXP x) { return x->ncol; }
API void *stub_data(SEXP x) { return x->data; }
API SEXP stub_elt(SEXP x, int i) { return ((SEXP *)x->data)[i]; }

/* .Call(name, ...): look the routine up as R does, check the arity, run it under a jump target */
API int stub_call(const char *name, int nargs, SEXP *a, SEXP *out)
{
    typedef SEXP (*F)();
    const R_CallMethodDef *m = registered;
    for (; m && m->name; ++m) if (strcmp(m->name, name) == 0) break;
    if (!m || !m->name) { snprintf(last_error, sizeof last_error, "no such .Call routine %s", name); return 2; }
    if (m->numArgs != nargs) {
        snprintf(last_error, sizeof last_error, "%s takes %d arguments, got %d", name, m->numArgs, nargs);
        return 3;
    }
    F f = (F)m->fun;
    jump_armed = 1;
    if (setjmp(jump)) { jump_armed = 0; return 1; }
    SEXP r;
    switch (nargs) {
    case 1: r = f(a[0]); break;
    case 2: r = f(a[0], a[1]); break;
    case 3: r = f(a[0], a[1], a[2]); break;
    case 4: r = f(a[0], a[1], a[2], a[3]); break;
    case 5: r = f(a[0], a[1], a[2], a[3], a[4]); break;
    case 6: r = f(a[0], a[1], a[2], a[3], a[4], a[5]); break;
    case 13: r = f(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]); break;
    default: jump_armed = 0; snprintf(last_error, sizeof last_error, "stub_call: %d arguments", nargs); return 4;
    }
    jump_armed = 0;
    *out = r;
    return 0;
}

/* garbage collection at session end: run the finalizers, free everything */
API void stub_reset(void)
{
    for (SEXP s = all_values; s; s = s->next)
        if (s->type == EXTPTRSXP && s->fin && *(void **)s->data) s->fin(s);
    while (all_values) {
        SEXP n = all_values->next;
        free(all_values->data);
        free(all_values);
        all_values = n;
    }
}
