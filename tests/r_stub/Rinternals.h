/* Minimal stand-in for R's C API -- TEST INFRASTRUCTURE ONLY.
 * R is absent from the development image, so integration/src/rcall_shim.c cannot be built with
 * R CMD SHLIB here.  tests/test_r_shim.py compiles the shim against these declarations (the
 * subset of Rinternals.h / R_ext/Rdynload.h the shim uses, with R's documented signatures) so that
 * every call into include/saver_cuda.h is type-checked by gcc, and links it with a tiny fake
 * runtime (r_stub_runtime.c) to drive one .Call end to end on the CPU-visible entry points. */
#ifndef SAVER_TEST_RINTERNALS_H
#define SAVER_TEST_RINTERNALS_H
#include <stddef.h>

typedef struct SEXPREC *SEXP;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef ptrdiff_t R_xlen_t;
typedef void *(*DL_FUNC)(void);
typedef void (*R_CFinalizer_t)(SEXP);
typedef struct _DllInfo DllInfo;
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct { const char *name; DL_FUNC fun; int numArgs; void *types; } R_CMethodDef;
typedef R_CMethodDef R_FortranMethodDef;
typedef R_CallMethodDef R_ExternalMethodDef;

#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define NA_LOGICAL (-2147483647 - 1)
#define NA_INTEGER (-2147483647 - 1)

extern SEXP R_NilValue;
double *REAL(SEXP x);
int *INTEGER(SEXP x);
int *LOGICAL(SEXP x);
SEXP PROTECT(SEXP x);
void UNPROTECT(int n);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
int Rf_length(SEXP x);
SEXP Rf_allocVector(unsigned int type, R_xlen_t n);
SEXP Rf_allocMatrix(unsigned int type, int nrow, int ncol);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
Rboolean Rf_isNull(SEXP x);
Rboolean Rf_isReal(SEXP x);
Rboolean Rf_isInteger(SEXP x);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_asLogical(SEXP x);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarLogical(int v);
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void *R_ExternalPtrAddr(SEXP p);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void R_ClearExternalPtr(SEXP p);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);
int R_registerRoutines(DllInfo *info, const R_CMethodDef *c, const R_CallMethodDef *call,
                       const R_FortranMethodDef *f, const R_ExternalMethodDef *ext);
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value);
#endif
