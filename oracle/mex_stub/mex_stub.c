/* Tiny mx/mex runtime behind oracle/mex_stub/mex.h.  TEST INFRASTRUCTURE ONLY.
 * mexErrMsgTxt does not return in MATLAB (it longjmps into the interpreter,
 * siftmatch.c:155-190 relies on that); stub_call_mex() reproduces this with
 * setjmp/longjmp and reports the message through stub_last_error(). */
#include "mex.h"

#include <math.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static jmp_buf g_jmp;
static int g_jmp_armed = 0;
static char g_err[512];

double mxGetInf(void) { return INFINITY; }
double mxGetNaN(void) { return NAN; }

int mxIsNumeric(const mxArray *a) {
  switch (a->cls) {
    case mxDOUBLE_CLASS: case mxSINGLE_CLASS: case mxINT8_CLASS: case mxUINT8_CLASS:
    case mxINT16_CLASS: case mxUINT16_CLASS: case mxINT32_CLASS: case mxUINT32_CLASS:
    case mxINT64_CLASS: case mxUINT64_CLASS:
      return 1;
    default:
      return 0;
  }
}
int mxIsDouble(const mxArray *a) { return a->cls == mxDOUBLE_CLASS; }
int mxIsComplex(const mxArray *a) { return a->is_complex; }
int mxIsChar(const mxArray *a) { return a->cls == mxCHAR_CLASS; }
int mxGetNumberOfDimensions(const mxArray *a) { return a->ndim; }
size_t mxGetM(const mxArray *a) { return a->m; }
size_t mxGetN(const mxArray *a) { return a->n; }
mxClassID mxGetClassID(const mxArray *a) { return a->cls; }
void *mxGetData(const mxArray *a) { return a->data; }
double *mxGetPr(const mxArray *a) { return (double *)a->data; }
double mxGetScalar(const mxArray *a) { return ((double *)a->data)[0]; }
void *mxMalloc(size_t n) { return malloc(n ? n : 1); }
void *mxCalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz ? sz : 1); }
void mxFree(void *p) { free(p); }

mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c) {
  mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
  a->cls = mxDOUBLE_CLASS;
  a->m = m;
  a->n = n;
  a->ndim = 2;
  a->is_complex = (c == mxCOMPLEX);
  a->owns_data = 1;
  a->data = calloc(m * n ? m * n : 1, sizeof(double));
  return a;
}

/* 1 x 1 structs: data -> stub_struct {nfields, names[], vals[]} */
typedef struct {
  int nfields;
  const char **names;
  mxArray **vals;
} stub_struct;

void mxDestroyArray(mxArray *a) {
  if (!a) return;
  if (a->cls == mxSTRUCT_CLASS && a->data) {
    stub_struct *s = (stub_struct *)a->data;
    for (int i = 0; i < s->nfields; ++i) mxDestroyArray(s->vals[i]);
    free(s->names);
    free(s->vals);
    free(s);
    free(a);
    return;
  }
  if (a->owns_data) free(a->data);
  free(a);
}

int mxIsEmpty(const mxArray *a) { return a->m * a->n == 0; }
int mxIsSparse(const mxArray *a) { (void)a; return 0; }
mwIndex *mxGetIr(const mxArray *a) { (void)a; return 0; }
mwIndex *mxGetJc(const mxArray *a) { (void)a; return 0; }
size_t mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
mxArray *mxCreateDoubleScalar(double v) {
  mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL);
  ((double *)a->data)[0] = v;
  return a;
}
mxArray *mxCreateLogicalMatrix(size_t m, size_t n) {
  mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
  a->cls = mxLOGICAL_CLASS;
  a->m = m;
  a->n = n;
  a->ndim = 2;
  a->owns_data = 1;
  a->data = calloc(m * n ? m * n : 1, 1);
  return a;
}
int mxIsStruct(const mxArray *a) { return a->cls == mxSTRUCT_CLASS; }
int stub_struct_nfields(const mxArray *a) { return a->cls == mxSTRUCT_CLASS ? ((stub_struct *)a->data)->nfields : 0; }
const char *stub_struct_field_name(const mxArray *a, int k) { return ((stub_struct *)a->data)->names[k]; }
mxArray *mxCreateStructMatrix(size_t m, size_t n, int nfields, const char **names) {
  mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
  stub_struct *s = (stub_struct *)calloc(1, sizeof(stub_struct));
  s->nfields = nfields;
  s->names = (const char **)calloc((size_t)(nfields ? nfields : 1), sizeof(char *));
  s->vals = (mxArray **)calloc((size_t)(nfields ? nfields : 1), sizeof(mxArray *));
  for (int i = 0; i < nfields; ++i) s->names[i] = names[i];
  a->cls = mxSTRUCT_CLASS;
  a->m = m;
  a->n = n;
  a->ndim = 2;
  a->owns_data = 0;
  a->data = s;
  return a;
}
void mxSetField(mxArray *a, mwIndex i, const char *name, mxArray *v) {
  (void)i;
  stub_struct *s = (stub_struct *)a->data;
  for (int k = 0; k < s->nfields; ++k)
    if (strcmp(s->names[k], name) == 0) {
      mxDestroyArray(s->vals[k]);
      s->vals[k] = v;
    }
}
mxArray *mxGetField(const mxArray *a, mwIndex i, const char *name) {
  (void)i;
  if (a->cls != mxSTRUCT_CLASS) return 0;
  stub_struct *s = (stub_struct *)a->data;
  for (int k = 0; k < s->nfields; ++k)
    if (strcmp(s->names[k], name) == 0) return s->vals[k];
  return 0;
}

void mexErrMsgTxt(const char *msg) {
  snprintf(g_err, sizeof g_err, "%s", msg ? msg : "");
  if (g_jmp_armed) longjmp(g_jmp, 1);
  fprintf(stderr, "mexErrMsgTxt outside stub_call_mex: %s\n", g_err);
  abort();
}

void mexErrMsgIdAndTxt(const char *id, const char *msg, ...) {
  (void)id;
  mexErrMsgTxt(msg);
}

int mexPrintf(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  int r = vfprintf(stdout, fmt, ap);
  va_end(ap);
  return r;
}

static void (*g_atexit)(void) = 0;
int mexAtExit(void (*fn)(void)) {
  g_atexit = fn;
  return 0;
}
void stub_run_atexit(void) {
  if (g_atexit) g_atexit();
  g_atexit = 0;
}

mxArray *stub_wrap(int cls, size_t m, size_t n, void *data) {
  mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
  a->cls = (mxClassID)cls;
  a->m = m;
  a->n = n;
  a->ndim = 2;
  a->is_complex = 0;
  a->owns_data = 0;
  a->data = data;
  return a;
}

int stub_call_mex(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  g_err[0] = 0;
  g_jmp_armed = 1;
  if (setjmp(g_jmp)) {
    g_jmp_armed = 0;
    return 1; /* error raised by the gateway */
  }
  mexFunction(nlhs, plhs, nrhs, prhs);
  g_jmp_armed = 0;
  return 0;
}

const char *stub_last_error(void) { return g_err; }
