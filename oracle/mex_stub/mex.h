/* Minimal functional stand-in for MATLAB's mex.h.
 *
 * TEST INFRASTRUCTURE ONLY.  It exists so that the reference's own
 * matlab_code/sift/siftmatch.c (which #includes "mexutils.c" -> "mex.h")
 * can be compiled where it lies under /root/reference and driven through its
 * real mexFunction gateway from ctypes (see oracle/Makefile, oracle/refmex.py).
 * Nothing in the product path includes this header.
 *
 * Only the handful of mx / mex entry points that siftmatch.c and mexutils.c use
 * are provided (siftmatch.c:139-250, mexutils.c:15-98).
 */
#ifndef PRE3_ORACLE_MEX_STUB_H
#define PRE3_ORACLE_MEX_STUB_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  mxUNKNOWN_CLASS = 0,
  mxCELL_CLASS,
  mxSTRUCT_CLASS,
  mxLOGICAL_CLASS,
  mxCHAR_CLASS,
  mxVOID_CLASS,
  mxDOUBLE_CLASS,
  mxSINGLE_CLASS,
  mxINT8_CLASS,
  mxUINT8_CLASS,
  mxINT16_CLASS,
  mxUINT16_CLASS,
  mxINT32_CLASS,
  mxUINT32_CLASS,
  mxINT64_CLASS,
  mxUINT64_CLASS
} mxClassID;

typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

typedef struct mxArray_tag {
  mxClassID cls;
  size_t m, n;
  int ndim;
  int is_complex;
  int owns_data;
  void *data;
} mxArray;

typedef size_t mwSize;
typedef int bool_t_stub;

double mxGetInf(void);
double mxGetNaN(void);
int mxIsNumeric(const mxArray *a);
int mxIsDouble(const mxArray *a);
int mxIsComplex(const mxArray *a);
int mxIsChar(const mxArray *a);
int mxGetNumberOfDimensions(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
mxClassID mxGetClassID(const mxArray *a);
void *mxGetData(const mxArray *a);
double *mxGetPr(const mxArray *a);
double mxGetScalar(const mxArray *a);
void *mxMalloc(size_t n);
void *mxCalloc(size_t n, size_t sz);
void mxFree(void *p);
mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c);
void mxDestroyArray(mxArray *a);
/* additions for the product's own gateways (3pre_b200/mex_files/*.cpp): scalars and 1 x 1 structs */
typedef size_t mwIndex;
int mxIsEmpty(const mxArray *a);
int mxIsSparse(const mxArray *a); /* always 0: the stub has no sparse arrays */
mwIndex *mxGetIr(const mxArray *a);
mwIndex *mxGetJc(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateLogicalMatrix(size_t m, size_t n); /* uint8 0/1 data, class mxLOGICAL_CLASS */
int stub_struct_nfields(const mxArray *a);
const char *stub_struct_field_name(const mxArray *a, int k);
int mxIsStruct(const mxArray *a);
mxArray *mxCreateStructMatrix(size_t m, size_t n, int nfields, const char **names);
void mxSetField(mxArray *a, mwIndex i, const char *name, mxArray *v);
mxArray *mxGetField(const mxArray *a, mwIndex i, const char *name);
void mexErrMsgTxt(const char *msg);
void mexErrMsgIdAndTxt(const char *id, const char *msg, ...);
int mexPrintf(const char *fmt, ...);
int mexAtExit(void (*fn)(void));

/* the gateway every MEX file exports */
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

/* --- helpers for the ctypes driver (not part of MATLAB's API) --- */
mxArray *stub_wrap(int cls, size_t m, size_t n, void *data);
int stub_call_mex(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
const char *stub_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
