/* mex_shim/mex.h -- declarations-only stand-in for MATLAB's / Octave's mex.h.
 *
 * The build container has no MATLAB, Octave or mex.h (SURVEY.md 8b "build reality").  The
 * gateways in this directory are written against the documented MEX C API; this header
 * declares exactly the subset they use so that `__graft_entry__.build()` can compile them
 * (object files only).  On a user's machine they are built with the real header:
 *     mex -O siftmatch.cpp -I<repo>/include -L<repo>/3pre_b200/lib -lpre3
 *     mkoctfile --mex siftmatch.cpp -I<repo>/include -L<repo>/3pre_b200/lib -lpre3
 */
#ifndef PRE3_MEX_SHIM_H
#define PRE3_MEX_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum {
  mxUNKNOWN_CLASS = 0, mxCELL_CLASS, mxSTRUCT_CLASS, mxLOGICAL_CLASS, mxCHAR_CLASS, mxVOID_CLASS,
  mxDOUBLE_CLASS, mxSINGLE_CLASS, mxINT8_CLASS, mxUINT8_CLASS, mxINT16_CLASS, mxUINT16_CLASS,
  mxINT32_CLASS, mxUINT32_CLASS, mxINT64_CLASS, mxUINT64_CLASS
} mxClassID;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
int mxIsNumeric(const mxArray *a);
int mxIsComplex(const mxArray *a);
int mxIsStruct(const mxArray *a);
int mxIsEmpty(const mxArray *a);
int mxIsSparse(const mxArray *a);
mwIndex *mxGetIr(const mxArray *a);
mwIndex *mxGetJc(const mxArray *a);
mwSize mxGetNumberOfDimensions(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
mxClassID mxGetClassID(const mxArray *a);
void *mxGetData(const mxArray *a);
double *mxGetPr(const mxArray *a);
double mxGetScalar(const mxArray *a);
mxArray *mxGetField(const mxArray *a, mwIndex i, const char *name);
void mxSetField(mxArray *a, mwIndex i, const char *name, mxArray *v);
mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c);
mxArray *mxCreateDoubleScalar(double v);
mxArray *mxCreateLogicalMatrix(size_t m, size_t n);
mxArray *mxCreateStructMatrix(size_t m, size_t n, int nfields, const char **names);
void *mxMalloc(size_t n);
void mxFree(void *p);
void mexErrMsgTxt(const char *msg);
void mexErrMsgIdAndTxt(const char *id, const char *msg, ...);
int mexAtExit(void (*fn)(void));
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
