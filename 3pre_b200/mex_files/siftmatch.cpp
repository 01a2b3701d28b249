/* siftmatch MEX gateway over libpre3.so -- drop-in for M/sift/siftmatch.c:139-250.
 *
 *   matches = siftmatch(L1, L2)            [matches, D] = siftmatch(L1, L2, thresh)
 *
 * Same argument checks and messages as the reference gateway; the compare_mx*_CLASS loops
 * (siftmatch.c:83-132) run on the GPU (pre3_siftmatch).  Placed earlier on the MATLAB path
 * than M/sift, the built siftmatch.mexa64 shadows the reference's MEX file. */
#include "pre3_mex_common.h"

extern "C" void mexFunction(int nout, mxArray *out[], int nin, const mxArray *in[]) {
  enum { L1 = 0, L2, THRESH };
  enum { MATCHES = 0, D };
  double thresh = 1.5; /* siftmatch.c:146 */
  if (nin < 2) mexErrMsgTxt("At least two input arguments required");
  if (nout > 2) mexErrMsgTxt("Too many output arguments");
  if (!mxIsNumeric(in[L1]) || !mxIsNumeric(in[L2]) || mxGetNumberOfDimensions(in[L1]) > 2 ||
      mxGetNumberOfDimensions(in[L2]) > 2)
    mexErrMsgTxt("L1 and L2 must be two dimensional numeric arrays");
  const int K1 = (int)mxGetN(in[L1]), K2 = (int)mxGetN(in[L2]), ND = (int)mxGetM(in[L1]);
  if ((int)mxGetM(in[L2]) != ND) mexErrMsgTxt("L1 and L2 must have the same number of rows");
  const mxClassID cls = mxGetClassID(in[L1]);
  if (mxGetClassID(in[L2]) != cls) mexErrMsgTxt("L1 and L2 must be of the same class");
  if (nin == 3) {
    if (!mxIsNumeric(in[THRESH]) || mxIsComplex(in[THRESH]) || mxGetClassID(in[THRESH]) != mxDOUBLE_CLASS ||
        mxGetM(in[THRESH]) != 1 || mxGetN(in[THRESH]) != 1)
      mexErrMsgTxt("THRESH should be a real scalar");
    thresh = *mxGetPr(in[THRESH]);
  } else if (nin > 3) {
    mexErrMsgTxt("At most three arguments are allowed");
  }
  int pcls;
  switch (cls) {
    case mxDOUBLE_CLASS: pcls = PRE3_CLASS_DOUBLE; break;
    case mxSINGLE_CLASS: pcls = PRE3_CLASS_SINGLE; break;
    case mxINT8_CLASS: pcls = PRE3_CLASS_INT8; break;
    case mxUINT8_CLASS: pcls = PRE3_CLASS_UINT8; break;
    default: mexErrMsgTxt("Unsupported numeric class"); return;
  }
  pre3_ctx *ctx = pre3_mex_ctx();
  int32_t *pairs = (int32_t *)mxMalloc(sizeof(int32_t) * 2 * (size_t)(K1 > 0 ? K1 : 1));
  /* scores only when asked for ([matches, D] = ...): without them the certified accept branch of the matcher skips
   * the exact distance and nothing is copied back */
  double *score = nout > 1 ? (double *)mxMalloc(sizeof(double) * (size_t)(K1 > 0 ? K1 : 1)) : NULL;
  int32_t n = 0;
  int rc = pre3_siftmatch(ctx, mxGetData(in[L1]), mxGetData(in[L2]), pcls, K1, K2, ND, thresh, pairs, score, &n);
  if (rc != PRE3_OK) {
    mxFree(pairs);
    if (score) mxFree(score);
    pre3_mex_check(rc);
  }
  out[MATCHES] = mxCreateDoubleMatrix(2, (size_t)n, mxREAL);
  double *M = mxGetPr(out[MATCHES]);
  double *Dp = NULL;
  if (nout > 1) {
    out[D] = mxCreateDoubleMatrix(1, (size_t)n, mxREAL);
    Dp = mxGetPr(out[D]);
  }
  for (int i = 0; i < n; ++i) {
    M[2 * i] = pairs[2 * i] + 1; /* 1-based, siftmatch.c:241-242 */
    M[2 * i + 1] = pairs[2 * i + 1] + 1;
    if (Dp) Dp[i] = score[i];
  }
  mxFree(pairs);
  if (score) mxFree(score);
}
