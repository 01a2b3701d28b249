/* [s, R, T, err] = absoluteOrientationQuaternion_mex(A, B, doScale)
 * GPU drop-in for M/absoluteOrientationQuaternion.m:28-127 (A, B: 3 x N; B ~ s*R*A + T). */
#include "pre3_mex_common.h"

extern "C" void mexFunction(int nout, mxArray *out[], int nin, const mxArray *in[]) {
  if (nin < 2 || nin > 3) mexErrMsgTxt("absoluteOrientationQuaternion: A, B [, doScale]");
  if (nout > 4) mexErrMsgTxt("Too many output arguments");
  int doScale = 1; /* nargin < 3 -> doScale = 1 (absoluteOrientationQuaternion.m:32-34) */
  if (nin == 3) doScale = mxGetScalar(in[2]) != 0.0;
  if (mxGetM(in[0]) != mxGetM(in[1]) || mxGetN(in[0]) != mxGetN(in[1]))
    mexErrMsgTxt("Point sets need to have same size."); /* :41-43 */
  if (mxGetM(in[0]) != 3) mexErrMsgTxt("Need points of dimension 3"); /* :46-48 */
  if (mxGetN(in[0]) < 4) mexErrMsgTxt("Need at least 4 point pairs"); /* :51-54 */
  if (mxGetClassID(in[0]) != mxDOUBLE_CLASS || mxGetClassID(in[1]) != mxDOUBLE_CLASS)
    mexErrMsgTxt("A and B must be double");
  double s = 1.0, R[9], T[3], err = 0.0;
  pre3_mex_check(pre3_horn(pre3_mex_ctx(), mxGetPr(in[0]), mxGetPr(in[1]), (int)mxGetN(in[0]), doScale, &s, R, T, &err));
  out[0] = mxCreateDoubleScalar(s);
  if (nout > 1) {
    out[1] = mxCreateDoubleMatrix(3, 3, mxREAL);
    memcpy(mxGetPr(out[1]), R, sizeof R);
  }
  if (nout > 2) {
    out[2] = mxCreateDoubleMatrix(3, 1, mxREAL);
    memcpy(mxGetPr(out[2]), T, sizeof T);
  }
  if (nout > 3) out[3] = mxCreateDoubleScalar(err);
}
