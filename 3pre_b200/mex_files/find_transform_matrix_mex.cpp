/* [rot, trans, state] = find_transform_matrix_mex(pset1, pset2)
 * GPU drop-in for M/mex_files/RANSAC_CALCULATION/find_transform_matrix.m:2-42 (pset: 3 x n). */
#include "pre3_mex_common.h"

extern "C" void mexFunction(int nout, mxArray *out[], int nin, const mxArray *in[]) {
  if (nin != 2) mexErrMsgTxt("find_transform_matrix: two input arguments required");
  if (nout > 3) mexErrMsgTxt("Too many output arguments");
  for (int i = 0; i < 2; ++i)
    if (mxGetClassID(in[i]) != mxDOUBLE_CLASS || mxIsComplex(in[i]) || mxGetM(in[i]) != 3)
      mexErrMsgTxt("pset1 and pset2 must be real double 3 x n matrices");
  if (mxGetN(in[0]) != mxGetN(in[1])) mexErrMsgTxt("pset1 and pset2 must have the same size");
  double rot[9], trans[3];
  int32_t state = 0;
  pre3_mex_check(pre3_find_transform_matrix(pre3_mex_ctx(), mxGetPr(in[0]), mxGetPr(in[1]), (int)mxGetN(in[0]), rot,
                                            trans, &state));
  out[0] = mxCreateDoubleMatrix(3, 3, mxREAL);
  memcpy(mxGetPr(out[0]), rot, sizeof rot); /* both column-major */
  if (nout > 1) {
    out[1] = mxCreateDoubleMatrix(3, 1, mxREAL);
    memcpy(mxGetPr(out[1]), trans, sizeof trans);
  }
  if (nout > 2) out[2] = mxCreateDoubleScalar((double)state);
}
