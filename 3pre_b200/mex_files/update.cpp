/* [x_k_k, p_k_k, K] = update(x_km1_k, p_km1_k, H, R, z, h)
 *
 * Same name and signature as M/update.m:27 of 3PRE (the EKF update every partial update goes through:
 * M/@ekf_filter/ekf_update_li_inliers.m:28, ekf_update_hi_inliers.m:31): built as update.mex* and placed earlier on
 * the MATLAB path it shadows the .m file -- no shim needed.  The arithmetic (:32-48: S, K = P H' inv(S), x + K (z - h),
 * P - K S K', symmetrisation, quaternion normalisation Jacobian) runs on the GPU (pre3_ekf_update_dense).
 * H may be sparse (the callers stack sparse 2 x n blocks) or full; R full or sparse (the callers pass eye(m)). */
#include "pre3_mex_common.h"

static double *densify(const mxArray *a, size_t rows, size_t cols) {
  double *d = (double *)mxMalloc(sizeof(double) * (rows * cols > 0 ? rows * cols : 1));
  const double *pr = mxGetPr(a);
  if (mxIsSparse(a)) {
    memset(d, 0, sizeof(double) * rows * cols);
    const mwIndex *ir = mxGetIr(a), *jc = mxGetJc(a);
    for (size_t c = 0; c < cols; ++c)
      for (mwIndex k = jc[c]; k < jc[c + 1]; ++k) d[c * rows + ir[k]] = pr[k];
  } else {
    memcpy(d, pr, sizeof(double) * rows * cols);
  }
  return d;
}

extern "C" void mexFunction(int nout, mxArray *out[], int nin, const mxArray *in[]) {
  if (nin != 6) mexErrMsgTxt("update: six inputs required (x_km1_k, p_km1_k, H, R, z, h)");
  if (nout > 3) mexErrMsgTxt("Too many output arguments");
  for (int i = 0; i < 6; ++i)
    if (mxGetClassID(in[i]) != mxDOUBLE_CLASS || mxIsComplex(in[i])) mexErrMsgTxt("update: inputs must be real double");
  const size_t n = mxGetNumberOfElements(in[0]);
  const size_t m = mxGetM(in[4]) * mxGetN(in[4]);
  if (mxGetM(in[1]) != n || mxGetN(in[1]) != n) mexErrMsgTxt("update: p_km1_k must be n x n");
  out[0] = mxCreateDoubleMatrix(mxGetM(in[0]), mxGetN(in[0]), mxREAL);
  if (m == 0) { /* :50-54 */
    memcpy(mxGetPr(out[0]), mxGetPr(in[0]), sizeof(double) * n);
    if (nout > 1) {
      out[1] = mxCreateDoubleMatrix(n, n, mxREAL);
      memcpy(mxGetPr(out[1]), mxGetPr(in[1]), sizeof(double) * n * n);
    }
    if (nout > 2) out[2] = mxCreateDoubleScalar(0.0);
    return;
  }
  if (mxGetM(in[2]) != m || mxGetN(in[2]) != n) mexErrMsgTxt("update: H must be length(z) x n");
  if (mxGetM(in[3]) != m || mxGetN(in[3]) != m) mexErrMsgTxt("update: R must be length(z) x length(z)");
  if (mxGetNumberOfElements(in[5]) != m) mexErrMsgTxt("update: h must have as many entries as z");
  if (mxIsSparse(in[1])) mexErrMsgTxt("update: p_km1_k must be a full matrix");
  double *H = densify(in[2], m, n), *R = densify(in[3], m, m);
  mxArray *P = mxCreateDoubleMatrix(n, n, mxREAL);
  mxArray *K = mxCreateDoubleMatrix(n, m, mxREAL);
  const int rc = pre3_ekf_update_dense(pre3_mex_ctx(), (int)n, (int)m, mxGetPr(in[0]), mxGetPr(in[1]), H, R,
                                       mxGetPr(in[4]), mxGetPr(in[5]), mxGetPr(out[0]), mxGetPr(P), mxGetPr(K));
  mxFree(H);
  mxFree(R);
  pre3_mex_check(rc);
  if (nout > 1) out[1] = P;
  if (nout > 2) out[2] = K;
}
