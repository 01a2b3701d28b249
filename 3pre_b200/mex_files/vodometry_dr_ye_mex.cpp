/* [rot, trans, sta, op_num, good, stat] = vodometry_dr_ye_mex(pset1, pset2, match [, samples | seed [, MaxIteration]])
 *
 * The RANSAC part of M/code_from_dr_ye/vodometry_dr_ye.m:147-220 (loop body ransac_dr_ye.m:20-71, refit
 * find_transform_matrix_dr_ye.m) on the GPU (pre3_vodometry_dr_ye_batch, P = 1) -- the variant the live
 * EKF calls (M/fv.m:47 -> Calculate_V_Omega_RANSAC_dr_ye.m:19-22).  Called by matlab/vodometry_dr_ye.m,
 * which keeps the reference's file reading, SIFT extraction and output packing.
 *   pset1, pset2  3 x pnum double: [-x(ROW,COL); -y(ROW,COL); z(ROW,COL)] of every match (ransac_dr_ye.m:13-19)
 *   match         2 x pnum double, what siftmatch returned (only compared for equality by the sampler)
 *   in[3]         4 x H matrix of 1-based draws num_rs(1..4) per iteration, or a scalar seed (MATLAB's RNG
 *                 stream cannot be reproduced outside MATLAB), optional
 *   in[4]         MaxIteration (default 700, vodometry_dr_ye.m:162), optional
 * Outputs: rot 3 x 3, trans 3 x 1, sta (state of the refit; 4 = fewer than 4 matches or no consensus,
 * :158,:191), op_num, good (1 x op_num 1-based columns of match in the support set), stat = struct of
 * nIterationRansac, ErrorMean, ErrorStd (:214-216), dist, nLoops. */
#include "pre3_mex_common.h"

extern "C" void mexFunction(int nout, mxArray *out[], int nin, const mxArray *in[]) {
  if (nin < 3) mexErrMsgTxt("vodometry_dr_ye_mex: pset1, pset2 and match are required");
  if (nout > 6) mexErrMsgTxt("Too many output arguments");
  for (int i = 0; i < 2; ++i)
    if (mxGetClassID(in[i]) != mxDOUBLE_CLASS || mxIsComplex(in[i]) || mxGetM(in[i]) != 3)
      mexErrMsgTxt("pset1 and pset2 must be real double 3 x pnum matrices");
  const int N = (int)mxGetN(in[0]);
  if ((int)mxGetN(in[1]) != N) mexErrMsgTxt("pset1 and pset2 must have the same size");
  if (mxGetClassID(in[2]) != mxDOUBLE_CLASS || (N > 0 && (mxGetM(in[2]) != 2 || (int)mxGetN(in[2]) != N)))
    mexErrMsgTxt("match must be a double 2 x pnum matrix");
  pre3_ransac_opts o;
  memset(&o, 0, sizeof o);
  o.method = PRE3_METHOD_DR_YE;
  o.k = 4;
  o.max_iteration = 700;
  o.ratio = 1.5;
  if (nin > 4 && !mxIsEmpty(in[4])) o.max_iteration = (int32_t)mxGetScalar(in[4]);
  o.H = o.max_iteration;
  int32_t *samples = NULL;
  if (nin > 3 && !mxIsEmpty(in[3])) {
    if (mxGetNumberOfElements(in[3]) == 1) {
      o.seed = (uint64_t)mxGetScalar(in[3]);
    } else {
      if (mxGetClassID(in[3]) != mxDOUBLE_CLASS || mxGetM(in[3]) != 4)
        mexErrMsgTxt("samples must be a double 4 x H matrix");
      o.H = (int32_t)mxGetN(in[3]);
      samples = (int32_t *)mxMalloc(sizeof(int32_t) * 4 * (size_t)o.H);
      const double *sp = mxGetPr(in[3]);
      for (size_t i = 0; i < 4 * (size_t)o.H; ++i) samples[i] = (int32_t)sp[i] - 1;
    }
  }
  int32_t *match = (int32_t *)mxMalloc(sizeof(int32_t) * 2 * (size_t)(N > 0 ? N : 1));
  const double *mp = mxGetPr(in[2]);
  for (size_t i = 0; i < 2 * (size_t)N; ++i) match[i] = (int32_t)mp[i];
  uint8_t *mask = (uint8_t *)mxMalloc((size_t)(N > 0 ? N : 1));
  pre3_pair_result res;
  pre3_dr_ye_stat st;
  int rc = pre3_vodometry_dr_ye_batch(pre3_mex_ctx(), mxGetPr(in[0]), mxGetPr(in[1]), NULL, match, 1, N, &o, samples,
                                      &res, mask, &st, NULL);
  if (samples) mxFree(samples);
  mxFree(match);
  if (rc != PRE3_OK) {
    mxFree(mask);
    pre3_mex_check(rc);
  }
  if (res.status == 5) {
    mxFree(mask);
    mexErrMsgIdAndTxt("pre3:dr_ye", "ransac_dr_ye: no point farther than 0.4 m (min of an empty set)");
  }
  const int failed = res.status != 0; /* pnum < 4 or no consensus: SolutionState 4 */
  const int op_num = failed ? (res.status == 4 ? res.best_fit : 0) : res.best_fit;
  out[0] = mxCreateDoubleMatrix(3, 3, mxREAL);
  if (!failed) memcpy(mxGetPr(out[0]), res.R, sizeof res.R);
  if (nout > 1) {
    out[1] = mxCreateDoubleMatrix(3, 1, mxREAL);
    if (!failed) memcpy(mxGetPr(out[1]), res.T, sizeof res.T);
  }
  if (nout > 2) out[2] = mxCreateDoubleScalar(failed ? 4.0 : (double)res.state);
  if (nout > 3) out[3] = mxCreateDoubleScalar((double)op_num);
  if (nout > 4) {
    const int ng = failed ? 0 : res.best_fit;
    out[4] = mxCreateDoubleMatrix(1, (size_t)ng, mxREAL);
    double *g = mxGetPr(out[4]);
    int j = 0;
    for (int i = 0; i < N && j < ng; ++i)
      if (mask[i]) g[j++] = (double)(i + 1);
  }
  if (nout > 5) {
    const char *names[5] = {"nIterationRansac", "ErrorMean", "ErrorStd", "dist", "nLoops"};
    out[5] = mxCreateStructMatrix(1, 1, 5, names);
    mxSetField(out[5], 0, "nIterationRansac", mxCreateDoubleScalar((double)st.n_iteration_ransac));
    mxSetField(out[5], 0, "ErrorMean", mxCreateDoubleScalar(st.error_mean));
    mxSetField(out[5], 0, "ErrorStd", mxCreateDoubleScalar(st.error_std));
    mxSetField(out[5], 0, "dist", mxCreateDoubleScalar(st.dist));
    mxSetField(out[5], 0, "nLoops", mxCreateDoubleScalar((double)st.n_loops));
  }
  mxFree(mask);
}
