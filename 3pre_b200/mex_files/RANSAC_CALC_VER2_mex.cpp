/* [R, T, error, BestFit, State_RANSAC] = RANSAC_CALC_VER2_mex(Ya, Yb, options [, samples | seed [, k [, adaptive [, method]]]])
 *
 * Fills the slot M/mex_files/RANSAC_CALCULATION/RANSAC_CALC.prj:95 declares (param.mex.outputfile
 * = RANSAC_CALC_VER2_mex; entry-point types :199-236: Ya, Yb double 3 x :3000, options struct of
 * DistanceThreshold and MaxIteration) -- the MATLAB-Coder build of which failed in the reference.
 * Algorithm: M/mex_files/RANSAC_CALCULATION/RANSAC_CALC_VER2.m:2-201 on the GPU (pre3_ransac).
 *
 * Optional trailing inputs (never required):
 *   in[3]  samples: k x H matrix of 1-based indices (what get_rand(k, N) marks, ascending), or a
 *          scalar seed for the built-in generator (MATLAB's RNG stream cannot be reproduced)
 *   in[4]  k (default 5, RANSAC_CALC_VER2.m:85)   in[5] adaptive stop (default 1)
 *   in[6]  method: 0 = find_transform_matrix (default), 1 = Horn (M/RANSAC_CALC_VER_test.m)
 * `error` is returned as the struct the reference ends up with (:199-201): fields ErrorSum,
 * mYa, mYb (the support set of the winner). */
#include "pre3_mex_common.h"

extern "C" void mexFunction(int nout, mxArray *out[], int nin, const mxArray *in[]) {
  if (nin < 3) mexErrMsgTxt("RANSAC_CALC_VER2: Ya, Yb and options are required");
  if (nout > 5) mexErrMsgTxt("Too many output arguments");
  for (int i = 0; i < 2; ++i)
    if (mxGetClassID(in[i]) != mxDOUBLE_CLASS || mxIsComplex(in[i]) || mxGetM(in[i]) != 3)
      mexErrMsgTxt("Ya and Yb must be real double 3 x N matrices");
  const int N = (int)mxGetN(in[0]);
  if ((int)mxGetN(in[1]) != N) mexErrMsgTxt("Ya and Yb must have the same size");
  if (!mxIsStruct(in[2])) mexErrMsgTxt("options must be a struct with DistanceThreshold and MaxIteration");
  const mxArray *fd = mxGetField(in[2], 0, "DistanceThreshold"), *fm = mxGetField(in[2], 0, "MaxIteration");
  if (!fd || !fm) mexErrMsgTxt("options must be a struct with DistanceThreshold and MaxIteration");
  pre3_ransac_opts o;
  memset(&o, 0, sizeof o);
  o.method = PRE3_METHOD_SVD;
  o.k = 5;
  o.max_iteration = (int32_t)mxGetScalar(fm);
  o.adaptive = 1;
  o.H = o.max_iteration;
  o.distance_threshold = mxGetScalar(fd);
  o.ratio = 1.5;
  o.seed = 0;
  int32_t *samples = NULL;
  if (nin > 4 && !mxIsEmpty(in[4])) o.k = (int32_t)mxGetScalar(in[4]);
  if (nin > 5 && !mxIsEmpty(in[5])) o.adaptive = mxGetScalar(in[5]) != 0.0;
  if (nin > 6 && !mxIsEmpty(in[6])) o.method = mxGetScalar(in[6]) != 0.0 ? PRE3_METHOD_HORN : PRE3_METHOD_SVD;
  if (nin > 3 && !mxIsEmpty(in[3])) {
    if (mxGetNumberOfElements(in[3]) == 1) {
      o.seed = (uint64_t)mxGetScalar(in[3]);
    } else {
      if (mxGetClassID(in[3]) != mxDOUBLE_CLASS) mexErrMsgTxt("samples must be a double k x H matrix");
      o.k = (int32_t)mxGetM(in[3]);
      o.H = (int32_t)mxGetN(in[3]);
      samples = (int32_t *)mxMalloc(sizeof(int32_t) * (size_t)o.k * (size_t)o.H);
      const double *sp = mxGetPr(in[3]);
      for (size_t i = 0; i < (size_t)o.k * (size_t)o.H; ++i) samples[i] = (int32_t)sp[i] - 1;
    }
  }
  pre3_pair_result res;
  uint8_t *mask = (uint8_t *)mxMalloc((size_t)(N > 0 ? N : 1));
  int rc = pre3_ransac(pre3_mex_ctx(), mxGetPr(in[0]), mxGetPr(in[1]), N, &o, samples, &res, mask, NULL, NULL);
  if (samples) mxFree(samples);
  if (rc != PRE3_OK) {
    mxFree(mask);
    pre3_mex_check(rc);
  }
  if (res.status == 1) {
    mxFree(mask);
    mexErrMsgIdAndTxt("pre3:get_rand", "get_rand: fewer correspondences than the minimal sample size");
  }
  if (res.status == 2) {
    mxFree(mask);
    mexErrMsgIdAndTxt("pre3:ransac", "RANSAC_CALC_VER2: no hypothesis was recorded");
  }
  out[0] = mxCreateDoubleMatrix(3, 3, mxREAL);
  memcpy(mxGetPr(out[0]), res.R, sizeof res.R);
  if (nout > 1) {
    out[1] = mxCreateDoubleMatrix(3, 1, mxREAL);
    memcpy(mxGetPr(out[1]), res.T, sizeof res.T);
  }
  if (nout > 2) {
    const char *names[3] = {"ErrorSum", "mYa", "mYb"};
    out[2] = mxCreateStructMatrix(1, 1, 3, names);
    mxArray *mYa = mxCreateDoubleMatrix(3, (size_t)res.best_fit, mxREAL);
    mxArray *mYb = mxCreateDoubleMatrix(3, (size_t)res.best_fit, mxREAL);
    const double *Ya = mxGetPr(in[0]), *Yb = mxGetPr(in[1]);
    double *pa = mxGetPr(mYa), *pb = mxGetPr(mYb);
    size_t j = 0;
    for (int i = 0; i < N && j < (size_t)res.best_fit; ++i)
      if (mask[i]) {
        memcpy(pa + 3 * j, Ya + 3 * (size_t)i, 3 * sizeof(double));
        memcpy(pb + 3 * j, Yb + 3 * (size_t)i, 3 * sizeof(double));
        ++j;
      }
    mxSetField(out[2], 0, "ErrorSum", mxCreateDoubleScalar(res.error_sum));
    mxSetField(out[2], 0, "mYa", mYa);
    mxSetField(out[2], 0, "mYb", mYb);
  }
  if (nout > 3) out[3] = mxCreateDoubleScalar((double)res.best_fit);
  if (nout > 4) out[4] = mxCreateDoubleScalar((double)res.state);
  mxFree(mask);
}
