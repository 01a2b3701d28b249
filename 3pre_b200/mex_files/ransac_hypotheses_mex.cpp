/* [low_innovation_inlier, stats] = ransac_hypotheses_mex(x, P, std_z, cam, type, pos, has_z, ic, z, h,
 *                                                        Hcam, Hfeat, R, li0 [, sel | seed [, n_hyp [, adaptive]]])
 *
 * The compute of M/ransac_hypotheses.m:27-85 on plain arrays.  The old-style @ekf_filter object and
 * the features_info struct array stay on the MATLAB side: the shim mex_files/matlab/ransac_hypotheses.m
 * (same signature as the reference) reads them through get_x_k_km1 / get_p_k_km1 / get_std_z
 * (M/@ekf_filter/) and forwards:
 *   x n x 1, P n x n, std_z scalar, cam struct (f, Cx, Cy, k1, k2),
 *   type 1 x F (0 'inversedepth', 1 'cartesian'), pos 1 x F (1-based first state of the feature),
 *   has_z, ic, li0 1 x F (0/1), z, h 2 x F, Hcam 2 x 13 x F (= H(:,1:13)), Hfeat 2 x 6 x F
 *   (= H(:,pos:pos+5), cartesian: first 3 columns), R 2 x 2 x F,
 *   sel (optional) 3 x H matrix of 1-based feature positions (select_random_match.m:58), or a scalar seed.
 * Outputs: low_innovation_inlier 1 x F; stats = [n_hyp, max_support, best_hyp (1-based, 0 none),
 * hypotheses evaluated, num_IC]. */
#include "pre3_mex_common.h"

static const double *dbl(const mxArray *a, size_t count, const char *what) {
  if (mxGetClassID(a) != mxDOUBLE_CLASS || mxIsComplex(a) || mxGetNumberOfElements(a) != count)
    mexErrMsgIdAndTxt("pre3:arg", "ransac_hypotheses: %s has the wrong class or size", what);
  return mxGetPr(a);
}

extern "C" void mexFunction(int nout, mxArray *out[], int nin, const mxArray *in[]) {
  if (nin < 14) mexErrMsgTxt("ransac_hypotheses_mex: 14 input arguments required");
  if (nout > 2) mexErrMsgTxt("Too many output arguments");
  const int n = (int)mxGetNumberOfElements(in[0]);
  const int F = (int)mxGetNumberOfElements(in[4]);
  const double *x = dbl(in[0], (size_t)n, "x"), *P = dbl(in[1], (size_t)n * n, "P");
  const double std_z = mxGetScalar(in[2]);
  if (!mxIsStruct(in[3])) mexErrMsgTxt("cam must be a struct");
  pre3_cam cam;
  const char *names[5] = {"f", "Cx", "Cy", "k1", "k2"};
  double *cv[5] = {&cam.f, &cam.Cx, &cam.Cy, &cam.k1, &cam.k2};
  for (int i = 0; i < 5; ++i) {
    const mxArray *f = mxGetField(in[3], 0, names[i]);
    if (!f) mexErrMsgIdAndTxt("pre3:arg", "cam.%s is missing", names[i]);
    *cv[i] = mxGetScalar(f);
  }
  const double *type = dbl(in[4], (size_t)F, "type"), *pos = dbl(in[5], (size_t)F, "pos");
  const double *has_z = dbl(in[6], (size_t)F, "has_z"), *ic = dbl(in[7], (size_t)F, "ic");
  const double *z = dbl(in[8], 2 * (size_t)F, "z"), *h = dbl(in[9], 2 * (size_t)F, "h");
  const double *Hcam = dbl(in[10], 26 * (size_t)F, "Hcam"), *Hfeat = dbl(in[11], 12 * (size_t)F, "Hfeat");
  const double *R = dbl(in[12], 4 * (size_t)F, "R"), *li0 = dbl(in[13], (size_t)F, "li0");
  pre3_ekf_opts o;
  memset(&o, 0, sizeof o);
  o.n_hyp_init = 1000; /* ransac_hypotheses.m:35 */
  o.H = 1000;
  o.adaptive = 1;
  int32_t *sel = NULL;
  if (nin > 15 && !mxIsEmpty(in[15])) o.n_hyp_init = o.H = (int32_t)mxGetScalar(in[15]);
  if (nin > 16 && !mxIsEmpty(in[16])) o.adaptive = mxGetScalar(in[16]) != 0.0;
  if (nin > 14 && !mxIsEmpty(in[14])) {
    if (mxGetNumberOfElements(in[14]) == 1) {
      o.seed = (uint64_t)mxGetScalar(in[14]);
    } else {
      const int rows = (int)mxGetM(in[14]);
      if (mxGetClassID(in[14]) != mxDOUBLE_CLASS || rows < 1 || rows > 3) mexErrMsgTxt("sel must be a double m x H matrix, m <= 3");
      o.H = (int32_t)mxGetN(in[14]);
      sel = (int32_t *)mxMalloc(sizeof(int32_t) * 3 * (size_t)(o.H > 0 ? o.H : 1));
      const double *sp = mxGetPr(in[14]);
      for (int i = 0; i < o.H; ++i)
        for (int a = 0; a < 3; ++a) sel[3 * (size_t)i + a] = a < rows ? (int32_t)sp[(size_t)i * rows + a] - 1 : 0;
    }
  }
  int32_t *ti = (int32_t *)mxMalloc(sizeof(int32_t) * 2 * (size_t)(F > 0 ? F : 1)), *pi = ti + F;
  uint8_t *fl = (uint8_t *)mxMalloc(3 * (size_t)(F > 0 ? F : 1)), *hz = fl, *icb = fl + F, *li = fl + 2 * F;
  for (int i = 0; i < F; ++i) {
    ti[i] = (int32_t)type[i];
    pi[i] = (int32_t)pos[i] - 1;
    hz[i] = has_z[i] != 0.0;
    icb[i] = ic[i] != 0.0;
    li[i] = li0[i] != 0.0;
  }
  pre3_ekf_result res;
  const int rc = pre3_ransac_hypotheses_batch(pre3_mex_ctx(), 1, n, F, x, P, std_z, &cam, ti, pi, hz, icb, z, h, Hcam,
                                              Hfeat, R, sel, &o, 0, li, &res, NULL);
  if (sel) mxFree(sel);
  if (rc != PRE3_OK) {
    mxFree(ti);
    mxFree(fl);
    pre3_mex_check(rc);
  }
  if (res.status == 1) {
    mxFree(ti);
    mxFree(fl);
    mexErrMsgIdAndTxt("pre3:select_random_match", "Index exceeds matrix dimensions (no individually compatible match)");
  }
  if (res.status == 3) {
    mxFree(ti);
    mxFree(fl);
    mexErrMsgIdAndTxt("pre3:ekf", "an individually compatible feature has no measurement z");
  }
  out[0] = mxCreateDoubleMatrix(1, (size_t)F, mxREAL);
  for (int i = 0; i < F; ++i) mxGetPr(out[0])[i] = (double)li[i];
  if (nout > 1) {
    out[1] = mxCreateDoubleMatrix(1, 5, mxREAL);
    double *s = mxGetPr(out[1]);
    s[0] = res.n_hyp;
    s[1] = (double)res.max_support;
    s[2] = (double)(res.best_hyp + 1);
    s[3] = (double)res.n_evaluated;
    s[4] = (double)res.num_ic;
  }
  mxFree(ti);
  mxFree(fl);
}
