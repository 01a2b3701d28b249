/* [hypothesis_support, positions_li_inliers_id, positions_li_inliers_euc] =
 *     compute_hypothesis_support_fast_mex(xi, cam, state_vector_pattern, z_id, z_euc, threshold)
 *
 * Drop-in for M/compute_hypothesis_support_fast.m:27-116 (same argument list; the .m shim of the
 * same base name forwards here).  xi may hold several hypothesised states as columns (n x B): the
 * outputs then have one column per state.  cam: struct with f, Cx, Cy, k1, k2 (M/initialize_cam.m:64-76).
 * Empty z_id / z_euc give [] masks like the reference (:73-77, :112-116). */
#include "pre3_mex_common.h"

static double cam_field(const mxArray *cam, const char *name) {
  const mxArray *f = mxGetField(cam, 0, name);
  if (!f) mexErrMsgIdAndTxt("pre3:arg", "cam.%s is missing", name);
  return mxGetScalar(f);
}

extern "C" void mexFunction(int nout, mxArray *out[], int nin, const mxArray *in[]) {
  if (nin != 6) mexErrMsgTxt("compute_hypothesis_support_fast: six input arguments required");
  if (nout > 3) mexErrMsgTxt("Too many output arguments");
  for (int i = 0; i < 6; ++i)
    if (i != 1 && !mxIsEmpty(in[i]) && (mxGetClassID(in[i]) != mxDOUBLE_CLASS || mxIsComplex(in[i])))
      mexErrMsgTxt("arguments must be real double arrays");
  if (!mxIsStruct(in[1])) mexErrMsgTxt("cam must be a struct");
  const int n = (int)mxGetM(in[0]), B = (int)mxGetN(in[0]);
  if ((int)mxGetM(in[2]) != n || (int)mxGetN(in[2]) != 4) mexErrMsgTxt("state_vector_pattern must be n x 4");
  const int n_id = mxIsEmpty(in[3]) ? 0 : (int)mxGetN(in[3]), n_euc = mxIsEmpty(in[4]) ? 0 : (int)mxGetN(in[4]);
  pre3_cam cam;
  cam.f = cam_field(in[1], "f");
  cam.Cx = cam_field(in[1], "Cx");
  cam.Cy = cam_field(in[1], "Cy");
  cam.k1 = cam_field(in[1], "k1");
  cam.k2 = cam_field(in[1], "k2");
  int32_t *support = (int32_t *)mxMalloc(sizeof(int32_t) * (size_t)(B > 0 ? B : 1));
  uint8_t *li = (uint8_t *)mxMalloc((size_t)B * (size_t)n_id + 1), *le = (uint8_t *)mxMalloc((size_t)B * (size_t)n_euc + 1);
  const int rc = pre3_ekf_support(pre3_mex_ctx(), mxGetPr(in[0]), n, B, &cam, mxGetPr(in[2]),
                                  n_id ? mxGetPr(in[3]) : NULL, n_id, n_euc ? mxGetPr(in[4]) : NULL, n_euc,
                                  mxGetScalar(in[5]), support, li, le);
  if (rc != PRE3_OK) {
    mxFree(support);
    mxFree(li);
    mxFree(le);
    pre3_mex_check(rc);
  }
  out[0] = mxCreateDoubleMatrix(1, (size_t)B, mxREAL);
  for (int b = 0; b < B; ++b) mxGetPr(out[0])[b] = (double)support[b];
  if (nout > 1) {
    out[1] = mxCreateLogicalMatrix(n_id ? (size_t)B : 0, (size_t)n_id); /* one ROW per state: 1 x n_id for B = 1 */
    unsigned char *p = (unsigned char *)mxGetData(out[1]);
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < n_id; ++j) p[(size_t)j * B + b] = li[(size_t)b * n_id + j];
  }
  if (nout > 2) {
    out[2] = mxCreateLogicalMatrix(n_euc ? (size_t)B : 0, (size_t)n_euc);
    unsigned char *p = (unsigned char *)mxGetData(out[2]);
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < n_euc; ++j) p[(size_t)j * B + b] = le[(size_t)b * n_euc + j];
  }
  mxFree(support);
  mxFree(li);
  mxFree(le);
}
