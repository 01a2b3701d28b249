/* [x, y, z, confidence_map, max_confidence] = read_xyz_sr4000_mex(sr_data [, sigma [, boundary]])
 * [xyz_data, idxRemain] = read_xyz_sr4000_mex(sr_data, frames, 'features' [, mode [, use_confidence]])
 *
 * The arithmetic of M/read_xyz_sr4000.m:8-26 (first form: x, y, z = imfilter(., fspecial('gaussian',[3 3],sigma),
 * 'same'); sigma defaults to 2, boundary 0 = zero padding, 1 = 'replicate' as in
 * M/code_from_dr_ye/read_sr4000_data_dr_ye.m:8,88-90) and of the per-feature loop of M/SIFT_extract_save.m:75-88 with
 * M/inittialize_depth_my_version.m:16,40-85 (second form: frames = 4 x N as sift returns them, 0-based positions;
 * xyz_data 3 x N with NaN columns for rejected features, idxRemain 1-based; mode 1 = the code_from_dr_ye flavour) on
 * the GPU (pre3_read_xyz_sr4000_batch / pre3_features_xyz_batch, one frame).  sr_data = load('d1_%04d.dat'): 576,
 * 720 or 721 rows x 176 columns.  Called by matlab/read_xyz_sr4000.m. */
#include "pre3_mex_common.h"

extern "C" void mexFunction(int nout, mxArray *out[], int nin, const mxArray *in[]) {
  if (nin < 1) mexErrMsgTxt("read_xyz_sr4000_mex: sr_data is required");
  if (mxGetClassID(in[0]) != mxDOUBLE_CLASS || mxIsComplex(in[0]) || mxGetN(in[0]) != 176)
    mexErrMsgTxt("sr_data must be a real double matrix with 176 columns");
  const int rows = (int)mxGetM(in[0]);
  if (rows != 576 && rows != 720 && rows != 721) mexErrMsgTxt("sr_data must have 576, 720 or 721 rows");
  pre3_frame_opts o;
  memset(&o, 0, sizeof o);
  o.sigma = 2.0;
  o.rows = rows;
  o.use_confidence = 1;
  const bool features = nin >= 3;
  if (!features) {
    if (nout > 5) mexErrMsgTxt("Too many output arguments");
    if (nin > 1 && !mxIsEmpty(in[1])) o.sigma = mxGetScalar(in[1]);
    if (nin > 2 && !mxIsEmpty(in[2])) o.boundary = mxGetScalar(in[2]) != 0.0;
    mxArray *m[3];
    for (int i = 0; i < 3; ++i) m[i] = mxCreateDoubleMatrix(144, 176, mxREAL);
    double mc = 0.0;
    pre3_mex_check(pre3_read_xyz_sr4000_batch(pre3_mex_ctx(), mxGetPr(in[0]), 1, &o, mxGetPr(m[0]), mxGetPr(m[1]),
                                              mxGetPr(m[2]), &mc));
    out[0] = m[0];
    if (nout > 1) out[1] = m[1];
    if (nout > 2) out[2] = m[2];
    if (nout > 3) { /* confidence_map = sr_data(577:720, :) raw, or [] (read_xyz_sr4000.m:25-34) */
      if (rows >= 720) {
        out[3] = mxCreateDoubleMatrix(144, 176, mxREAL);
        const double *s = mxGetPr(in[0]);
        double *d = mxGetPr(out[3]);
        for (int c = 0; c < 176; ++c) memcpy(d + 144 * (size_t)c, s + (size_t)c * rows + 576, 144 * sizeof(double));
      } else {
        out[3] = mxCreateDoubleMatrix(0, 0, mxREAL);
      }
    }
    if (nout > 4) out[4] = mxCreateDoubleScalar(mc);
    return;
  }
  if (nout > 2) mexErrMsgTxt("Too many output arguments");
  if (mxGetClassID(in[1]) != mxDOUBLE_CLASS || mxGetM(in[1]) < 2) mexErrMsgTxt("frames must be a double matrix with >= 2 rows");
  const int ld = (int)mxGetM(in[1]), K = (int)mxGetN(in[1]);
  if (nin > 3 && !mxIsEmpty(in[3])) o.mode = mxGetScalar(in[3]) != 0.0;
  if (nin > 4 && !mxIsEmpty(in[4])) o.use_confidence = mxGetScalar(in[4]) != 0.0;
  if (o.mode == 1) {
    o.sigma = 1.0;
    o.boundary = 1;
  }
  out[0] = mxCreateDoubleMatrix(3, (size_t)K, mxREAL);
  int32_t *idx = (int32_t *)mxMalloc(sizeof(int32_t) * (size_t)(K > 0 ? K : 1));
  int32_t nk = 0, oob = 0;
  const int rc = pre3_features_xyz_batch(pre3_mex_ctx(), mxGetPr(in[0]), 1, &o, mxGetPr(in[1]), ld, K, NULL,
                                         mxGetPr(out[0]), NULL, &nk, idx, NULL, NULL, 0, 0, NULL, NULL, &oob);
  if (rc != PRE3_OK || oob) {
    mxFree(idx);
    pre3_mex_check(rc);
    mexErrMsgTxt("Index exceeds matrix dimensions."); /* x(round(uv(1)),round(uv(2))), inittialize_depth_my_version.m:40 */
  }
  if (nout > 1) {
    out[1] = mxCreateDoubleMatrix(1, (size_t)nk, mxREAL);
    double *d = mxGetPr(out[1]);
    for (int i = 0; i < nk; ++i) d[i] = (double)(idx[i] + 1);
  }
  mxFree(idx);
}
