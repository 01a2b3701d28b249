/* pre3_oracle_frames.c -- CPU restatement of the step before the matching path:
 * SR4000 frame -> filtered x, y, z maps -> per-feature 3-D points.
 *
 * TEST INFRASTRUCTURE ONLY (see pre3_oracle.c).  PARITY UNPINNED: the reference holds no vectors for this step
 * and MATLAB (fspecial / imfilter of the Image Processing Toolbox) is absent; checked against an independent
 * scipy.ndimage restatement (oracle/ref_numpy.py: read_xyz_sr4000, sift_extract_xyz) and by properties.
 *
 * Follows  M/read_xyz_sr4000.m:8-21, M/code_from_dr_ye/read_sr4000_data_dr_ye.m:8,88-90,
 *          M/inittialize_depth_my_version.m:16,31-85, M/SIFT_extract_save.m:55-56,75-88,
 *          M/code_from_dr_ye/confidence_filtering.m:1-13, ransac_dr_ye.m:13-19.
 * imfilter's summation order is not documented; the order fixed here (and in 3pre_b200/csrc/frames.cu) is
 *   acc = 0; for dc = -1..1, for dr = -1..1: acc = acc + h(dr,dc) * A(r+dr, c+dc)
 * with out-of-range taps reading 0.0 ('same', zero padding) or the clamped pixel ('replicate'). */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define FR_ROWS 144
#define FR_COLS 176

/* fspecial('gaussian',[3 3],sigma): h(dr+1 + 3*(dc+1)) */
ORC_API void orc_gaussian3(double sigma, double *h) {
  double mx = 0.0;
  for (int dc = -1; dc <= 1; ++dc)
    for (int dr = -1; dr <= 1; ++dr) {
      const double arg = -((double)(dc * dc) + (double)(dr * dr)) / (2.0 * sigma * sigma);
      const double v = exp(arg);
      h[(dr + 1) + 3 * (dc + 1)] = v;
      mx = v > mx ? v : mx;
    }
  double sum = 0.0;
  for (int i = 0; i < 9; ++i) {
    if (h[i] < 2.220446049250313e-16 * mx) h[i] = 0.0;
    sum = sum + h[i];
  }
  if (sum != 0.0)
    for (int i = 0; i < 9; ++i) h[i] = h[i] / sum;
}

static double fr_stencil(const double *map, int ld, int r, int c, const double *h, int boundary) {
  double acc = 0.0;
  for (int dc = -1; dc <= 1; ++dc)
    for (int dr = -1; dr <= 1; ++dr) {
      int rr = r + dr, cc = c + dc;
      double v;
      if (boundary == 1) {
        rr = rr < 0 ? 0 : (rr > FR_ROWS - 1 ? FR_ROWS - 1 : rr);
        cc = cc < 0 ? 0 : (cc > FR_COLS - 1 ? FR_COLS - 1 : cc);
        v = map[(size_t)cc * ld + rr];
      } else {
        v = (rr >= 0 && rr < FR_ROWS && cc >= 0 && cc < FR_COLS) ? map[(size_t)cc * ld + rr] : 0.0;
      }
      acc = acc + h[(dr + 1) + 3 * (dc + 1)] * v;
    }
  return acc;
}

/* [x,y,z] = read_xyz_sr4000 for one frame (sr: rows x 176 column-major); outputs 144 x 176 column-major */
ORC_API void orc_read_xyz(const double *sr, int rows, double sigma, int boundary, double *x, double *y, double *z) {
  double h[9];
  orc_gaussian3(sigma, h);
  for (int c = 0; c < FR_COLS; ++c)
    for (int r = 0; r < FR_ROWS; ++r) {
      z[(size_t)c * FR_ROWS + r] = fr_stencil(sr, rows, r, c, h, boundary);
      x[(size_t)c * FR_ROWS + r] = fr_stencil(sr + FR_ROWS, rows, r, c, h, boundary);
      y[(size_t)c * FR_ROWS + r] = fr_stencil(sr + 2 * FR_ROWS, rows, r, c, h, boundary);
    }
}

ORC_API double orc_max_confidence(const double *sr, int rows) {
  if (rows < 720) return NAN;
  double m = NAN;
  for (int c = 0; c < FR_COLS; ++c)
    for (int r = 0; r < FR_ROWS; ++r) {
      const double v = sr[(size_t)c * rows + 4 * FR_ROWS + r];
      if (v == v && (!(m == m) || v > m)) m = v;
    }
  return m;
}

/* The loop of SIFT_extract_save.m:75-88 (mode 0) or confidence_filtering + lookup (mode 1) for one frame.
 * frames: frame_ld x K (0-based x, y in rows 1:2).  xyz_all: 3 x K (NaN for rejected); keep: K; idx_remain: K
 * (0-based, -1 beyond the return value).  Returns the number of survivors; *n_oob counts features whose rounded
 * position is outside the image (index error in the reference). */
ORC_API int orc_features_xyz(const double *sr, int rows, double sigma, int boundary, int mode, int use_conf,
                             const double *frames, int frame_ld, int K, double *xyz_all, uint8_t *keep,
                             int32_t *idx_remain, int32_t *n_oob) {
  double h[9];
  orc_gaussian3(sigma, h);
  const int has_conf = rows >= 720;
  const double mc = orc_max_confidence(sr, rows);
  int nk = 0, oob = 0;
  for (int k = 0; k < K; ++k) {
    const double *fr = frames + (size_t)k * frame_ld;
    const long long c1 = (long long)round(fr[0] + 1.0), r1 = (long long)round(fr[1] + 1.0);
    int kp = 0;
    double px = NAN, py = NAN, pz = NAN;
    if (r1 >= 1 && r1 <= FR_ROWS && c1 >= 1 && c1 <= FR_COLS) {
      const int r = (int)r1 - 1, c = (int)c1 - 1;
      const double xf = fr_stencil(sr + FR_ROWS, rows, r, c, h, boundary);
      const double yf = fr_stencil(sr + 2 * FR_ROWS, rows, r, c, h, boundary);
      const double zf = fr_stencil(sr, rows, r, c, h, boundary);
      const double conf = has_conf ? sr[(size_t)c * rows + 4 * FR_ROWS + r] : 0.0;
      if (mode == 0) {
        if (!(xf != xf)) { /* ~isnan(x(...)) :40 */
          const double df = sqrt((xf * xf + yf * yf) + zf * zf); /* :45 */
          kp = !(df < 0.4 || (has_conf && conf <= (2.0 / 4.0) * mc)); /* :74 */
        }
      } else {
        kp = !(use_conf && has_conf && conf < 0.5 * mc); /* confidence_filtering.m:8 */
      }
      if (kp) {
        px = -xf; /* :85 */
        py = -yf;
        pz = zf;
      }
    } else {
      ++oob;
    }
    xyz_all[3 * (size_t)k] = px;
    xyz_all[3 * (size_t)k + 1] = py;
    xyz_all[3 * (size_t)k + 2] = pz;
    keep[k] = (uint8_t)kp;
    if (kp) idx_remain[nk++] = k;
  }
  for (int k = nk; k < K; ++k) idx_remain[k] = -1;
  if (n_oob) *n_oob = oob;
  return nk;
}
