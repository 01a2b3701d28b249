/* pre3_oracle_ekf.c -- CPU restatement of the 1-point-RANSAC EKF hypothesis path of 3PRE
 * (BASELINE.json config 4): ransac_hypotheses + compute_hypothesis_support_fast.
 *
 * TEST INFRASTRUCTURE ONLY (see pre3_oracle.c): imported by tests/, smoke() and bench.py's
 * cpu_baseline / --impl reference legs, never by the product path.
 *
 * Parity status: "parity unpinned".  The reference is MATLAB source (no interpreter here) and
 * ships no golden vectors for this path.  The restatement follows the cited
 * .m files statement by statement; MATLAB built-ins whose arithmetic cannot be pinned are
 * replaced by the fixed-order algorithms SPECIFIED below (the CUDA kernels in
 * 3pre_b200/csrc/ekf.cu implement the same operation order, so supports / masks / selection
 * are comparable bit for bit):
 *   sin, cos  -> orc_sincos   (Cody-Waite reduction + the classic degree-13/14 kernels)
 *   inv(S)    -> orc_inv      (Gauss-Jordan, partial pivoting)
 *   sparse * dense products -> sums over the structural non-zeros in ascending index order
 * oracle/ref_numpy_ekf.py is the independent restatement (numpy sin/cos, LAPACK inv, dense
 * products) that cross-checks the states to 1e-9 and the supports away from the threshold.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile); only + - * / sqrt and the
 * functions defined here touch the data.  `M/` = /root/reference/matlab_code/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* sin / cos (stand in for MATLAB sin/cos in M/m.m:32-34)
 *
 * SPEC (shared with ekf.cu):
 *   |x| <= pi/4:  s = ksin(x, 0), c = kcos(x, 0)
 *   otherwise     fn = trunc(x * (2/pi) + (x >= 0 ? 0.5 : -0.5))      (|x| < 2^20 * pi/2, else NaN)
 *                 r = (x - fn*P1) - fn*P2 ;  w = fn*P2T ;  y0 = r - w ;  y1 = (r - y0) - w
 *                 quadrant n = fn mod 4 selects (+-ksin, +-kcos)
 *   ksin(x,y): z = x*x; v = z*x; r = S2 + z*(S3 + z*(S4 + z*(S5 + z*S6)));
 *              y == 0 ? x + v*(S1 + z*r) : x - ((z*(0.5*y - v*r) - y) - v*S1)
 *   kcos(x,y): z = x*x; r = z*(C1 + z*(C2 + z*(C3 + z*(C4 + z*(C5 + z*C6)))));
 *              hz = 0.5*z; w = 1 - hz; w + (((1 - w) - hz) + (z*r - x*y))
 *   P1, P2 carry 33 significant bits of pi/2 each (fn*P1, fn*P2 are exact), P2T the tail.
 * Coefficients are the classic minimax ones (Sun fdlibm k_sin.c / k_cos.c, public).  Error
 * < 1 ulp; tests compare with libm. */
static const double ORC_S1 = -1.66666666666666324348e-01, ORC_S2 = 8.33333333332248946124e-03,
                    ORC_S3 = -1.98412698298579493134e-04, ORC_S4 = 2.75573137070700676789e-06,
                    ORC_S5 = -2.50507602534068634195e-08, ORC_S6 = 1.58969099521155010221e-10;
static const double ORC_C1 = 4.16666666666666019037e-02, ORC_C2 = -1.38888888888741095749e-03,
                    ORC_C3 = 2.48015872894767294178e-05, ORC_C4 = -2.75573143513906633035e-07,
                    ORC_C5 = 2.08757232129817482790e-09, ORC_C6 = -1.13596475577881948265e-11;
static const double ORC_INVPIO2 = 6.36619772367581382433e-01, ORC_P1 = 1.57079632673412561417e+00,
                    ORC_P2 = 6.07710050630396597660e-11, ORC_P2T = 2.02226624879595063154e-21;

static double orc_ksin(double x, double y) {
  const double z = x * x, v = z * x;
  const double r = ORC_S2 + z * (ORC_S3 + z * (ORC_S4 + z * (ORC_S5 + z * ORC_S6)));
  if (y == 0.0) return x + v * (ORC_S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * ORC_S1);
}

static double orc_kcos(double x, double y) {
  const double z = x * x;
  const double r = z * (ORC_C1 + z * (ORC_C2 + z * (ORC_C3 + z * (ORC_C4 + z * (ORC_C5 + z * ORC_C6)))));
  const double hz = 0.5 * z, w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * r - x * y));
}

ORC_API void orc_sincos(double x, double *s, double *c) {
  if (!(fabs(x) < 1647099.0)) { /* 2^20 * pi/2, also catches NaN / Inf */
    *s = *c = NAN;
    return;
  }
  if (fabs(x) <= 0.78539816339744830962) {
    *s = orc_ksin(x, 0.0);
    *c = orc_kcos(x, 0.0);
    return;
  }
  const double fn = (double)(long long)(x * ORC_INVPIO2 + (x >= 0.0 ? 0.5 : -0.5));
  const double r = (x - fn * ORC_P1) - fn * ORC_P2;
  const double w = fn * ORC_P2T;
  const double y0 = r - w;
  const double y1 = (r - y0) - w;
  const double ks = orc_ksin(y0, y1), kc = orc_kcos(y0, y1);
  switch ((int)(((long long)fn) & 3)) {
    case 0: *s = ks; *c = kc; break;
    case 1: *s = kc; *c = -ks; break;
    case 2: *s = -ks; *c = -kc; break;
    default: *s = -kc; *c = ks; break;
  }
}

/* ------------------------------------------------------------------------- */
/* inv(S) (M/ransac_hypotheses.m:62), S of order d in {2, 6}, row-major a[r*d+c].
 *
 * SPEC (shared with ekf.cu): Gauss-Jordan on [S | I] with partial pivoting:
 *   for col = 0..d-1: pivot row = first row >= col with the largest |a[row][col]|; swap;
 *     piv = a[col][col]; row col := row col / piv (every entry, both halves);
 *     for every other row r: f = a[r][col]; row r := row r - f * row col.
 * A zero pivot yields Inf/NaN exactly like a singular inv() would poison K. */
static void orc_inv(const double *S, int d, double *out) {
  double a[6][12];
  for (int r = 0; r < d; ++r)
    for (int c = 0; c < d; ++c) {
      a[r][c] = S[r * d + c];
      a[r][d + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int col = 0; col < d; ++col) {
    int pr = col;
    double best = fabs(a[col][col]);
    for (int r = col + 1; r < d; ++r)
      if (fabs(a[r][col]) > best) {
        best = fabs(a[r][col]);
        pr = r;
      }
    if (pr != col)
      for (int c = 0; c < 2 * d; ++c) {
        double t = a[col][c];
        a[col][c] = a[pr][c];
        a[pr][c] = t;
      }
    const double piv = a[col][col];
    for (int c = 0; c < 2 * d; ++c) a[col][c] = a[col][c] / piv;
    for (int r = 0; r < d; ++r) {
      if (r == col) continue;
      const double f = a[r][col];
      for (int c = 0; c < 2 * d; ++c) a[r][c] = a[r][c] - f * a[col][c];
    }
  }
  for (int r = 0; r < d; ++r)
    for (int c = 0; c < d; ++c) out[r * d + c] = a[r][d + c];
}

/* ------------------------------------------------------------------------- */
/* Feature table of one frame: what features_info carries into ransac_hypotheses
 * (M/ransac_hypotheses.m:38,48,54-60).  All matrices column-major like MATLAB.
 *   type[i]   0 = 'inversedepth' (6 states), 1 = 'cartesian' (3 states)
 *   pos[i]    0-based offset of the feature's states in x (generate_state_vector_pattern.m:30-51
 *             walks position = 14, +6 / +3)
 *   has_z[i]  ~isempty(features_info(i).z)           ic[i]  individually_compatible
 *   z, h      2 x F        Hcam 2 x 13 x F  (H(:,1:13))      Hfeat 2 x 6 x F (H(:, pos+1:pos+nf);
 *             cartesian uses the first 3 columns)            R 2 x 2 x F
 * Structural non-zeros of a row of H: columns 0..12 and pos..pos+nf-1
 * (M/calculate_Hi_inverse_depth_my_version.m:44-49). */
typedef struct {
  double f, Cx, Cy, k1, k2; /* M/initialize_cam.m:64-76 */
} orc_cam;

static int nf_of(int type) { return type == 0 ? 6 : 3; }

/* column index of structural non-zero number t (0..12+nf) of feature i */
static int nz_col(const int32_t *pos, const int32_t *type, int i, int t) { return t < 13 ? t : pos[i] + (t - 13); }
static double nz_val(const double *Hcam, const double *Hfeat, int i, int c, int t) {
  return t < 13 ? Hcam[(size_t)i * 26 + 2 * t + c] : Hfeat[(size_t)i * 12 + 2 * (t - 13) + c];
}

/* xi = x + K*(zi - hi), K = P*Hi'*inv(Hi*P*Hi' + R)  (M/ransac_hypotheses.m:54-63) for the
 * m features sel[0..m) (0-based feature indices, in the order select_random_match returned them).
 *
 * SPEC of the operation order (shared with ekf.cu); rows of Hi are (feature a, component c) in
 * stacking order, r = 2*a + c:
 *   W[r][j]  = sum_{t} Hi[r][k_t] * P[k_t][j]          (Hi*P, left to right over the non-zeros)
 *   S[r][s]  = (sum_{t} W[r][k_t] * Hi[s][k_t]) + Rblk[r][s]   (k_t: non-zeros of row s)
 *   G[e][r]  = sum_{t} P[e][k_t] * Hi[r][k_t]          (P*Hi')
 *   Sinv     = orc_inv(S)
 *   K[e][j]  = sum_{c=0..2m-1} G[e][c] * Sinv[c][j]
 *   xi[e]    = x[e] + sum_{j=0..2m-1} K[e][j] * (zi[j] - hi[j])
 * every sum accumulated left to right starting from its first term. */
ORC_API void orc_ekf_update(const double *x, const double *P, int n, const int32_t *pos, const int32_t *type,
                            const double *z, const double *h, const double *Hcam, const double *Hfeat,
                            const double *R, const int32_t *sel, int m, double *xi) {
  const int d = 2 * m;
  double S[36], Sinv[36], innov[6];
  /* S */
  for (int a = 0; a < m; ++a)
    for (int ca = 0; ca < 2; ++ca) {
      const int fa = sel[a], r = 2 * a + ca, nta = 13 + nf_of(type[fa]);
      for (int b = 0; b < m; ++b)
        for (int cb = 0; cb < 2; ++cb) {
          const int fb = sel[b], s = 2 * b + cb, ntb = 13 + nf_of(type[fb]);
          double acc = 0.0;
          for (int tb = 0; tb < ntb; ++tb) {
            const int k = nz_col(pos, type, fb, tb);
            double w = 0.0; /* W[r][k] */
            for (int ta = 0; ta < nta; ++ta) {
              const int kk = nz_col(pos, type, fa, ta);
              const double term = nz_val(Hcam, Hfeat, fa, ca, ta) * P[(size_t)k * n + kk];
              w = (ta == 0) ? term : w + term;
            }
            const double term = w * nz_val(Hcam, Hfeat, fb, cb, tb);
            acc = (tb == 0) ? term : acc + term;
          }
          const double rblk = (a == b) ? R[(size_t)fa * 4 + 2 * cb + ca] : 0.0;
          S[r * d + s] = acc + rblk;
        }
    }
  orc_inv(S, d, Sinv);
  for (int a = 0; a < m; ++a)
    for (int c = 0; c < 2; ++c) innov[2 * a + c] = z[(size_t)sel[a] * 2 + c] - h[(size_t)sel[a] * 2 + c];
  for (int e = 0; e < n; ++e) {
    double G[6], K[6];
    for (int a = 0; a < m; ++a)
      for (int c = 0; c < 2; ++c) {
        const int fa = sel[a], nta = 13 + nf_of(type[fa]);
        double g = 0.0;
        for (int t = 0; t < nta; ++t) {
          const int k = nz_col(pos, type, fa, t);
          const double term = P[(size_t)k * n + e] * nz_val(Hcam, Hfeat, fa, c, t);
          g = (t == 0) ? term : g + term;
        }
        G[2 * a + c] = g;
      }
    for (int j = 0; j < d; ++j) {
      double k = 0.0;
      for (int c = 0; c < d; ++c) {
        const double term = G[c] * Sinv[c * d + j];
        k = (c == 0) ? term : k + term;
      }
      K[j] = k;
    }
    double dx = 0.0;
    for (int j = 0; j < d; ++j) {
      const double term = K[j] * innov[j];
      dx = (j == 0) ? term : dx + term;
    }
    xi[e] = x[e] + dx;
  }
}

/* ------------------------------------------------------------------------- */
/* compute_hypothesis_support_fast (M/compute_hypothesis_support_fast.m:27-116).
 * Index lists replace the logical pattern columns (:35-37,:83): idx_r 3*n_id, idx_ang 2*n_id,
 * idx_rho n_id, idx_xyz 3*n_euc, each in state order (0-based).  Returns the support; li_id /
 * li_euc receive the logical masks.
 *
 * SPEC of the operation order (shared with ekf.cu):
 *   q2r (M/q2r.m:29-36), r=q0 x=q1 y=q2 z=q3, every entry left to right, e.g.
 *     R00 = ((r*r + x*x) - y*y) - z*z ;  R01 = 2*(x*y - r*z) ; ...
 *   rotcw = rotwc' ;  hc_i = (rotwc[0][i]*v0 + rotwc[1][i]*v1) + rotwc[2][i]*v2
 *   inverse depth: (s_t,c_t) = sincos(theta), (s_p,c_p) = sincos(phi); mi = [c_p*s_t; -s_p; c_p*c_t]
 *     (M/m.m:32-34);  v = (ri - rwc)*rho + mi        (:49-55)
 *   cartesian: v = xyz - rwc                          (:92-94)
 *   hn = hc(1:2)/hc(3);  uv = f*hn + [u0;v0]          (:57-64)
 *   distort (M/distort_fm_my_version.m:52-61): xu = (u-Cx)/f; yu = (v-Cy)/f; ru = sqrt(xu*xu+yu*yu);
 *     ru2 = ru*ru; D = (1 + k1*ru2) + k2*(ru2*ru2); ud = (xu*D)*f + Cx; vd = (yu*D)*f + Cy
 *   residual = sqrt(nu0*nu0 + nu1*nu1), nu = z - [ud;vd]   (:68-69)
 *   inverse depth inliers: residual < min(residuals) + threshold  (:70; min skips NaN like MATLAB)
 *   cartesian inliers:     residual < threshold                   (:109) */
static void orc_q2r(const double *q, double R[3][3]) {
  const double r = q[0], x = q[1], y = q[2], z = q[3];
  R[0][0] = ((r * r + x * x) - y * y) - z * z;
  R[0][1] = 2.0 * (x * y - r * z);
  R[0][2] = 2.0 * (z * x + r * y);
  R[1][0] = 2.0 * (x * y + r * z);
  R[1][1] = ((r * r - x * x) + y * y) - z * z;
  R[1][2] = 2.0 * (y * z - r * x);
  R[2][0] = 2.0 * (z * x - r * y);
  R[2][1] = 2.0 * (y * z + r * x);
  R[2][2] = ((r * r - x * x) - y * y) + z * z;
}

static double orc_project_residual(const double Rwc[3][3], const double *v, const orc_cam *cam, const double *z) {
  double hc[3];
  for (int i = 0; i < 3; ++i) hc[i] = (Rwc[0][i] * v[0] + Rwc[1][i] * v[1]) + Rwc[2][i] * v[2];
  const double u = cam->f * (hc[0] / hc[2]) + cam->Cx;
  const double w = cam->f * (hc[1] / hc[2]) + cam->Cy;
  const double xu = (u - cam->Cx) / cam->f, yu = (w - cam->Cy) / cam->f;
  const double ru = sqrt(xu * xu + yu * yu);
  const double ru2 = ru * ru;
  const double D = (1.0 + cam->k1 * ru2) + cam->k2 * (ru2 * ru2);
  const double ud = (xu * D) * cam->f + cam->Cx, vd = (yu * D) * cam->f + cam->Cy;
  const double n0 = z[0] - ud, n1 = z[1] - vd;
  return sqrt(n0 * n0 + n1 * n1);
}

ORC_API int orc_ekf_support(const double *xi, const orc_cam *cam, const int32_t *idx_r, const int32_t *idx_ang,
                            const int32_t *idx_rho, const double *z_id, int n_id, const int32_t *idx_xyz,
                            const double *z_euc, int n_euc, double threshold, uint8_t *li_id, uint8_t *li_euc,
                            double *residuals_out) {
  int support = 0;
  double Rwc[3][3];
  orc_q2r(xi + 3, Rwc);
  if (n_id > 0) {
    double *res = (double *)malloc(sizeof(double) * (size_t)n_id);
    double mn = NAN;
    for (int j = 0; j < n_id; ++j) {
      double st, ct, sp, cp, v[3], mi[3];
      orc_sincos(xi[idx_ang[2 * j]], &st, &ct);
      orc_sincos(xi[idx_ang[2 * j + 1]], &sp, &cp);
      mi[0] = cp * st;
      mi[1] = -sp;
      mi[2] = cp * ct;
      const double rho = xi[idx_rho[j]];
      for (int a = 0; a < 3; ++a) v[a] = (xi[idx_r[3 * j + a]] - xi[a]) * rho + mi[a];
      res[j] = orc_project_residual(Rwc, v, cam, z_id + 2 * (size_t)j);
      if (!(res[j] != res[j]) && (mn != mn || res[j] < mn)) mn = res[j];
    }
    const double lim = mn + threshold;
    for (int j = 0; j < n_id; ++j) {
      const int in = res[j] < lim;
      if (li_id) li_id[j] = (uint8_t)in;
      support += in;
      if (residuals_out) residuals_out[j] = res[j];
    }
    free(res);
  }
  for (int j = 0; j < n_euc; ++j) {
    double v[3];
    for (int a = 0; a < 3; ++a) v[a] = xi[idx_xyz[3 * j + a]] - xi[a];
    const double r = orc_project_residual(Rwc, v, cam, z_euc + 2 * (size_t)j);
    const int in = r < threshold;
    if (li_euc) li_euc[j] = (uint8_t)in;
    support += in;
    if (residuals_out) residuals_out[n_id + j] = r;
  }
  return support;
}

/* ------------------------------------------------------------------------- */
/* n_hyp = ceil(log(1-p)/log(1-(1-epsilon))), epsilon = 1 - support/num_IC
 * (M/ransac_hypotheses.m:29,77-78).  MATLAB semantics at the edges: log(0) = -Inf gives
 * n_hyp = 0; a negative argument (support > num_IC: features with z that are not IC) gives a
 * complex quotient whose REAL part is what the relational test at :80 sees. */
ORC_API double orc_ekf_nhyp(int support, int num_ic) {
  const double p = 0.99;
  const double epsilon = 1.0 - ((double)support / (double)num_ic);
  const double a = 1.0 - (1.0 - epsilon);
  const double L = log(1.0 - p);
  if (a > 0.0) return ceil(L / log(a));
  if (a == 0.0) return 0.0;
  const double la = log(-a), pi = 3.14159265358979323846;
  return ceil((L * la) / (la * la + pi * pi));
}

/* Seeded stand-in for select_random_match (M/select_random_match.m:37-58): the first m entries
 * of a random permutation of the IC list, IN PERMUTATION ORDER (not sorted).
 * SPEC (shared with ekf.cu): draw d = 0..m-1 picks t_d uniform in [0, num_ic - d) from
 * splitmix64(seed ^ frame*A ^ hyp*B ^ (d+1)*C) (high 32 bits, multiply-shift), then maps it to
 * the t_d-th not yet chosen rank (earlier picks visited in ascending order). */
static uint64_t orc_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

ORC_API void orc_ekf_select(uint64_t seed, uint32_t frame, uint32_t hyp, int num_ic, int m, int32_t *rank_out) {
  int chosen[3];
  for (int d = 0; d < m; ++d) {
    const uint64_t x = orc_splitmix64(seed ^ ((uint64_t)frame * 0x9E3779B97F4A7C15ULL) ^
                                      ((uint64_t)hyp * 0xD1B54A32D192ED03ULL) ^
                                      ((uint64_t)(d + 1) * 0x8CB92BA72F3D8DD7ULL));
    const uint32_t r = (uint32_t)(x >> 32);
    int t = (int)(((uint64_t)r * (uint64_t)(num_ic - d)) >> 32);
    /* skip the ranks already taken, visiting them in ascending order */
    int sorted[3], ns = d;
    for (int i = 0; i < d; ++i) sorted[i] = chosen[i];
    for (int i = 1; i < ns; ++i)
      for (int j = i; j > 0 && sorted[j - 1] > sorted[j]; --j) {
        int tmp = sorted[j - 1];
        sorted[j - 1] = sorted[j];
        sorted[j] = tmp;
      }
    for (int i = 0; i < ns; ++i)
      if (t >= sorted[i]) ++t;
    chosen[d] = t;
    rank_out[d] = t;
  }
}

/* ------------------------------------------------------------------------- */
/* ransac_hypotheses (M/ransac_hypotheses.m:27-85) for one frame.
 *   sel: H x 3 supplied match selections (0-based FEATURE indices; only the first m of a row are
 *        used) or NULL for the seeded generator above (rank -> index into the IC list).
 *   n_hyp_init: 1000 (:35).  H: selections available (the loop also ends when they run out).
 * Loop semantics restated:
 *   for i = 1:n_hyp_init   -- the range is fixed when the loop starts
 *     if n_hyp == 0, break                                  (:41-46)
 *     m = 3 if num_IC > 3 else 1                            (select_random_match.m:47-51)
 *     ... update, support ...
 *     if support > max_support (strict): record masks, n_hyp := orc_ekf_nhyp   (:74-79)
 *     if n_hyp <= i: break -- where `i` is the INNER loop's variable (:57), i.e. m   (:80)
 * adaptive == 0 (not in the reference): both breaks are disabled, every selection is evaluated
 * and the first maximum wins -- the fixed-H mode of the GPU path.
 * Outputs: low_innovation_inlier per feature (only features with z are written, others keep
 * their input value, set_as_most_supported_hypothesis.m:32-53), stats. */
typedef struct {
  int32_t status;      /* 0 ok; 1 no individually compatible match (select_random_match errors) */
  int32_t n_evaluated; /* hypotheses evaluated */
  int32_t best_hyp;    /* 0-based index of the most supported hypothesis (-1: none had support) */
  int32_t max_support; /* StatData.RANSAC_HYP_SUPPORT (:85) */
  int32_t num_ic;
  int32_t m;
  double n_hyp;        /* StatData.RANSAC_ITER (:84): the final n_hyp */
} orc_ekf_result;

ORC_API int orc_ransac_hypotheses(const double *x, const double *P, int n, double std_z, const orc_cam *cam, int F,
                                  const int32_t *type, const int32_t *pos, const uint8_t *has_z, const uint8_t *ic,
                                  const double *z, const double *h, const double *Hcam, const double *Hfeat,
                                  const double *R, const int32_t *sel, int H, int n_hyp_init, int adaptive,
                                  uint64_t seed, uint32_t frame_id, uint8_t *li_inlier, orc_ekf_result *out,
                                  int32_t *supports) {
  /* generate_state_vector_pattern.m:29-51 as index lists, in feature order */
  int n_id = 0, n_euc = 0, num_ic = 0;
  int32_t *idx_r = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)(F + 1));
  int32_t *idx_ang = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(F + 1));
  int32_t *idx_rho = (int32_t *)malloc(sizeof(int32_t) * (size_t)(F + 1));
  int32_t *idx_xyz = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)(F + 1));
  int32_t *feat_id = (int32_t *)malloc(sizeof(int32_t) * (size_t)(F + 1));
  int32_t *feat_euc = (int32_t *)malloc(sizeof(int32_t) * (size_t)(F + 1));
  int32_t *ic_list = (int32_t *)malloc(sizeof(int32_t) * (size_t)(F + 1));
  double *z_id = (double *)malloc(sizeof(double) * 2 * (size_t)(F + 1));
  double *z_euc = (double *)malloc(sizeof(double) * 2 * (size_t)(F + 1));
  for (int i = 0; i < F; ++i) {
    if (ic[i]) ic_list[num_ic++] = i;
    if (!has_z[i]) continue;
    if (type[i] == 0) {
      for (int a = 0; a < 3; ++a) idx_r[3 * n_id + a] = pos[i] + a;
      idx_ang[2 * n_id] = pos[i] + 3;
      idx_ang[2 * n_id + 1] = pos[i] + 4;
      idx_rho[n_id] = pos[i] + 5;
      z_id[2 * n_id] = z[2 * (size_t)i];
      z_id[2 * n_id + 1] = z[2 * (size_t)i + 1];
      feat_id[n_id++] = i;
    } else {
      for (int a = 0; a < 3; ++a) idx_xyz[3 * n_euc + a] = pos[i] + a;
      z_euc[2 * n_euc] = z[2 * (size_t)i];
      z_euc[2 * n_euc + 1] = z[2 * (size_t)i + 1];
      feat_euc[n_euc++] = i;
    }
  }
  memset(out, 0, sizeof *out);
  out->best_hyp = -1;
  out->num_ic = num_ic;
  out->n_hyp = (double)n_hyp_init;
  int rc = 0;
  if (num_ic == 0) {
    out->status = 1;
    rc = 1;
  } else {
    const int m = num_ic > 3 ? 3 : 1;
    out->m = m;
    double *xi = (double *)malloc(sizeof(double) * (size_t)n);
    uint8_t *mi = (uint8_t *)malloc((size_t)(n_id + 1));
    uint8_t *me = (uint8_t *)malloc((size_t)(n_euc + 1));
    double n_hyp = (double)n_hyp_init;
    int max_support = 0;
    const int iters = n_hyp_init < H ? n_hyp_init : H;
    for (int i = 0; i < iters; ++i) {
      if (adaptive && n_hyp == 0.0) break;
      int32_t s[3];
      if (sel) {
        for (int a = 0; a < m; ++a) s[a] = sel[(size_t)i * 3 + a];
      } else {
        int32_t rk[3];
        orc_ekf_select(seed, frame_id, (uint32_t)i, num_ic, m, rk);
        for (int a = 0; a < m; ++a) s[a] = ic_list[rk[a]];
      }
      orc_ekf_update(x, P, n, pos, type, z, h, Hcam, Hfeat, R, s, m, xi);
      const int sup = orc_ekf_support(xi, cam, idx_r, idx_ang, idx_rho, z_id, n_id, idx_xyz, z_euc, n_euc, std_z, mi,
                                      me, NULL);
      if (supports) supports[i] = sup;
      out->n_evaluated = i + 1;
      if (sup > max_support) {
        max_support = sup;
        out->best_hyp = i;
        for (int j = 0; j < n_id; ++j) li_inlier[feat_id[j]] = mi[j];
        for (int j = 0; j < n_euc; ++j) li_inlier[feat_euc[j]] = me[j];
        n_hyp = orc_ekf_nhyp(sup, num_ic);
      }
      if (adaptive && n_hyp <= (double)m) break;
    }
    out->max_support = max_support;
    out->n_hyp = n_hyp;
    free(xi);
    free(mi);
    free(me);
  }
  free(idx_r);
  free(idx_ang);
  free(idx_rho);
  free(idx_xyz);
  free(feat_id);
  free(feat_euc);
  free(ic_list);
  free(z_id);
  free(z_euc);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* Re-prediction at x_k_k: M/@ekf_filter/rescue_hi_inliers.m:32-33
 *   features_info = predict_camera_measurements(filter.x_k_k, cam, features_info)   (M/predict_camera_measurements.m:27-68)
 *   features_info = calculate_derivatives(filter.x_k_k, cam, features_info)         (M/calculate_derivatives.m:27-59)
 * predict: hi_inverse_depth.m:27-86 / hi_cartesian.m:27-80 -- hrl = rotcw * v, the +-60 degree field-of-view test on
 *   atan2(hrl(1), hrl(3)) and atan2(hrl(2), hrl(3)) (:60-66), hu_my_version.m:32-41, distort_fm_my_version.m:52-61, the
 *   image-bounds test (:79); a feature that fails a test KEEPS its previous h (predict_camera_measurements.m:37-39,:62-64).
 * derivatives, for every feature whose h is not empty (calculate_derivatives.m:34), zi = that h:
 *   calculate_Hi_inverse_depth_my_version.m:27-190 / calculate_Hi_cartesian_my_version.m:27-168
 *     Rrw = inv(q2r(q))                          dhu_dhrl = [f/hz 0 -hx f/hz^2; 0 f/hz -hy f/hz^2]
 *     dhd_dhu = inv(inv(jacob_distor(cam, zi)))   (jacob_undistor_fm_my_version.m:38 inverts once, dhd_dhu again)
 *     jacob_distor_fm_my_version.m:38-60          dh_dhrl = dhd_dhu * dhu_dhrl
 *     H(:,1:3) = dh_dhrl * (-Rrw * rho)  [cartesian: -Rrw]
 *     H(:,4:7) = dh_dhrl * dRq_times_a_by_dq(qconj(q), a) * diag([1 -1 -1 -1])   (dRq_times_a_by_dq.m:26-103)
 *     H(:,8:13) = 0
 *     feature block = dh_dhrl * [rho*Rrw, Rrw*dm/dtheta, Rrw*dm/dphi, Rrw*(y - rw)]   [cartesian: dh_dhrl * Rrw]
 * Specification shared with 3pre_b200/csrc/ekf.cu (k_ekf_predict): inv() = orc_inv (Gauss-Jordan above), sin / cos =
 * orc_sincos, every product  C(i,j) = sum_k A(i,k) B(k,j)  accumulated left to right in k starting from the k = 0
 * term.  MATLAB's inv / mtimes (LAPACK / BLAS, possibly FMA) differ from this by rounding: checked against an
 * independent numpy restatement at 1e-9 (tests/test_oracle_ekf_update_cpu.py). */
static void mm(const double *A, int ra, int ca, const double *B, int cb, double *C) { /* row-major */
  for (int i = 0; i < ra; ++i)
    for (int j = 0; j < cb; ++j) {
      double s = A[i * ca] * B[j];
      for (int k = 1; k < ca; ++k) s = s + A[i * ca + k] * B[k * cb + j];
      C[i * cb + j] = s;
    }
}

ORC_API void orc_ekf_predict(const double *x, int n, const orc_cam *cam, int nRows, int nCols, int F,
                             const int32_t *type, const int32_t *pos, const uint8_t *has_h_in, const double *h_in,
                             double *h_out, uint8_t *has_h_out, uint8_t *predicted, double *Hcam, double *Hfeat) {
  (void)n;
  double Rwc[3][3], Rf[9], Rrw[9];
  orc_q2r(x + 3, Rwc);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rf[i * 3 + j] = Rwc[i][j];
  orc_inv(Rf, 3, Rrw); /* inv(q2r(q)) of the Jacobians */
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < F; ++i) {
    const double *y = x + pos[i];
    double a[3], mi[3] = {0, 0, 0}, st = 0, ct = 0, sp = 0, cp = 0, rho = 1.0;
    if (type[i] == 0) {
      orc_sincos(y[3], &st, &ct);
      orc_sincos(y[4], &sp, &cp);
      mi[0] = cp * st, mi[1] = -sp, mi[2] = cp * ct; /* m.m:32-34 */
      rho = y[5];
      for (int k = 0; k < 3; ++k) a[k] = (y[k] - x[k]) * rho + mi[k];
    } else {
      for (int k = 0; k < 3; ++k) a[k] = y[k] - x[k];
    }
    /* ---- predict_camera_measurements: r_cw = r_wc' (hi_inverse_depth.m:33) / inv(r_wc) (hi_cartesian.m:33) ---- */
    double hrl[3];
    if (type[i] == 0) {
      for (int k = 0; k < 3; ++k) hrl[k] = (Rwc[0][k] * a[0] + Rwc[1][k] * a[1]) + Rwc[2][k] * a[2];
    } else {
      mm(Rrw, 3, 3, a, 1, hrl);
    }
    int ok = 1;
    const double ax = atan2(hrl[0], hrl[2]) * 180 / PI, ay = atan2(hrl[1], hrl[2]) * 180 / PI;
    if (ax < -60 || ax > 60 || ay < -60 || ay > 60) ok = 0;
    double ud = 0, vd = 0;
    if (ok) {
      const double u = cam->Cx + (hrl[0] / hrl[2]) * cam->f, v = cam->Cy + (hrl[1] / hrl[2]) * cam->f;
      const double xu = (u - cam->Cx) / cam->f, yu = (v - cam->Cy) / cam->f;
      const double ru = sqrt(xu * xu + yu * yu);
      const double ru2 = ru * ru;
      const double D = (1.0 + cam->k1 * ru2) + cam->k2 * (ru2 * ru2);
      ud = (xu * D) * cam->f + cam->Cx, vd = (yu * D) * cam->f + cam->Cy;
      if (!(ud > 0 && ud < nCols && vd > 0 && vd < nRows)) ok = 0;
    }
    predicted[i] = (uint8_t)ok;
    has_h_out[i] = (uint8_t)(ok || has_h_in[i]);
    h_out[2 * i] = ok ? ud : h_in[2 * i];
    h_out[2 * i + 1] = ok ? vd : h_in[2 * i + 1];
    for (int k = 0; k < 26; ++k) Hcam[26 * (size_t)i + k] = 0.0;
    for (int k = 0; k < 12; ++k) Hfeat[12 * (size_t)i + k] = 0.0;
    if (!has_h_out[i]) continue; /* calculate_derivatives.m:34 */
    /* ---- calculate_derivatives ---- */
    double hc[3];
    mm(Rrw, 3, 3, a, 1, hc);
    const double f = cam->f;
    const double dhu[6] = {f / hc[2], 0.0, -hc[0] * f / (hc[2] * hc[2]), 0.0, f / hc[2], -hc[1] * f / (hc[2] * hc[2])};
    const double uu = h_out[2 * i], vv = h_out[2 * i + 1];
    const double xd = uu - cam->Cx, yd = vv - cam->Cy;
    const double r2 = (xd * xd + yd * yd) / (f * f), r4 = r2 * r2;
    const double k1 = cam->k1, k2 = cam->k2;
    double J[4], Ji[4], Jii[4];
    J[0] = (1 + k1 * r2 + k2 * r4) + (uu - cam->Cx) * (k1 + 2 * k2 * r2) * (2 * (uu - cam->Cx) / (f * f));
    J[3] = (1 + k1 * r2 + k2 * r4) + (vv - cam->Cy) * (k1 + 2 * k2 * r2) * (2 * (vv - cam->Cy) / (f * f));
    J[1] = (uu - cam->Cx) * (k1 + 2 * k2 * r2) * (2 * (vv - cam->Cy) / (f * f));
    J[2] = (vv - cam->Cy) * (k1 + 2 * k2 * r2) * (2 * (uu - cam->Cx) / (f * f));
    orc_inv(J, 2, Ji);
    orc_inv(Ji, 2, Jii);
    double dh[6];
    mm(Jii, 2, 2, dhu, 3, dh); /* dh_dhrl, 2 x 3 */
    double drw[9], Hrw[6];
    for (int k = 0; k < 9; ++k) drw[k] = type[i] == 0 ? -Rrw[k] * rho : -Rrw[k];
    mm(dh, 2, 3, drw, 3, Hrw);
    /* dRq_times_a_by_dq(qconj(q), a) * diag([1 -1 -1 -1]) */
    const double q0 = x[3], qx = -x[4], qy = -x[5], qz = -x[6];
    const double dR[4][9] = {{2 * q0, -2 * qz, 2 * qy, 2 * qz, 2 * q0, -2 * qx, -2 * qy, 2 * qx, 2 * q0},
                             {2 * qx, 2 * qy, 2 * qz, 2 * qy, -2 * qx, -2 * q0, 2 * qz, 2 * q0, -2 * qx},
                             {-2 * qy, 2 * qx, 2 * q0, 2 * qx, 2 * qy, 2 * qz, -2 * q0, 2 * qz, -2 * qy},
                             {-2 * qz, -2 * q0, 2 * qx, 2 * q0, -2 * qz, 2 * qy, 2 * qx, 2 * qy, 2 * qz}};
    double dq[12]; /* 3 x 4 row-major */
    for (int c = 0; c < 4; ++c) {
      double t[3];
      mm(dR[c], 3, 3, a, 1, t);
      for (int r = 0; r < 3; ++r) dq[r * 4 + c] = c == 0 ? t[r] : t[r] * -1.0;
    }
    double Hq[8];
    mm(dh, 2, 3, dq, 4, Hq);
    for (int r = 0; r < 2; ++r) {
      for (int c = 0; c < 3; ++c) Hcam[26 * (size_t)i + 2 * c + r] = Hrw[r * 3 + c];
      for (int c = 0; c < 4; ++c) Hcam[26 * (size_t)i + 2 * (3 + c) + r] = Hq[r * 4 + c];
    }
    if (type[i] == 0) {
      double dy[18]; /* 3 x 6 row-major */
      const double dth[3] = {cp * ct, 0.0, -cp * st}, dph[3] = {-sp * st, -cp, -sp * ct};
      double yr[3] = {y[0] - x[0], y[1] - x[1], y[2] - x[2]}, c4[3], c5[3], c6[3];
      mm(Rrw, 3, 3, dth, 1, c4);
      mm(Rrw, 3, 3, dph, 1, c5);
      mm(Rrw, 3, 3, yr, 1, c6);
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) dy[r * 6 + c] = rho * Rrw[r * 3 + c];
        dy[r * 6 + 3] = c4[r], dy[r * 6 + 4] = c5[r], dy[r * 6 + 5] = c6[r];
      }
      double Hy[12];
      mm(dh, 2, 3, dy, 6, Hy);
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 6; ++c) Hfeat[12 * (size_t)i + 2 * c + r] = Hy[r * 6 + c];
    } else {
      double Hy[6];
      mm(dh, 2, 3, Rrw, 3, Hy);
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) Hfeat[12 * (size_t)i + 2 * c + r] = Hy[r * 3 + c];
    }
  }
}
