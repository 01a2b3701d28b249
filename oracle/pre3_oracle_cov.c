/* pre3_oracle_cov.c -- CPU restatement of the covariance of the RANSAC pose.
 *
 * TEST INFRASTRUCTURE ONLY (see pre3_oracle.c).  PARITY UNPINNED: the reference holds no vectors for this step and its
 * call site is commented out (M/mex_files/RANSAC_CALCULATION/RANSAC_CALC_VER2.m:204-206); checked against an independent
 * numpy restatement (tests) and against the analytic derivatives of the same error function.
 *
 * Follows  M/cov_est_RANSAC_deriv.m:1-244 (the variant that runs: central differences through M/deriv.m:2-9 with
 * eps_xy = 1e-6, :26), with cart2sph / sph2cart as MATLAB defines them, q2R of
 * M/slamToolbox_11_02_18/FrameTransforms/Rotations/q2R.m:19-35 and R2q of .../R2q.m:11-55 (orc_R2q).
 * M/covariance_estimate_RANSAC.m is the same computation through the third-party `derivest` (absent from the tree) and
 * refers to undefined variables (qq4 :117, T_ab__1_ :122, k :180): it cannot run and is not restated.
 *
 * x = [T1 T2 T3 q1 q2 q3 q4] (:72-93), noise z = [b_theta b_phi b_r a_theta a_phi a_r] (:117-145).  Per support point:
 *   E      :150    gradEk :153 (7 central differences of E)
 *   d2Ek_dx2 :156  column j = deriv(gradEk, x_j)          (nested central differences, 4 E evaluations per entry)
 *   d2Ek_dxde_k :160-175  = deriv(gradEk, z_k)
 * summed over the points (:150-192; the six noise derivatives are summed over the points as well, :184-189), then
 *   dA_dz = G2tot \ [d1 .. d6] :200,  sm_cov_censi = dA_dz * blkdiag(R, R) * dA_dz' :215,  s2 = Etot / (k - 3) :217. */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

void orc_R2q(const double *R, double *q);

/* param vector: 0 b_theta 1 b_phi 2 b_r 3 a_theta 4 a_phi 5 a_r 6..9 quat 10..12 T */
static double cov_E(const double *p) { /* E_k_func, cov_est_RANSAC_deriv.m:236-241 */
  /* sph2cart: z = r sin(elev); rcoselev = r cos(elev); x = rcoselev cos(az); y = rcoselev sin(az) */
  const double za = p[5] * sin(p[4]), rca = p[5] * cos(p[4]);
  const double xa = rca * cos(p[3]), ya = rca * sin(p[3]);
  const double zb = p[2] * sin(p[1]), rcb = p[2] * cos(p[1]);
  const double xb = rcb * cos(p[0]), yb = rcb * sin(p[0]);
  const double a = p[6], b = p[7], c = p[8], d = p[9];
  const double aa = a * a, ab = 2 * a * b, ac = 2 * a * c, ad = 2 * a * d, bb = b * b, bc = 2 * b * c, bd = 2 * b * d,
               cc = c * c, cd = 2 * c * d, dd = d * d;
  const double r11 = aa + bb - cc - dd, r12 = bc - ad, r13 = bd + ac;
  const double r21 = bc + ad, r22 = aa - bb + cc - dd, r23 = cd - ab;
  const double r31 = bd - ac, r32 = cd + ab, r33 = aa - bb - cc + dd;
  const double v1 = xa - (((r11 * xb + r12 * yb) + r13 * zb) + p[10]);
  const double v2 = ya - (((r21 * xb + r22 * yb) + r23 * zb) + p[11]);
  const double v3 = za - (((r31 * xb + r32 * yb) + r33 * zb) + p[12]);
  const double nrm = sqrt((v1 * v1 + v2 * v2) + v3 * v3);
  return nrm * nrm;
}

static const int XIDX[7] = {10, 11, 12, 6, 7, 8, 9};

static void cov_grad(const double *p, double eps, double *g) { /* gradEk :72-93, deriv.m:7-9 */
  double w[13];
  memcpy(w, p, sizeof(w));
  for (int i = 0; i < 7; ++i) {
    const int k = XIDX[i];
    w[k] = p[k] + eps / 2;
    const double f1 = cov_E(w);
    w[k] = p[k] - eps / 2;
    const double f0 = cov_E(w);
    w[k] = p[k];
    g[i] = (f1 - f0) / eps;
  }
}

static void cov_dgrad(const double *p, int k, double eps, double *col) { /* deriv(@(t) gradEk(... t ...), p_k, eps) */
  double w[13], g1[7], g0[7];
  memcpy(w, p, sizeof(w));
  w[k] = p[k] + eps / 2;
  cov_grad(w, eps, g1);
  w[k] = p[k] - eps / 2;
  cov_grad(w, eps, g0);
  for (int i = 0; i < 7; ++i) col[i] = (g1[i] - g0[i]) / eps;
}

/* A (7 x 7, column-major, destroyed) \ B (7 x nb, column-major, overwritten): LU with partial pivoting.
 * Returns 0, or 1 when a pivot is exactly zero. */
static int solve7(double *A, double *B, int nb) {
  int sing = 0;
  for (int k = 0; k < 7; ++k) {
    int piv = k;
    double best = fabs(A[k + 7 * k]);
    for (int r = k + 1; r < 7; ++r)
      if (fabs(A[r + 7 * k]) > best) best = fabs(A[r + 7 * k]), piv = r;
    if (piv != k) {
      for (int c = 0; c < 7; ++c) {
        const double t = A[k + 7 * c];
        A[k + 7 * c] = A[piv + 7 * c], A[piv + 7 * c] = t;
      }
      for (int c = 0; c < nb; ++c) {
        const double t = B[k + 7 * c];
        B[k + 7 * c] = B[piv + 7 * c], B[piv + 7 * c] = t;
      }
    }
    if (A[k + 7 * k] == 0.0) sing = 1;
    for (int r = k + 1; r < 7; ++r) {
      const double f = A[r + 7 * k] / A[k + 7 * k];
      for (int c = k + 1; c < 7; ++c) A[r + 7 * c] = A[r + 7 * c] - f * A[k + 7 * c];
      for (int c = 0; c < nb; ++c) B[r + 7 * c] = B[r + 7 * c] - f * B[k + 7 * c];
    }
  }
  for (int c = 0; c < nb; ++c)
    for (int r = 6; r >= 0; --r) {
      double s = B[r + 7 * c];
      for (int j = r + 1; j < 7; ++j) s = s - A[r + 7 * j] * B[j + 7 * c];
      B[r + 7 * c] = s / A[r + 7 * r];
    }
  return sing;
}

/* Ya, Yb: n x 3 (point-major); R row-major, Ya ~ R Yb + T.  Outputs column-major: cov 7x7, G2tot 7x7, Gtot 7,
 * dA_dz 7x6; scal = {Etot, s2}.  Returns 0, 1 (singular G2tot). */
ORC_API int orc_cov_est_ransac_deriv(const double *Ya, const double *Yb, int n, const double *R, const double *T,
                                     double *cov, double *G2tot, double *Gtot, double *dA_dz, double *scal) {
  const double eps = 0.000001; /* :26 */
  double q[4], Etot = 0.0, D[42];
  orc_R2q(R, q); /* :14 */
  memset(G2tot, 0, 49 * sizeof(double));
  memset(Gtot, 0, 7 * sizeof(double));
  memset(D, 0, sizeof(D));
  for (int i = 0; i < n; ++i) {
    double p[13], g[7], col[7];
    const double *pb = Yb + 3 * i, *pa = Ya + 3 * i;
    /* cart2sph: az = atan2(y, x); elev = atan2(z, hypot(x, y)); r = hypot(hypot(x, y), z)   (:50-51) */
    p[0] = atan2(pb[1], pb[0]), p[1] = atan2(pb[2], hypot(pb[0], pb[1])), p[2] = hypot(hypot(pb[0], pb[1]), pb[2]);
    p[3] = atan2(pa[1], pa[0]), p[4] = atan2(pa[2], hypot(pa[0], pa[1])), p[5] = hypot(hypot(pa[0], pa[1]), pa[2]);
    p[6] = q[0], p[7] = q[1], p[8] = q[2], p[9] = q[3];
    p[10] = T[0], p[11] = T[1], p[12] = T[2];
    Etot = Etot + cov_E(p); /* :150 */
    cov_grad(p, eps, g);    /* :153 */
    for (int r = 0; r < 7; ++r) Gtot[r] = Gtot[r] + g[r];
    for (int j = 0; j < 7; ++j) { /* :156 */
      cov_dgrad(p, XIDX[j], eps, col);
      for (int r = 0; r < 7; ++r) G2tot[r + 7 * j] = G2tot[r + 7 * j] + col[r];
    }
    for (int k = 0; k < 6; ++k) { /* :160-189 */
      cov_dgrad(p, k, eps, col);
      for (int r = 0; r < 7; ++r) D[r + 7 * k] = D[r + 7 * k] + col[r];
    }
  }
  double A[49];
  memcpy(A, G2tot, sizeof(A));
  memcpy(dA_dz, D, sizeof(D));
  const int sing = solve7(A, dA_dz, 6); /* :200 */
  const double pi = 3.14159265358979323846;
  const double sg[3] = {0.02 * pi / 180, 0.02 * pi / 180, 0.015}; /* :209 */
  double s2v[6];
  for (int k = 0; k < 6; ++k) s2v[k] = sg[k % 3] * sg[k % 3]; /* R = diag(sigma.^2), blkdiag(R, R) :213-215 */
  for (int c = 0; c < 7; ++c)
    for (int r = 0; r < 7; ++r) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s = s + (dA_dz[r + 7 * k] * s2v[k]) * dA_dz[c + 7 * k];
      cov[r + 7 * c] = s;
    }
  scal[0] = Etot;
  scal[1] = Etot / (double)(n - 3); /* :216-217 */
  return sing;
}
