/* pre3_oracle.c -- CPU restatement of the 3PRE frame-to-frame hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs, never by the product path
 * (3pre_b200/ fails loudly when its CUDA library is missing).
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - orc_siftmatch_*  : PINNED.  Checked bit-for-bit against the reference's own
 *     matlab_code/sift/siftmatch.c compiled from /root/reference into
 *     oracle/_ref/ (oracle/Makefile) and against the committed golden fixtures.
 *   - everything else  : "parity unpinned".  The reference (MATLAB R2011a + MKL
 *     svd/eig/rand) cannot be executed here and ships no golden vectors for
 *     RANSAC / Horn / support scoring.  The restatement follows
 *     the cited .m files statement by statement; svd/eig are replaced by the
 *     fixed-order Jacobi iterations specified below (the CUDA kernels implement
 *     the same operation order, so masks / counts / selection are comparable
 *     bit-for-bit); an independent LAPACK restatement (oracle/ref_numpy.py)
 *     cross-checks R, T to 1e-9.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 * No FMA contraction, IEEE double/float throughout, only + - * / sqrt on the
 * RANSAC path.  `M/` below = /root/reference/matlab_code/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Stage 1: siftmatch  (M/sift/siftmatch.c:83-132)                            */
/* ------------------------------------------------------------------------- */
/* Output: pairs[2*i]=k1 (0-based), pairs[2*i+1]=k2, score[i]=best; returns the
 * number of accepted pairs.  Accumulation type follows PROMOTE_* (:61-64), the
 * start values MAXVAL_* (:66-69), the update rule :110-116, the float-cast
 * ratio test :122-123 (thresh narrowed to float at the call, :87,:205). */
#define ORC_MATCH(NAME, T, ACC, MAXVAL)                                        \
  ORC_API int NAME(const T *L1, const T *L2, int K1, int K2, int ND,          \
                   double thresh_d, int32_t *pairs, double *score) {           \
    const float thresh = (float)thresh_d;                                      \
    int n = 0;                                                                 \
    for (int k1 = 0; k1 < K1; ++k1) {                                          \
      const T *a = L1 + (size_t)k1 * ND;                                       \
      ACC best = (MAXVAL), second_best = (MAXVAL);                             \
      int bestk = -1;                                                          \
      for (int k2 = 0; k2 < K2; ++k2) {                                        \
        const T *b = L2 + (size_t)k2 * ND;                                     \
        ACC acc = 0;                                                           \
        for (int bin = 0; bin < ND; ++bin) {                                   \
          ACC delta = ((ACC)a[bin]) - ((ACC)b[bin]);                           \
          acc += delta * delta;                                                \
        }                                                                      \
        if (acc < best) {                                                      \
          second_best = best;                                                  \
          best = acc;                                                          \
          bestk = k2;                                                          \
        } else if (acc < second_best) {                                        \
          second_best = acc;                                                   \
        }                                                                      \
      }                                                                        \
      if (thresh * (float)best <= (float)second_best && bestk != -1) {         \
        pairs[2 * n] = k1;                                                     \
        pairs[2 * n + 1] = bestk;                                              \
        score[n] = (double)best;                                               \
        ++n;                                                                   \
      }                                                                        \
    }                                                                          \
    return n;                                                                  \
  }

ORC_MATCH(orc_siftmatch_f64, double, double, INFINITY)
ORC_MATCH(orc_siftmatch_f32, float, float, ((float)INFINITY))
ORC_MATCH(orc_siftmatch_i8, signed char, int, 0x7fffffff)
ORC_MATCH(orc_siftmatch_u8, unsigned char, int, 0x7fffffff)

/* ------------------------------------------------------------------------- */
/* 3x3 one-sided Jacobi SVD (stands in for MATLAB svd, find_transform_matrix.m:17; used for H that is not clearly
 * well-conditioned, see orc_polar_fast below)
 *
 * SPEC (shared with 3pre_b200/csrc/fit.cuh, which implements the same order):
 *   A := H (row-major a[r][c]),  V := I.
 *   repeat up to 15 sweeps over the column pairs (0,1),(0,2),(1,2):
 *     alpha = (a0p*a0p + a1p*a1p) + a2p*a2p, beta likewise for q,
 *     gamma = (a0p*a0q + a1p*a1q) + a2p*a2q
 *     skip the pair if gamma == 0 or gamma*gamma <= 1e-28*(alpha*beta)
 *     d = beta - alpha;  g2 = 2*gamma
 *     t = g2 / (|d| + sqrt(d*d + g2*g2));  if d < 0: t = -t
 *     c = 1/sqrt(1 + t*t);  s = c*t
 *     for each row r of A and of V: (xp, xq) := (c*xp - s*xq, s*xp + c*xq)
 *   stop when a full sweep skipped every pair.
 *   sigma_j = sqrt((a0j*a0j + a1j*a1j) + a2j*a2j);   columns are NOT sorted.
 */
static void orc_svd3_cols(double a[3][3], double v[3][3], double sig[3]) {
  static const int PP[3] = {0, 0, 1}, QQ[3] = {1, 2, 2};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) v[r][c] = (r == c) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 15; ++sweep) {
    int rotated = 0;
    for (int k = 0; k < 3; ++k) {
      const int p = PP[k], q = QQ[k];
      double alpha = (a[0][p] * a[0][p] + a[1][p] * a[1][p]) + a[2][p] * a[2][p];
      double beta = (a[0][q] * a[0][q] + a[1][q] * a[1][q]) + a[2][q] * a[2][q];
      double gamma = (a[0][p] * a[0][q] + a[1][p] * a[1][q]) + a[2][p] * a[2][q];
      if (gamma == 0.0) continue;
      if (gamma * gamma <= 1e-28 * (alpha * beta)) continue;
      rotated = 1;
      double d = beta - alpha;
      double g2 = 2.0 * gamma;
      double t = g2 / (fabs(d) + sqrt(d * d + g2 * g2));
      if (d < 0.0) t = -t;
      double c = 1.0 / sqrt(1.0 + t * t);
      double s = c * t;
      for (int r = 0; r < 3; ++r) {
        double xp = a[r][p], xq = a[r][q];
        a[r][p] = c * xp - s * xq;
        a[r][q] = s * xp + c * xq;
      }
      for (int r = 0; r < 3; ++r) {
        double xp = v[r][p], xq = v[r][q];
        v[r][p] = c * xp - s * xq;
        v[r][q] = s * xp + c * xq;
      }
    }
    if (!rotated) break;
  }
  for (int j = 0; j < 3; ++j)
    sig[j] = sqrt((a[0][j] * a[0][j] + a[1][j] * a[1][j]) + a[2][j] * a[2][j]);
}

static double orc_det3(double m[3][3]) {
  return (m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
          m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])) +
         m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

/* V*U' of a WELL-CONDITIONED H without the SVD: Xq = V*U' is the transpose of the orthogonal polar factor of H
 * (H = U S V' = (U V')(V S V')), which Newton's iteration X <- (X + X^-T)/2 reaches quadratically.
 *
 * SPEC (shared with 3pre_b200/csrc/fit.cuh: same operations in the same order; + - * / sqrt only):
 *   nf2  = sum of the squares of the entries of H, row-major order, accumulated from 0.0
 *   detH = orc_det3(H)
 *   the fast path applies iff  nf2 > 0,  |detH| > 50*threshold*nf2   (then sigma_min >= 2|detH|/nf2 > 100*threshold: no
 *   singular value anywhere near the reference's test `S(i,i) < threshold`, find_transform_matrix.m:20,:26)  and
 *   |detH| > 1e-9*(nf2*sqrt(nf2))  (condition number below ~3e4); otherwise the Jacobi SVD above is used.
 *     detH < 0:  det(V*U') = -1 with no small singular value  ->  state -1 (:25-36), nothing else to compute.
 *     detH > 0:  X = H*(1/sqrt(nf2));  repeat at most 40 times:
 *                  C = cofactors of X (C00 = x11*x22 - x12*x21, C01 = x12*x20 - x10*x22, C02 = x10*x21 - x11*x20,
 *                      C10 = x02*x21 - x01*x22, C11 = x00*x22 - x02*x20, C12 = x01*x20 - x00*x21,
 *                      C20 = x01*x12 - x02*x11, C21 = x02*x10 - x00*x12, C22 = x00*x11 - x01*x10)
 *                  dt = (x00*C00 + x01*C01) + x02*C02
 *                  iterations 0..2 are SCALED (Higham: gamma = (||X^-T||_F / ||X||_F)^(1/2); 7 iterations then suffice
 *                  for every admitted H, whatever its conditioning):
 *                      a = sum C[r][c]^2, b = sum X[r][c]^2 (row-major, from 0.0)
 *                      g = sqrt(sqrt(a / ((dt*dt)*b)));  hx = 0.5*g;  hc = 0.5/(dt*g)
 *                  later iterations: hx = 0.5;  hc = 0.5/dt
 *                  Y[r][c] = hx*X[r][c] + C[r][c]*hc;   e = Y[r][c] - X[r][c], diff2 += e*e (row-major, from 0.0)
 *                  X = Y;  stop when diff2 <= 1e-28
 *                Xq = X'  ->  state 1.
 * Returns 1 (Xq filled), -1 (reflection), 0 (not applicable: use the SVD). */
static int orc_polar_fast(double H[3][3], double threshold, double Xq[3][3]) {
  double nf2 = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) nf2 = nf2 + H[r][c] * H[r][c];
  const double detH = orc_det3(H);
  if (!(nf2 > 0.0)) return 0;
  if (!(fabs(detH) > 50.0 * threshold * nf2)) return 0;
  const double rn = sqrt(nf2);
  if (!(fabs(detH) > 0.000000001 * (nf2 * rn))) return 0;
  if (detH < 0.0) return -1;
  const double sc = 1.0 / rn;
  double X[3][3], Y[3][3], C[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) X[r][c] = H[r][c] * sc;
  for (int it = 0; it < 40; ++it) {
    C[0][0] = X[1][1] * X[2][2] - X[1][2] * X[2][1];
    C[0][1] = X[1][2] * X[2][0] - X[1][0] * X[2][2];
    C[0][2] = X[1][0] * X[2][1] - X[1][1] * X[2][0];
    C[1][0] = X[0][2] * X[2][1] - X[0][1] * X[2][2];
    C[1][1] = X[0][0] * X[2][2] - X[0][2] * X[2][0];
    C[1][2] = X[0][1] * X[2][0] - X[0][0] * X[2][1];
    C[2][0] = X[0][1] * X[1][2] - X[0][2] * X[1][1];
    C[2][1] = X[0][2] * X[1][0] - X[0][0] * X[1][2];
    C[2][2] = X[0][0] * X[1][1] - X[0][1] * X[1][0];
    const double dt = (X[0][0] * C[0][0] + X[0][1] * C[0][1]) + X[0][2] * C[0][2];
    double hx = 0.5, hc;
    if (it < 3) {
      double a = 0.0, b = 0.0;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          a = a + C[r][c] * C[r][c];
          b = b + X[r][c] * X[r][c];
        }
      const double g = sqrt(sqrt(a / ((dt * dt) * b)));
      hx = 0.5 * g;
      hc = 0.5 / (dt * g);
    } else {
      hc = 0.5 / dt;
    }
    double diff2 = 0.0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        Y[r][c] = hx * X[r][c] + C[r][c] * hc;
        const double e = Y[r][c] - X[r][c];
        diff2 = diff2 + e * e;
      }
    memcpy(X, Y, sizeof X);
    if (diff2 <= 1e-28) break;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Xq[r][c] = X[c][r];
  return 1;
}

/* find_transform_matrix  (M/mex_files/RANSAC_CALCULATION/find_transform_matrix.m:2-42)
 * pset1, pset2: 3 x pnum column-major.  rot row-major 3x3, trans[3].
 * Returns state in {1, 2, 0, -1}.  Model: pset1 ~ rot*pset2 + trans.
 *
 * `idx` (may be NULL) gathers columns: point i is column idx[i] (0-based) of the
 * arrays -- this is Ya(:,RandIndex) / Ya(:,position_inliers) of the callers.
 *
 * Conventions where MATLAB's svd output is not determined by the mathematics
 * (null directions of a rank-deficient H; SURVEY.md 7 "hard parts"):
 *   - exactly one sigma_j < 1e-11 (:20,:26): u_j := u_{j+1} x u_{j+2} (cyclic), so
 *     det(U)=+1, det(V U')=+1 and the state is 1.  MATLAB gets state 1 or 2
 *     depending on LAPACK's arbitrary sign, with the same rot after its column
 *     flip (:29-30).
 *   - two or three sigma_j < 1e-11: state -1 (MATLAB: 1 or -1 by the same luck).
 *   - any non-finite entry: state 0 (round(mdet) is neither 1 nor -1, :38-42).
 */
ORC_API int orc_find_transform_thr(const double *pset1, const double *pset2,
                                   const int32_t *idx, int pnum, double threshold,
                                   double *rot, double *trans);

ORC_API int orc_find_transform(const double *pset1, const double *pset2,
                               const int32_t *idx, int pnum, double *rot,
                               double *trans) {
  return orc_find_transform_thr(pset1, pset2, idx, pnum, 0.00000000001 /* :20 */, rot, trans);
}

/* threshold: 1e-11 in find_transform_matrix.m:20, 1e-14 in
 * M/code_from_dr_ye/find_transform_matrix_dr_ye.m:19 (the only difference between the two files) */
ORC_API int orc_find_transform_thr(const double *pset1, const double *pset2,
                                   const int32_t *idx, int pnum, double threshold,
                                   double *rot, double *trans) {
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double ct1[3] = {0, 0, 0}, ct2[3] = {0, 0, 0};
  for (int i = 0; i < pnum; ++i) { /* sum(pset,2) :11 */
    const double *p1 = pset1 + 3 * (size_t)(idx ? idx[i] : i);
    const double *p2 = pset2 + 3 * (size_t)(idx ? idx[i] : i);
    for (int r = 0; r < 3; ++r) {
      ct1[r] += p1[r];
      ct2[r] += p2[r];
    }
  }
  for (int r = 0; r < 3; ++r) {
    ct1[r] = ct1[r] / (double)pnum;
    ct2[r] = ct2[r] / (double)pnum;
  }
  for (int i = 0; i < pnum; ++i) { /* :12-15  H = H + q2*q1' */
    const double *p1 = pset1 + 3 * (size_t)(idx ? idx[i] : i);
    const double *p2 = pset2 + 3 * (size_t)(idx ? idx[i] : i);
    double q1[3], q2[3];
    for (int r = 0; r < 3; ++r) {
      q1[r] = p1[r] - ct1[r];
      q2[r] = p2[r] - ct2[r];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) H[r][c] = H[r][c] + q2[r] * q1[c];
  }
  int finite = 1;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      if (!isfinite(H[r][c])) finite = 0;
  int state;
  double U[3][3], V[3][3], sig[3], A[3][3], Xq[3][3];
  memcpy(A, H, sizeof A);
  int nsmall = 0, jsmall = -1;
  const int fast = finite ? orc_polar_fast(H, threshold, Xq) : 0;
  if (finite && fast == 0) {
    orc_svd3_cols(A, V, sig);
    for (int j = 0; j < 3; ++j)
      if (sig[j] < threshold) {
        ++nsmall;
        jsmall = j;
      }
  }
  if (!finite) {
    state = 0;
  } else if (fast != 0) {
    state = fast; /* 1: Xq = V*U' from the polar iteration; -1: reflection, no singular value near the threshold */
  } else if (nsmall >= 2) {
    state = -1;
  } else {
    for (int j = 0; j < 3; ++j)
      if (j != jsmall) { /* u_j = a_j * (1/sigma_j) */
        double inv = 1.0 / sig[j];
        for (int r = 0; r < 3; ++r) U[r][j] = A[r][j] * inv;
      }
    if (nsmall == 1) {
      const int j1 = (jsmall + 1) % 3, j2 = (jsmall + 2) % 3;
      U[0][jsmall] = U[1][j1] * U[2][j2] - U[2][j1] * U[1][j2];
      U[1][jsmall] = U[2][j1] * U[0][j2] - U[0][j1] * U[2][j2];
      U[2][jsmall] = U[0][j1] * U[1][j2] - U[1][j1] * U[0][j2];
    }
    for (int r = 0; r < 3; ++r) /* Xq = V*U' :18 */
      for (int c = 0; c < 3; ++c)
        Xq[r][c] = (V[r][0] * U[c][0] + V[r][1] * U[c][1]) + V[r][2] * U[c][2];
    double mdet = orc_det3(Xq); /* :19 */
    double rd = round(mdet);
    if (rd == 1.0)
      state = 1;
    else if (rd == -1.0)
      state = -1; /* nsmall==1 cannot reach here (det forced +1); nsmall==0 -> zs(2)==0 :28,:33 */
    else
      state = 0;
  }
  if (state == 1) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) rot[3 * r + c] = Xq[r][c];
    for (int r = 0; r < 3; ++r) /* trans = ct1 - rot*ct2 :23 */
      trans[r] = ct1[r] - ((Xq[r][0] * ct2[0] + Xq[r][1] * ct2[1]) + Xq[r][2] * ct2[2]);
  } else { /* rot=H; trans=0  :35-36,:40-41 */
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) rot[3 * r + c] = H[r][c];
    trans[0] = trans[1] = trans[2] = 0.0;
  }
  return state;
}

/* ------------------------------------------------------------------------- */
/* Symmetric 4x4 cyclic Jacobi (stands in for MATLAB eig, absoluteOrientationQuaternion.m:87)
 *
 * SPEC (shared with 3pre_b200/csrc/fit.cuh):
 *   only the upper triangle of M is read (m[p][q], p<=q), E := I.
 *   scale2 = sum over p<=q (p-major order) of m[p][q]^2, computed once.
 *   repeat up to 20 sweeps over (0,1),(0,2),(0,3),(1,2),(1,3),(2,3):
 *     apq = m[p][q];  skip if apq == 0 or apq*apq <= 1e-34*scale2
 *     d = aqq - app;  g2 = 2*apq
 *     t = g2 / (|d| + sqrt(d*d + g2*g2));  if d < 0: t = -t
 *     c = 1/sqrt(1 + t*t);  s = c*t
 *     app' = app - t*apq;  aqq' = aqq + t*apq;  apq' = 0
 *     for r != p,q:  (arp, arq) := (c*arp - s*arq, s*arp + c*arq)
 *     for each row r of E: (erp, erq) := (c*erp - s*erq, s*erp + c*erq)
 *   stop when a full sweep skipped every pair.
 *   Returns the column of E whose diagonal entry is largest (first on ties) --
 *   Horn's definition; the reference takes E(:,4) which is the same vector when
 *   eig returns ascending eigenvalues (SURVEY.md 7 "eig order hazard").
 */
static void orc_eig4_max(double m[4][4], double e[4]) {
  static const int PP[6] = {0, 0, 0, 1, 1, 2}, QQ[6] = {1, 2, 3, 2, 3, 3};
  double E[4][4];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) E[r][c] = (r == c) ? 1.0 : 0.0;
  double scale2 = 0.0;
  for (int r = 0; r < 4; ++r)
    for (int c = r; c < 4; ++c) scale2 = scale2 + m[r][c] * m[r][c];
  for (int r = 0; r < 4; ++r) /* symmetrise from the upper triangle */
    for (int c = 0; c < r; ++c) m[r][c] = m[c][r];
  for (int sweep = 0; sweep < 20; ++sweep) {
    int rotated = 0;
    for (int k = 0; k < 6; ++k) {
      const int p = PP[k], q = QQ[k];
      double apq = m[p][q];
      if (apq == 0.0) continue;
      if (apq * apq <= 1e-34 * scale2) continue;
      rotated = 1;
      double app = m[p][p], aqq = m[q][q];
      double d = aqq - app;
      double g2 = 2.0 * apq;
      double t = g2 / (fabs(d) + sqrt(d * d + g2 * g2));
      if (d < 0.0) t = -t;
      double c = 1.0 / sqrt(1.0 + t * t);
      double s = c * t;
      m[p][p] = app - t * apq;
      m[q][q] = aqq + t * apq;
      m[p][q] = 0.0;
      m[q][p] = 0.0;
      for (int r = 0; r < 4; ++r) {
        if (r == p || r == q) continue;
        double arp = m[r][p], arq = m[r][q];
        double np = c * arp - s * arq;
        double nq = s * arp + c * arq;
        m[r][p] = np;
        m[p][r] = np;
        m[r][q] = nq;
        m[q][r] = nq;
      }
      for (int r = 0; r < 4; ++r) {
        double erp = E[r][p], erq = E[r][q];
        E[r][p] = c * erp - s * erq;
        E[r][q] = s * erp + c * erq;
      }
    }
    if (!rotated) break;
  }
  int jmax = 0;
  for (int j = 1; j < 4; ++j)
    if (m[j][j] > m[jmax][jmax]) jmax = j;
  for (int r = 0; r < 4; ++r) e[r] = E[r][jmax];
}

/* absoluteOrientationQuaternion  (M/absoluteOrientationQuaternion.m:28-127)
 * A, B: 3 x n column-major (optionally gathered through idx).  Model B ~ s*R*A + T.
 * R row-major 3x3.  Returns 0, or -1 when n < 4 (the reference errors, :51-54)
 * unless allow_small is set (k=3 extension). */
ORC_API int orc_horn(const double *A, const double *B, const int32_t *idx, int n,
                     int doScale, int allow_small, double *s_out, double *R,
                     double *T, double *err_out) {
  if (n < 4 && !allow_small) return -1;
  double Ca[3] = {0, 0, 0}, Cb[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i) { /* mean(A,2) :60-61 */
    const double *a = A + 3 * (size_t)(idx ? idx[i] : i);
    const double *b = B + 3 * (size_t)(idx ? idx[i] : i);
    for (int r = 0; r < 3; ++r) {
      Ca[r] += a[r];
      Cb[r] += b[r];
    }
  }
  for (int r = 0; r < 3; ++r) {
    Ca[r] = Ca[r] / (double)n;
    Cb[r] = Cb[r] / (double)n;
  }
  double M[4][4];
  memset(M, 0, sizeof M);
  for (int i = 0; i < n; ++i) { /* :69-84 */
    const double *pa = A + 3 * (size_t)(idx ? idx[i] : i);
    const double *pb = B + 3 * (size_t)(idx ? idx[i] : i);
    double a[4], b[4];
    a[0] = 0.0;
    b[0] = 0.0;
    for (int r = 0; r < 3; ++r) {
      a[r + 1] = pa[r] - Ca[r];
      b[r + 1] = pb[r] - Cb[r];
    }
    double Ma[4][4] = {{a[0], -a[1], -a[2], -a[3]},
                       {a[1], a[0], a[3], -a[2]},
                       {a[2], -a[3], a[0], a[1]},
                       {a[3], a[2], -a[1], a[0]}};
    double Mb[4][4] = {{b[0], -b[1], -b[2], -b[3]},
                       {b[1], b[0], -b[3], b[2]},
                       {b[2], b[3], b[0], -b[1]},
                       {b[3], -b[2], b[1], b[0]}};
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) { /* (Ma'*Mb)(r,c), k ascending */
        double acc = ((Ma[0][r] * Mb[0][c] + Ma[1][r] * Mb[1][c]) + Ma[2][r] * Mb[2][c]) +
                     Ma[3][r] * Mb[3][c];
        M[r][c] = M[r][c] + acc;
      }
  }
  double e[4];
  orc_eig4_max(M, e); /* :87-90 */
  double M1[4][4] = {{e[0], -e[1], -e[2], -e[3]},
                     {e[1], e[0], e[3], -e[2]},
                     {e[2], -e[3], e[0], e[1]},
                     {e[3], e[2], -e[1], e[0]}};
  double M2[4][4] = {{e[0], -e[1], -e[2], -e[3]},
                     {e[1], e[0], -e[3], e[2]},
                     {e[2], e[3], e[0], -e[1]},
                     {e[3], -e[2], e[1], e[0]}};
  double Rm[3][3];
  for (int r = 1; r < 4; ++r) /* R = (M1'*M2)(2:4,2:4) :101-104 */
    for (int c = 1; c < 4; ++c)
      Rm[r - 1][c - 1] =
          ((M1[0][r] * M2[0][c] + M1[1][r] * M2[1][c]) + M1[2][r] * M2[2][c]) +
          M1[3][r] * M2[3][c];
  double s = 1.0;
  if (doScale) { /* :106-112 */
    double sa = 0.0, sb = 0.0;
    for (int i = 0; i < n; ++i) {
      const double *pa = A + 3 * (size_t)(idx ? idx[i] : i);
      const double *pb = B + 3 * (size_t)(idx ? idx[i] : i);
      double an[3], bn[3], ran[3];
      for (int r = 0; r < 3; ++r) {
        an[r] = pa[r] - Ca[r];
        bn[r] = pb[r] - Cb[r];
      }
      for (int r = 0; r < 3; ++r)
        ran[r] = (Rm[r][0] * an[0] + Rm[r][1] * an[1]) + Rm[r][2] * an[2];
      sa = sa + ((bn[0] * ran[0] + bn[1] * ran[1]) + bn[2] * ran[2]);
      sb = sb + ((bn[0] * bn[0] + bn[1] * bn[1]) + bn[2] * bn[2]);
    }
    s = sb / sa;
  }
  for (int r = 0; r < 3; ++r) /* T = Cb - s*R*Ca :118 */
    T[r] = Cb[r] - ((s * Rm[r][0] * Ca[0] + s * Rm[r][1] * Ca[1]) + s * Rm[r][2] * Ca[2]);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = Rm[r][c];
  if (s_out) *s_out = s;
  if (err_out) { /* :121-127 */
    double err = 0.0;
    for (int i = 0; i < n; ++i) {
      const double *pa = A + 3 * (size_t)(idx ? idx[i] : i);
      const double *pb = B + 3 * (size_t)(idx ? idx[i] : i);
      double d[3];
      for (int r = 0; r < 3; ++r)
        d[r] = pb[r] - (((s * Rm[r][0] * pa[0] + s * Rm[r][1] * pa[1]) + s * Rm[r][2] * pa[2]) + T[r]);
      err = err + sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    }
    *err_out = err;
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Stage 3: support scoring  (RANSAC_CALC_VER2.m:121-125,135)                 */
/* ------------------------------------------------------------------------- */
/* Y0b = Rot*Yb + Trans; residu = Y0b - Ya; normResidu = sqrt(rx^2+ry^2+rz^2);
 * inlier = normResidu < thr (strict).  mask may be NULL.  errsum = sum of inlier
 * norms in index order. */
ORC_API int orc_score(const double *R, const double *T, const double *Ya,
                      const double *Yb, int N, double thr, uint8_t *mask,
                      double *errsum) {
  int count = 0;
  double es = 0.0;
  for (int i = 0; i < N; ++i) {
    const double *ya = Ya + 3 * (size_t)i, *yb = Yb + 3 * (size_t)i;
    double r[3];
    for (int k = 0; k < 3; ++k) {
      double y0 = ((R[3 * k] * yb[0] + R[3 * k + 1] * yb[1]) + R[3 * k + 2] * yb[2]) + T[k];
      r[k] = y0 - ya[k];
    }
    double nr = sqrt((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]);
    int in = nr < thr;
    if (mask) mask[i] = (uint8_t)in;
    if (in) {
      ++count;
      es = es + nr;
    }
  }
  if (errsum) *errsum = es;
  return count;
}

/* threshold override  (RANSAC_CALC_VER2.m:69-72) */
ORC_API double orc_distance_threshold(const double *Yb, int N) {
  if (N <= 0) return 0.0;
  int p = 0;
  for (int i = 1; i < N; ++i)
    if (Yb[3 * (size_t)i + 2] < Yb[3 * (size_t)p + 2]) p = i; /* first argmin */
  const double *y = Yb + 3 * (size_t)p;
  double dist = sqrt((y[0] * y[0] + y[1] * y[1]) + y[2] * y[2]);
  return 0.01 * dist;
}

/* adaptive iteration bound  (RANSAC_CALC_VER2.m:139 with mult=5,
 * RANSAC_CALC_VER_test.m:102 with mult=1).  MATLAB doubles: log(0)=-Inf,
 * x/(-Inf) = -0 -> ceil -> 0; card==0 -> log(1)=0 -> log(eps)/0 = -Inf. */
ORC_API double orc_adaptive_niter(int card, int nPoints, int k, int mult) {
  double w = (double)card / (double)nPoints;
  double v = (double)mult * ceil(log(0.01) / log(1.0 - pow(w, (double)k)));
  return v;
}

/* seeded sample sets (the stand-in for get_rand.m:43-48, which yields an
 * ascending k-subset).  SPEC shared with 3pre_b200/csrc/sample.cuh:
 *   x = splitmix64(seed ^ (pair*0x9E3779B97F4A7C15) ^ (hyp*0xD1B54A32D192ED03) ^ (draw+1)*0x8CB92BA72F3D8DD7)
 *   Floyd: for j = N-k .. N-1: t = (hi32(x_j) * (j+1)) >> 32; pick t unless already
 *   picked, else j.  Result sorted ascending. */
static uint64_t orc_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

ORC_API void orc_sample_set(uint64_t seed, uint32_t pair, uint32_t hyp, int N, int k,
                            int32_t *out) {
  int n = 0;
  for (int d = 0; d < k; ++d) {
    int j = N - k + d;
    uint64_t x = orc_splitmix64(seed ^ ((uint64_t)pair * 0x9E3779B97F4A7C15ULL) ^
                                ((uint64_t)hyp * 0xD1B54A32D192ED03ULL) ^
                                ((uint64_t)(d + 1) * 0x8CB92BA72F3D8DD7ULL));
    uint32_t r = (uint32_t)(x >> 32);
    int t = (int)(((uint64_t)r * (uint64_t)(j + 1)) >> 32);
    int dup = 0;
    for (int i = 0; i < n; ++i)
      if (out[i] == t) dup = 1;
    int pick = dup ? j : t;
    /* insert keeping ascending order */
    int pos = n;
    while (pos > 0 && out[pos - 1] > pick) {
      out[pos] = out[pos - 1];
      --pos;
    }
    out[pos] = pick;
    ++n;
  }
}

/* ------------------------------------------------------------------------- */
/* Stages 2-4: the RANSAC loop                                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
  int32_t status;        /* 0 ok, 1 = fewer than k correspondences (get_rand errors), 2 = no iteration ran */
  int32_t state;         /* State_RANSAC of the refit (VER2) / 1 (Horn variant) */
  int32_t best_fit;      /* BestFit = max cardinality */
  int32_t best_sample;   /* 0-based index into the supplied sample sets of the winner */
  int32_t best_iter;     /* 1-based BestFitIdx in the reference's M(iter) numbering */
  int32_t n_iter;        /* length(M): hypotheses recorded */
  int32_t n_consumed;    /* sample sets consumed (recorded + skipped state -1) */
  int32_t pad;
  double thr;            /* distance threshold actually used */
  double error_sum;      /* M(BestFitIdx).ErrorSum */
  double R[9];           /* row-major refit rotation */
  double T[3];
  double R_hyp[9];       /* the winning hypothesis before the refit (M(BestFitIdx).R) */
  double T_hyp[3];
} orc_ransac_result;

/* method 0: RANSAC_CALC_VER2.m  (find_transform_matrix fit + refit, threshold
 *           override :69-72, adaptive rule x5 :139, skip on state -1 :97-99)
 * method 1: RANSAC_CALC_VER_test.m (Horn fit :71 + refit :152, threshold from
 *           options :95, adaptive rule without x5 :102)
 * samples: k x H int32, 0-based, column h = sample set h (ascending for get_rand
 *          semantics, but used as given).  The loop consumes them in order and
 *          stops early when they run out (documented deviation: the reference
 *          draws fresh random sets for ever).
 * adaptive: 0 disables the nIterations update (fixed-H runs).
 * mask_out (N bytes, may be NULL): PositionInliers of the winner.
 * counts_out / states_out (H each, may be NULL): per supplied sample set, for
 * stage-wise parity (count = -1 when not evaluated). */
ORC_API void orc_ransac(const double *Ya, const double *Yb, int N, int method,
                        int k, int max_iteration, double distance_threshold,
                        int adaptive, const int32_t *samples, int H,
                        orc_ransac_result *res, uint8_t *mask_out,
                        int32_t *counts_out, int8_t *states_out) {
  memset(res, 0, sizeof *res);
  for (int h = 0; h < H; ++h) {
    if (counts_out) counts_out[h] = -1;
    if (states_out) states_out[h] = 0;
  }
  if (N < k || N <= 0) {
    res->status = 1;
    return;
  }
  const int mult = (method == 0) ? 5 : 1;
  double thr = (method == 0) ? orc_distance_threshold(Yb, N) : distance_threshold;
  res->thr = thr;
  double nIterations = (double)max_iteration; /* :45 */
  int maxSupport = 5;                          /* :46 (literal, also when k != 5) */
  int iter = 1;                                /* :76 */
  /* per recorded hypothesis: cardinality, error sum, sample index */
  int cap = max_iteration > 1 ? max_iteration : 1;
  int32_t *card = (int32_t *)calloc((size_t)cap, sizeof(int32_t));
  double *esum = (double *)calloc((size_t)cap, sizeof(double));
  int32_t *sidx = (int32_t *)calloc((size_t)cap, sizeof(int32_t));
  int h = 0;
  while ((double)iter < fmin(nIterations, (double)max_iteration)) { /* :86 */
    if (h >= H) break;
    const int32_t *set = samples + (size_t)k * h;
    double Rot[9], Trans[3];
    int st;
    if (method == 0) {
      st = orc_find_transform(Ya, Yb, set, k, Rot, Trans); /* :96 */
    } else {
      double s_, e_;
      orc_horn(Yb, Ya, set, k, 0, 1, &s_, Rot, Trans, &e_); /* VER_test :71 */
      st = 1;
    }
    if (states_out) states_out[h] = (int8_t)st;
    ++h;
    if (method == 0 && st == -1) continue; /* :97-99, iter not advanced */
    double es;
    int c = orc_score(Rot, Trans, Ya, Yb, N, thr, NULL, &es); /* :121-125,:135 */
    if (counts_out) counts_out[h - 1] = c;
    card[iter - 1] = c;
    esum[iter - 1] = es;
    sidx[iter - 1] = h - 1;
    if (c >= maxSupport) { /* :137-140 */
      maxSupport = c;
      if (adaptive) nIterations = orc_adaptive_niter(c, N, k, mult);
    }
    iter = iter + 1;
  }
  res->n_consumed = h;
  int L = iter - 1; /* length(M) */
  res->n_iter = L;
  if (L < 1) {
    res->status = 2;
    free(card);
    free(esum);
    free(sidx);
    return;
  }
  /* selection :165-175 */
  int maxc = card[0];
  for (int i = 1; i < L; ++i)
    if (card[i] > maxc) maxc = card[i];
  int best = 0;
  double beste = 0;
  for (int i = 0; i < L; ++i) {
    /* eee2(eee2~=max)=0; eee1(eee2==0)=10000 : note max==0 marks everything */
    double e1 = (card[i] != maxc || card[i] == 0) ? 10000.0 : esum[i];
    if (i == 0 || e1 < beste) { /* [C,I]=min: first minimum */
      beste = e1;
      best = i;
    }
  }
  res->best_fit = maxc;
  res->best_iter = best + 1;
  res->best_sample = sidx[best];
  /* recompute the winner (M(BestFitIdx).R/T/ErrorSum/PositionInliers) */
  const int32_t *set = samples + (size_t)k * sidx[best];
  double Rot[9], Trans[3];
  if (method == 0) {
    orc_find_transform(Ya, Yb, set, k, Rot, Trans);
  } else {
    double s_, e_;
    orc_horn(Yb, Ya, set, k, 0, 1, &s_, Rot, Trans, &e_);
  }
  uint8_t *mask = (uint8_t *)malloc((size_t)N);
  double es;
  int c = orc_score(Rot, Trans, Ya, Yb, N, thr, mask, &es);
  res->error_sum = es;
  memcpy(res->R_hyp, Rot, sizeof Rot);
  memcpy(res->T_hyp, Trans, sizeof Trans);
  /* refit on the support set :186 / VER_test :152 */
  int32_t *sup = (int32_t *)malloc(sizeof(int32_t) * (size_t)(c > 0 ? c : 1));
  int ns = 0;
  for (int i = 0; i < N; ++i)
    if (mask[i]) sup[ns++] = i;
  if (method == 0) {
    res->state = orc_find_transform(Ya, Yb, sup, ns, res->R, res->T);
  } else {
    double s_, e_;
    orc_horn(Yb, Ya, sup, ns, 0, 1, &s_, res->R, res->T, &e_);
    res->state = 1;
  }
  if (mask_out) memcpy(mask_out, mask, (size_t)N);
  free(sup);
  free(mask);
  free(card);
  free(esum);
  free(sidx);
}

/* Whole pair: siftmatch -> gather -> RANSAC (SIFT_match_save.m:33,:40-53).
 * desc1/desc2 double 128 x K column-major; xyz1/xyz2 3 x K.  sample sets come from
 * orc_sample_set(seed, pair_id, h, N, k).  Returns number of matches; matches
 * written 0-based into pairs (capacity K1). */
ORC_API int orc_pair(const double *desc1, const double *desc2, const double *xyz1,
                     const double *xyz2, int K1, int K2, int ND, double ratio,
                     int method, int k, int max_iteration, double distance_threshold,
                     int adaptive, uint64_t seed, uint32_t pair_id, int H,
                     int32_t *pairs, orc_ransac_result *res, uint8_t *mask_out) {
  double *score = (double *)malloc(sizeof(double) * (size_t)(K1 > 0 ? K1 : 1));
  int n = orc_siftmatch_f64(desc1, desc2, K1, K2, ND, ratio, pairs, score);
  free(score);
  double *Ya = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
  double *Yb = (double *)malloc(sizeof(double) * 3 * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i)
    for (int r = 0; r < 3; ++r) {
      Ya[3 * i + r] = xyz1[3 * (size_t)pairs[2 * i] + r];     /* Ya(:,matches(1,:)) */
      Yb[3 * i + r] = xyz2[3 * (size_t)pairs[2 * i + 1] + r]; /* Yb(:,matches(2,:)) */
    }
  int32_t *samples = (int32_t *)malloc(sizeof(int32_t) * (size_t)k * (size_t)(H > 0 ? H : 1));
  if (n >= k)
    for (int h = 0; h < H; ++h) orc_sample_set(seed, pair_id, (uint32_t)h, n, k, samples + (size_t)k * h);
  orc_ransac(Ya, Yb, n, method, k, max_iteration, distance_threshold, adaptive,
             samples, H, res, mask_out, NULL, NULL);
  free(samples);
  free(Ya);
  free(Yb);
  return n;
}

/* ------------------------------------------------------------------------- */
/* slamToolbox R2q  (M/slamToolbox_11_02_18/FrameTransforms/Rotations/R2q.m:11-55) */
/* ------------------------------------------------------------------------- */
ORC_API void orc_R2q(const double *R /* row-major */, double *q) {
  double T = ((R[0] + R[4]) + R[8]) + 1.0;
  double a, b, c, d;
  if (T > 0.00000001) {
    double S = 2.0 * sqrt(T);
    a = 0.25 * S;
    b = (R[5] - R[7]) / S; /* R(2,3)-R(3,2) */
    c = (R[6] - R[2]) / S; /* R(3,1)-R(1,3) */
    d = (R[1] - R[3]) / S; /* R(1,2)-R(2,1) */
  } else if (R[0] > R[4] && R[0] > R[8]) {
    double S = 2.0 * sqrt(1.0 + R[0] - R[4] - R[8]);
    a = (R[5] - R[7]) / S;
    b = 0.25 * S;
    c = (R[1] + R[3]) / S;
    d = (R[6] + R[2]) / S;
  } else if (R[4] > R[8]) {
    double S = 2.0 * sqrt(1.0 + R[4] - R[0] - R[8]);
    a = (R[6] - R[2]) / S;
    b = (R[1] + R[3]) / S;
    c = 0.25 * S;
    d = (R[5] + R[7]) / S;
  } else {
    double S = 2.0 * sqrt(1.0 + R[8] - R[0] - R[4]);
    a = (R[1] - R[3]) / S;
    b = (R[6] + R[2]) / S;
    c = (R[5] + R[7]) / S;
    d = 0.25 * S;
  }
  q[0] = a;
  q[1] = -b;
  q[2] = -c;
  q[3] = -d;
}
