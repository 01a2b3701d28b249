// fit.cuh -- fp64 rigid fits, one hypothesis per thread.
//
// Implements, in the same operation order as the specification comments in
// oracle/pre3_oracle.c (which restates the reference .m files), so that the CPU checker
// and these kernels produce bit-identical (R, t, state) and therefore bit-identical
// inlier masks:
//   * find_transform_matrix  (M/mex_files/RANSAC_CALCULATION/find_transform_matrix.m:2-42)
//     with MATLAB's svd replaced by a fixed-order one-sided Jacobi SVD;
//   * absoluteOrientationQuaternion (M/absoluteOrientationQuaternion.m:56-127) with eig
//     replaced by a fixed-order cyclic Jacobi iteration, eigenvector of the largest
//     eigenvalue.
// The translation unit MUST be compiled with -fmad=false: only explicit fma()/fmaf()
// calls may fuse.  Uses + - * / sqrt only (all IEEE round-to-nearest in fp64 on sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pre3 {

struct Rigid {
  double R[9];  // row-major
  double t[3];
};

// One-sided Jacobi SVD of the 3x3 matrix a (row-major a[3*r+c]); on exit the columns of a
// are sigma_j*u_j, v accumulates the right rotations.  Columns are not sorted.
__device__ __forceinline__ void svd3_cols(double* a, double* v, double* sig) {
#pragma unroll
  for (int i = 0; i < 9; ++i) v[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 15; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int p = (k == 2) ? 1 : 0;
      const int q = (k == 0) ? 1 : 2;
      const double alpha = (a[p] * a[p] + a[3 + p] * a[3 + p]) + a[6 + p] * a[6 + p];
      const double beta = (a[q] * a[q] + a[3 + q] * a[3 + q]) + a[6 + q] * a[6 + q];
      const double gamma = (a[p] * a[q] + a[3 + p] * a[3 + q]) + a[6 + p] * a[6 + q];
      if (gamma == 0.0) continue;
      if (gamma * gamma <= 1e-28 * (alpha * beta)) continue;
      rotated = true;
      const double d = beta - alpha;
      const double g2 = 2.0 * gamma;
      double t = g2 / (fabs(d) + sqrt(d * d + g2 * g2));
      if (d < 0.0) t = -t;
      const double c = 1.0 / sqrt(1.0 + t * t);
      const double s = c * t;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double xp = a[3 * r + p], xq = a[3 * r + q];
        a[3 * r + p] = c * xp - s * xq;
        a[3 * r + q] = s * xp + c * xq;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const double xp = v[3 * r + p], xq = v[3 * r + q];
        v[3 * r + p] = c * xp - s * xq;
        v[3 * r + q] = s * xp + c * xq;
      }
    }
    if (!rotated) break;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) sig[j] = sqrt((a[j] * a[j] + a[3 + j] * a[3 + j]) + a[6 + j] * a[6 + j]);
}

__device__ __forceinline__ double det3(const double* m) {
  return (m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6])) +
         m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// V*U' of a well-conditioned H by Newton's polar iteration instead of the SVD: SPEC in oracle/pre3_oracle.c
// (orc_polar_fast), same operations in the same order.  Seven iterations (~1300 instructions) instead of ~5 Jacobi sweeps
// (~2800), the same count in every lane, and
// a reflection (det H < 0, no singular value near the threshold: state -1, skipped uncounted by the RANSAC loop) costs
// the determinant only.  Returns 1 (Xq filled), -1 (reflection), 0 (use the SVD).
__device__ __forceinline__ int polar_fast(const double* H, double threshold, double* Xq) {
  double nf2 = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) nf2 = nf2 + H[i] * H[i];
  const double detH = det3(H);
  if (!(nf2 > 0.0)) return 0;
  if (!(fabs(detH) > 50.0 * threshold * nf2)) return 0;
  const double rn = sqrt(nf2);
  if (!(fabs(detH) > 0.000000001 * (nf2 * rn))) return 0;
  if (detH < 0.0) return -1;
  const double sc = 1.0 / rn;
  double X[9], C[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) X[i] = H[i] * sc;
  for (int it = 0; it < 40; ++it) {
    C[0] = X[4] * X[8] - X[5] * X[7];
    C[1] = X[5] * X[6] - X[3] * X[8];
    C[2] = X[3] * X[7] - X[4] * X[6];
    C[3] = X[2] * X[7] - X[1] * X[8];
    C[4] = X[0] * X[8] - X[2] * X[6];
    C[5] = X[1] * X[6] - X[0] * X[7];
    C[6] = X[1] * X[5] - X[2] * X[4];
    C[7] = X[2] * X[3] - X[0] * X[5];
    C[8] = X[0] * X[4] - X[1] * X[3];
    const double dt = (X[0] * C[0] + X[1] * C[1]) + X[2] * C[2];
    double hx = 0.5, hc;
    if (it < 3) {  // scaled steps: 7 iterations suffice whatever the conditioning of the admitted H
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        a = a + C[i] * C[i];
        b = b + X[i] * X[i];
      }
      const double g = sqrt(sqrt(a / ((dt * dt) * b)));
      hx = 0.5 * g;
      hc = 0.5 / (dt * g);
    } else {
      hc = 0.5 / dt;
    }
    double diff2 = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double y = hx * X[i] + C[i] * hc;
      const double e = y - X[i];
      diff2 = diff2 + e * e;
      X[i] = y;
    }
    if (diff2 <= 1e-28) break;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Xq[3 * r + c] = X[3 * c + r];
  return 1;
}

// Tail of find_transform_matrix once H (row-major), ct1, ct2 are known (:17-42).
// Returns state; fills out (rot row-major, trans).
// threshold: 1e-11 (find_transform_matrix.m:20) or 1e-14 (code_from_dr_ye/find_transform_matrix_dr_ye.m:19).
__device__ __forceinline__ int kabsch_from_H(const double* H, const double* ct1, const double* ct2,
                                             Rigid& out, double threshold = 0.00000000001) {
  bool finite = true;
#pragma unroll
  for (int i = 0; i < 9; ++i) finite = finite && isfinite(H[i]);
  int state;
  double A[9], V[9], sig[3], U[9], Xq[9];
  int nsmall = 0, jsmall = -1;
  const int fast = finite ? polar_fast(H, threshold, Xq) : 0;
  if (finite && fast == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = H[i];
    svd3_cols(A, V, sig);
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (sig[j] < threshold) {
        ++nsmall;
        jsmall = j;
      }
  }
  if (!finite) {
    state = 0;
  } else if (fast != 0) {
    state = fast;  // 1: Xq from the polar iteration; -1: reflection
  } else if (nsmall >= 2) {
    state = -1;
  } else {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (j != jsmall) {
        const double inv = 1.0 / sig[j];
#pragma unroll
        for (int r = 0; r < 3; ++r) U[3 * r + j] = A[3 * r + j] * inv;
      }
    if (nsmall == 1) {
      // u_j := u_{j+1} x u_{j+2} (cyclic): written per case so that indices stay static
      if (jsmall == 0) {
        U[0] = U[3 + 1] * U[6 + 2] - U[6 + 1] * U[3 + 2];
        U[3] = U[6 + 1] * U[0 + 2] - U[0 + 1] * U[6 + 2];
        U[6] = U[0 + 1] * U[3 + 2] - U[3 + 1] * U[0 + 2];
      } else if (jsmall == 1) {
        U[1] = U[3 + 2] * U[6 + 0] - U[6 + 2] * U[3 + 0];
        U[4] = U[6 + 2] * U[0 + 0] - U[0 + 2] * U[6 + 0];
        U[7] = U[0 + 2] * U[3 + 0] - U[3 + 2] * U[0 + 0];
      } else {
        U[2] = U[3 + 0] * U[6 + 1] - U[6 + 0] * U[3 + 1];
        U[5] = U[6 + 0] * U[0 + 1] - U[0 + 0] * U[6 + 1];
        U[8] = U[0 + 0] * U[3 + 1] - U[3 + 0] * U[0 + 1];
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        Xq[3 * r + c] = (V[3 * r] * U[3 * c] + V[3 * r + 1] * U[3 * c + 1]) + V[3 * r + 2] * U[3 * c + 2];
    const double rd = round(det3(Xq));
    if (rd == 1.0)
      state = 1;
    else if (rd == -1.0)
      state = -1;
    else
      state = 0;
  }
  if (state == 1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) out.R[i] = Xq[i];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      out.t[r] = ct1[r] - ((Xq[3 * r] * ct2[0] + Xq[3 * r + 1] * ct2[1]) + Xq[3 * r + 2] * ct2[2]);
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) out.R[i] = H[i];
    out.t[0] = out.t[1] = out.t[2] = 0.0;
  }
  return state;
}

// find_transform_matrix on n points delivered by an accessor get(i, ya[3], yb[3])
// (ya = pset1 = previous frame, yb = pset2 = current frame).  KN > 0 fixes n at compile
// time (loops unroll, points live in registers); KN == 0 uses the runtime n.  Either way
// the arithmetic sequence is the one of find_transform_matrix.m:11-15.
template <int KN, typename Get>
__device__ __forceinline__ int fit_kabsch(int n_rt, Get get, Rigid& out, double threshold = 0.00000000001) {
  const int n = KN > 0 ? KN : n_rt;
  double ct1[3] = {0, 0, 0}, ct2[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double p1[3], p2[3];
    get(i, p1, p2);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      ct1[r] += p1[r];
      ct2[r] += p2[r];
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    ct1[r] = ct1[r] / (double)n;
    ct2[r] = ct2[r] / (double)n;
  }
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double p1[3], p2[3], q1[3], q2[3];
    get(i, p1, p2);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      q1[r] = p1[r] - ct1[r];
      q2[r] = p2[r] - ct2[r];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) H[3 * r + c] = H[3 * r + c] + q2[r] * q1[c];
  }
  return kabsch_from_H(H, ct1, ct2, out, threshold);
}

// Cyclic Jacobi on the symmetric 4x4 m (row-major, upper triangle authoritative);
// returns the eigenvector of the largest eigenvalue in e.
__device__ __forceinline__ void eig4_max(double* m, double* e) {
  double E[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) E[i] = (i % 5 == 0) ? 1.0 : 0.0;
  double scale2 = 0.0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = r; c < 4; ++c) scale2 = scale2 + m[4 * r + c] * m[4 * r + c];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < r; ++c) m[4 * r + c] = m[4 * c + r];
  for (int sweep = 0; sweep < 20; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int p = (k < 3) ? 0 : ((k < 5) ? 1 : 2);
      const int q = (k == 0) ? 1 : ((k == 1 || k == 3) ? 2 : 3);
      const double apq = m[4 * p + q];
      if (apq == 0.0) continue;
      if (apq * apq <= 1e-34 * scale2) continue;
      rotated = true;
      const double app = m[4 * p + p], aqq = m[4 * q + q];
      const double d = aqq - app;
      const double g2 = 2.0 * apq;
      double t = g2 / (fabs(d) + sqrt(d * d + g2 * g2));
      if (d < 0.0) t = -t;
      const double c = 1.0 / sqrt(1.0 + t * t);
      const double s = c * t;
      m[4 * p + p] = app - t * apq;
      m[4 * q + q] = aqq + t * apq;
      m[4 * p + q] = 0.0;
      m[4 * q + p] = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r == p || r == q) continue;
        const double arp = m[4 * r + p], arq = m[4 * r + q];
        const double np = c * arp - s * arq;
        const double nq = s * arp + c * arq;
        m[4 * r + p] = np;
        m[4 * p + r] = np;
        m[4 * r + q] = nq;
        m[4 * q + r] = nq;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double erp = E[4 * r + p], erq = E[4 * r + q];
        E[4 * r + p] = c * erp - s * erq;
        E[4 * r + q] = s * erp + c * erq;
      }
    }
    if (!rotated) break;
  }
  // largest diagonal entry, first on ties; static selects instead of dynamic indexing
  double best = m[0];
#pragma unroll
  for (int r = 0; r < 4; ++r) e[r] = E[4 * r];
#pragma unroll
  for (int j = 1; j < 4; ++j) {
    const bool gt = m[5 * j] > best;
    if (gt) {
      best = m[5 * j];
#pragma unroll
      for (int r = 0; r < 4; ++r) e[r] = E[4 * r + j];
    }
  }
}

// Adds (Ma'*Mb) of one centred point pair to M (absoluteOrientationQuaternion.m:71-83).
// a = [0; an], b = [0; bn].  The products with the structural zeros a(1), b(1) contribute
// exact +-0 terms in the reference's 4-term dot products; dropping them changes no value.
__device__ __forceinline__ void horn_accumulate(double* M, const double* an, const double* bn) {
  const double a1 = an[0], a2 = an[1], a3 = an[2];
  const double b1 = bn[0], b2 = bn[1], b3 = bn[2];
  // Ma = [0 -a1 -a2 -a3; a1 0 a3 -a2; a2 -a3 0 a1; a3 a2 -a1 0], Mb = [0 -b1 -b2 -b3; b1 0 -b3 b2; b2 b3 0 -b1; b3 -b2 b1 0]
  // (Ma'*Mb)(r,c) = sum_k Ma(k,r)*Mb(k,c), k ascending, zero products omitted.
  // Only the upper triangle is formed: eig4_max reads nothing else (the reference's M is
  // symmetric up to rounding and the specification symmetrises from the upper triangle).
  M[0] = M[0] + ((a1 * b1 + a2 * b2) + a3 * b3);
  M[1] = M[1] + (a2 * b3 + a3 * (-b2));
  M[2] = M[2] + (a1 * (-b3) + a3 * b1);
  M[3] = M[3] + (a1 * b2 + a2 * (-b1));
  M[5] = M[5] + (((-a1) * (-b1) + (-a3) * b3) + a2 * (-b2));
  M[6] = M[6] + ((-a1) * (-b2) + a2 * b1);
  M[7] = M[7] + ((-a1) * (-b3) + (-a3) * (-b1));
  M[10] = M[10] + (((-a2) * (-b2) + a3 * (-b3)) + (-a1) * b1);
  M[11] = M[11] + ((-a2) * (-b3) + a3 * b2);
  M[15] = M[15] + (((-a3) * (-b3) + (-a2) * b2) + a1 * (-b1));
}

// Tail of absoluteOrientationQuaternion once M, Ca, Cb are known (doScale = 0): R, T.
__device__ __forceinline__ void horn_from_M(double* M, const double* Ca, const double* Cb, Rigid& out) {
  double e[4];
  eig4_max(M, e);
  const double M1[16] = {e[0], -e[1], -e[2], -e[3], e[1], e[0], e[3], -e[2],
                         e[2], -e[3], e[0], e[1], e[3], e[2], -e[1], e[0]};
  const double M2[16] = {e[0], -e[1], -e[2], -e[3], e[1], e[0], -e[3], e[2],
                         e[2], e[3], e[0], -e[1], e[3], -e[2], e[1], e[0]};
#pragma unroll
  for (int r = 1; r < 4; ++r)
#pragma unroll
    for (int c = 1; c < 4; ++c)
      out.R[3 * (r - 1) + (c - 1)] =
          ((M1[r] * M2[c] + M1[4 + r] * M2[4 + c]) + M1[8 + r] * M2[8 + c]) + M1[12 + r] * M2[12 + c];
  // T = Cb - s*R*Ca with s = 1: (1*R(r,c))*Ca(c) == R(r,c)*Ca(c)
#pragma unroll
  for (int r = 0; r < 3; ++r)
    out.t[r] = Cb[r] - ((out.R[3 * r] * Ca[0] + out.R[3 * r + 1] * Ca[1]) + out.R[3 * r + 2] * Ca[2]);
}

// absoluteOrientationQuaternion(A = current (Yb), B = previous (Ya), 0) on n points from
// the accessor get(i, ya[3], yb[3]):  Ya ~ R*Yb + T (RANSAC_CALC_VER_test.m:71,:152).
template <int KN, typename Get>
__device__ __forceinline__ int fit_horn(int n_rt, Get get, Rigid& out) {
  const int n = KN > 0 ? KN : n_rt;
  double Ca[3] = {0, 0, 0}, Cb[3] = {0, 0, 0};  // Ca: centroid of A = Yb, Cb: of B = Ya
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double ya[3], yb[3];
    get(i, ya, yb);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Ca[r] += yb[r];
      Cb[r] += ya[r];
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    Ca[r] = Ca[r] / (double)n;
    Cb[r] = Cb[r] / (double)n;
  }
  double M[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) M[i] = 0.0;
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double ya[3], yb[3], an[3], bn[3];
    get(i, ya, yb);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      an[r] = yb[r] - Ca[r];
      bn[r] = ya[r] - Cb[r];
    }
    horn_accumulate(M, an, bn);
  }
  horn_from_M(M, Ca, Cb, out);
  return 1;
}

// Exact fp64 residual norm of one correspondence under (R,t), in the reference's operation
// order (RANSAC_CALC_VER2.m:121-123 as restated by orc_score).
__device__ __forceinline__ double residual_norm(const double* R, const double* t, const double* ya,
                                                const double* yb) {
  double r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double y0 = ((R[3 * k] * yb[0] + R[3 * k + 1] * yb[1]) + R[3 * k + 2] * yb[2]) + t[k];
    r[k] = y0 - ya[k];
  }
  return sqrt((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]);
}

// Squared distance d_diff of M/code_from_dr_ye/ransac_dr_ye.m:61-68 (sum of squares from 0.0, no sqrt).
__device__ __forceinline__ double residual_sq(const double* R, const double* t, const double* ya,
                                              const double* yb) {
  double d = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double y0 = ((R[3 * k] * yb[0] + R[3 * k + 1] * yb[1]) + R[3 * k + 2] * yb[2]) + t[k];
    const double e = y0 - ya[k];
    d = d + e * e;
  }
  return d;
}

}  // namespace pre3
