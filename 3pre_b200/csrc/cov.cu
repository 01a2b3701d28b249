// cov.cu -- covariance of the RANSAC pose estimate (M/cov_est_RANSAC_deriv.m:1-244, call site
// M/mex_files/RANSAC_CALCULATION/RANSAC_CALC_VER2.m:204-206).
//
// For the support set of the winning hypothesis: E_k = |p_a - (q2R(q) p_b + T)|^2 in spherical coordinates of both
// points, gradient and second derivatives by NESTED CENTRAL DIFFERENCES (M/deriv.m:7-9, eps_xy = 1e-6), summed over the
// points; then dA_dz = G2tot \ [d2E/dx dz_1 .. dz_6] and sm_cov_censi = dA_dz blkdiag(R, R) dA_dz'.
//   k_cov_points  grid (S, P): block = 18 point lanes x 14 task columns (7 state columns of the Hessian, 6 noise
//                 columns, 1 for E and the gradient); a task is 28 evaluations of E (4 sincos each).  Per-block partial
//                 sums in a fixed order -> workspace.
//   k_cov_finish  one thread per pair: sums the S partials in order, LU with partial pivoting (MATLAB's `\`), the 7 x 7
//                 covariance.
// Numerical differentiation with eps = 1e-6 amplifies rounding by 1e12: results agree with the oracle / MATLAB to
// ~1e-6 relative, not bit for bit (tests/test_gpu_cov.py states the tolerance).
#include "cov.cuh"

namespace pre3 {

namespace {

constexpr int CV_LANES = 18, CV_COLS = 14, CV_THREADS = CV_LANES * CV_COLS;  // 252
constexpr int CV_SMAX = 64;
constexpr double CV_EPS = 0.000001;  // cov_est_RANSAC_deriv.m:26

// parameter vector: 0 b_theta 1 b_phi 2 b_r 3 a_theta 4 a_phi 5 a_r 6..9 quaternion 10..12 T
__device__ __forceinline__ double cov_E(const double* p) {  // E_k_func :236-241
  double sa, ca, sb, cb, st, ct;
  sincos(p[4], &sa, &ca);
  const double za = p[5] * sa, rca = p[5] * ca;
  sincos(p[3], &st, &ct);
  const double xa = rca * ct, ya = rca * st;
  sincos(p[1], &sb, &cb);
  const double zb = p[2] * sb, rcb = p[2] * cb;
  sincos(p[0], &st, &ct);
  const double xb = rcb * ct, yb = rcb * st;
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

__device__ __forceinline__ int xidx(int i) { return i < 3 ? 10 + i : 3 + i; }  // x = [T1 T2 T3 q1 q2 q3 q4]

__device__ void cov_grad(double* w, double* g) {  // gradEk :72-93 (w is restored)
#pragma unroll 1
  for (int i = 0; i < 7; ++i) {
    const int k = xidx(i);
    const double x = w[k];
    w[k] = x + CV_EPS / 2;
    const double f1 = cov_E(w);
    w[k] = x - CV_EPS / 2;
    const double f0 = cov_E(w);
    w[k] = x;
    g[i] = (f1 - f0) / CV_EPS;
  }
}

// slamToolbox R2q (R2q.m:11-55); Rc column-major
__device__ void cov_R2q(const double* Rc, double* q) {
  const double r11 = Rc[0], r21 = Rc[1], r31 = Rc[2], r12 = Rc[3], r22 = Rc[4], r32 = Rc[5], r13 = Rc[6], r23 = Rc[7],
               r33 = Rc[8];
  const double T = ((r11 + r22) + r33) + 1.0;
  double a, b, c, d;
  if (T > 0.00000001) {
    const double S = 2.0 * sqrt(T);
    a = 0.25 * S, b = (r23 - r32) / S, c = (r31 - r13) / S, d = (r12 - r21) / S;
  } else if (r11 > r22 && r11 > r33) {
    const double S = 2.0 * sqrt(1.0 + r11 - r22 - r33);
    a = (r23 - r32) / S, b = 0.25 * S, c = (r12 + r21) / S, d = (r31 + r13) / S;
  } else if (r22 > r33) {
    const double S = 2.0 * sqrt(1.0 + r22 - r11 - r33);
    a = (r31 - r13) / S, b = (r12 + r21) / S, c = 0.25 * S, d = (r23 + r32) / S;
  } else {
    const double S = 2.0 * sqrt(1.0 + r33 - r11 - r22);
    a = (r12 - r21) / S, b = (r31 + r13) / S, c = (r23 + r32) / S, d = 0.25 * S;
  }
  q[0] = a, q[1] = -b, q[2] = -c, q[3] = -d;
}

// partial[(p * S + s) * 14 + col][8]: col 0..6 Hessian columns, 7..12 noise columns, 13: {Gtot[7], Etot}
__global__ void __launch_bounds__(CV_THREADS)
k_cov_points(const double* __restrict__ Ya, const double* __restrict__ Yb, const int32_t* __restrict__ n_corr,
             const uint8_t* __restrict__ masks, int Nmax, const double* __restrict__ RT, int rt_stride,
             double* __restrict__ partial) {
  __shared__ double sred[CV_LANES][CV_COLS][8];
  const int p = blockIdx.y, s = blockIdx.x, S = gridDim.x;
  const int col = threadIdx.x % CV_COLS, ln = threadIdx.x / CV_COLS;
  const int n = n_corr ? min(n_corr[p], Nmax) : Nmax;
  const int per = (n + S - 1) / S;
  const int i0 = s * per, i1 = min(n, i0 + per);
  double q[4];
  cov_R2q(RT + (size_t)p * rt_stride, q);
  const double* Tp = RT + (size_t)p * rt_stride + 9;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = i0 + ln; i < i1; i += CV_LANES) {
    if (masks && !masks[(size_t)p * Nmax + i]) continue;  // SupportSet of the winner (RANSAC_CALC_VER2.m:125-127)
    const double* pa = Ya + ((size_t)p * Nmax + i) * 3;
    const double* pb = Yb + ((size_t)p * Nmax + i) * 3;
    double w[13];
    // cart2sph :50-51
    w[0] = atan2(pb[1], pb[0]), w[1] = atan2(pb[2], hypot(pb[0], pb[1])), w[2] = hypot(hypot(pb[0], pb[1]), pb[2]);
    w[3] = atan2(pa[1], pa[0]), w[4] = atan2(pa[2], hypot(pa[0], pa[1])), w[5] = hypot(hypot(pa[0], pa[1]), pa[2]);
    w[6] = q[0], w[7] = q[1], w[8] = q[2], w[9] = q[3];
    w[10] = Tp[0], w[11] = Tp[1], w[12] = Tp[2];
    double g1[7];
    if (col == 13) {
      acc[7] += cov_E(w);  // :150
      cov_grad(w, g1);     // :153
#pragma unroll
      for (int r = 0; r < 7; ++r) acc[r] += g1[r];
    } else {
      const int k = col < 7 ? xidx(col) : col - 7;  // :156 / :160-175
      const double x = w[k];
      double g0[7];
      w[k] = x + CV_EPS / 2;
      cov_grad(w, g1);
      w[k] = x - CV_EPS / 2;
      cov_grad(w, g0);
#pragma unroll
      for (int r = 0; r < 7; ++r) acc[r] += (g1[r] - g0[r]) / CV_EPS;
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) sred[ln][col][r] = acc[r];
  __syncthreads();
  if (threadIdx.x < CV_COLS * 8) {
    const int c = threadIdx.x / 8, r = threadIdx.x % 8;
    double sum = 0.0;
    for (int l = 0; l < CV_LANES; ++l) sum += sred[l][c][r];
    partial[(((size_t)p * S + s) * CV_COLS + c) * 8 + r] = sum;
  }
}

__global__ void k_cov_finish(const double* __restrict__ partial, int S, int P, const int32_t* __restrict__ n_corr,
                             const uint8_t* __restrict__ masks, int Nmax, pre3_cov_result* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double sum[CV_COLS][8];
  for (int c = 0; c < CV_COLS; ++c)
    for (int r = 0; r < 8; ++r) sum[c][r] = 0.0;
  for (int s = 0; s < S; ++s)
    for (int c = 0; c < CV_COLS; ++c)
      for (int r = 0; r < 8; ++r) sum[c][r] += partial[(((size_t)p * S + s) * CV_COLS + c) * 8 + r];
  const int n = n_corr ? min(n_corr[p], Nmax) : Nmax;
  int k = n;
  if (masks) {
    k = 0;
    for (int i = 0; i < n; ++i) k += masks[(size_t)p * Nmax + i] ? 1 : 0;
  }
  pre3_cov_result o;
  double A[49], B[42];
  for (int j = 0; j < 7; ++j)
    for (int r = 0; r < 7; ++r) A[r + 7 * j] = o.G2tot[r + 7 * j] = sum[j][r];
  for (int j = 0; j < 6; ++j)
    for (int r = 0; r < 7; ++r) B[r + 7 * j] = sum[7 + j][r];
  for (int r = 0; r < 7; ++r) o.Gtot[r] = sum[13][r];
  o.Etot = sum[13][7];
  o.s2 = o.Etot / (double)(k - 3);  // :216-217
  o.n = k;
  // G2tot \ B: LU with partial pivoting (:200)
  int sing = 0;
  for (int kk = 0; kk < 7; ++kk) {
    int piv = kk;
    double best = fabs(A[kk + 7 * kk]);
    for (int r = kk + 1; r < 7; ++r)
      if (fabs(A[r + 7 * kk]) > best) best = fabs(A[r + 7 * kk]), piv = r;
    if (piv != kk) {
      for (int c = 0; c < 7; ++c) {
        const double t = A[kk + 7 * c];
        A[kk + 7 * c] = A[piv + 7 * c], A[piv + 7 * c] = t;
      }
      for (int c = 0; c < 6; ++c) {
        const double t = B[kk + 7 * c];
        B[kk + 7 * c] = B[piv + 7 * c], B[piv + 7 * c] = t;
      }
    }
    if (A[kk + 7 * kk] == 0.0) sing = 1;
    for (int r = kk + 1; r < 7; ++r) {
      const double f = A[r + 7 * kk] / A[kk + 7 * kk];
      for (int c = kk + 1; c < 7; ++c) A[r + 7 * c] = A[r + 7 * c] - f * A[kk + 7 * c];
      for (int c = 0; c < 6; ++c) B[r + 7 * c] = B[r + 7 * c] - f * B[kk + 7 * c];
    }
  }
  for (int c = 0; c < 6; ++c)
    for (int r = 6; r >= 0; --r) {
      double s = B[r + 7 * c];
      for (int j = r + 1; j < 7; ++j) s = s - A[r + 7 * j] * B[j + 7 * c];
      B[r + 7 * c] = s / A[r + 7 * r];
    }
  for (int i = 0; i < 42; ++i) o.dA_dz[i] = B[i];
  const double pi = 3.14159265358979323846;
  const double sg[3] = {0.02 * pi / 180, 0.02 * pi / 180, 0.015};  // :209
  for (int c = 0; c < 7; ++c)
    for (int r = 0; r < 7; ++r) {
      double s = 0.0;
      for (int j = 0; j < 6; ++j) s = s + (B[r + 7 * j] * (sg[j % 3] * sg[j % 3])) * B[c + 7 * j];  // :215
      o.cov[r + 7 * c] = s;
    }
  o.status = k < 1 ? 1 : (sing ? 2 : 0);
  out[p] = o;
}

int cov_splits(int P, int Nmax, int sm_count) {
  const int want = (2 * sm_count + P - 1) / P;                     // fill the chip when there are few pairs
  const int most = (Nmax + CV_LANES - 1) / CV_LANES;               // at least one point per lane
  return std::max(1, std::min(CV_SMAX, std::min(want, most)));
}

}  // namespace

size_t cov_workspace_bytes(int P, int Nmax) {
  return align_up((size_t)P * CV_SMAX * CV_COLS * 8 * sizeof(double)) + 4096;
}

int launch_cov_est(pre3_ctx* ctx, const double* dYa, const double* dYb, const int32_t* dn_corr, const uint8_t* dmasks,
                   int P, int Nmax, const double* dRT, int rt_stride, pre3_cov_result* dout) {
  Span span__(ctx, T_OTHER);
  if (P <= 0) return PRE3_OK;
  const int S = cov_splits(P, Nmax, ctx->sm_count);
  double* partial = ws_take<double>(ctx, (size_t)P * S * CV_COLS * 8);
  k_cov_points<<<dim3(S, P), CV_THREADS, 0, ctx->stream>>>(dYa, dYb, dn_corr, dmasks, Nmax, dRT, rt_stride, partial);
  k_cov_finish<<<(P + 63) / 64, 64, 0, ctx->stream>>>(partial, S, P, dn_corr, dmasks, Nmax, dout);
  count_launch(ctx, 2);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

}  // namespace pre3
