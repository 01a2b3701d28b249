// ekf_update.cu -- the EKF partial updates around ransac_hypotheses (first part).
//
// Reference:
//   M/update.m:27-56                                 S = H P H' + R;  K = P H' inv(S);  x + K (z - h);  P - K S K';
//                                                    0.5 P + 0.5 P';  quaternion normalisation Jacobian; q / |q|
//   M/@ekf_filter/ekf_update_li_inliers.m:15-29      z, h, H stacked from the low-innovation inliers, R = eye
//   M/@ekf_filter/ekf_update_hi_inliers.m:18-32      the same from the high-innovation inliers, on x_k_k / p_k_k
//   M/@ekf_filter/rescue_hi_inliers.m:35-46          nu' inv(H p_k_k H') nu < 5.9915
//   M/normJac.m:1-16
// Checker: oracle/ref_numpy_ekf.py (update, ekf_update_inliers, rescue_hi_inliers) -- dense numpy / LAPACK.  This is a
// floating-point path (dense fp64 linear algebra): results are specified to a tolerance (1e-9 relative to max|P|), not
// bit for bit, so the kernels use fused multiply-adds (explicit fma(): the translation unit is compiled -fmad=false).
//
// Per frame, with n states (cfg4: 1213) and m = 2 x (flagged features) stacked rows (cfg4: ~320):
//   k_upd_index   flagged features in feature order, m
//   k_upd_G       G = P H'            n x m   19 structural non-zeros of H per row (13 camera + 6 feature columns)
//   k_upd_S       S = H G + I         m x m
//   k_inv_panel / k_inv_update   inv(S): in-place blocked Gauss-Jordan (S is symmetric positive definite)   2 m^3 FLOP
//   k_dgemm       K = G inv(S)        n x m x m
//   k_dsyrk_update  0.5 (P + P') - G K'   lower tiles + mirrored write (K S K' = G K' is symmetric)   n^2 m FLOP: the bulk
//   k_upd_x       x + K (z - h)
//   k_upd_quat    rows / columns 4:7 through normJac, q / |q|
// Everything is column-major like MATLAB.
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace pre3 {
namespace {

struct UpdDims {
  int n, F, Mmax;  // Mmax = 2 F
};

// ---- flagged features in order ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_upd_index(const uint8_t* __restrict__ sel, int F, int32_t* __restrict__ idx,
                                                   int32_t* __restrict__ m_out) {
  __shared__ int s_w[8];
  __shared__ int s_base;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < F; i0 += 256) {
    const int i = i0 + tid;
    const bool on = i < F && sel[(size_t)f * F + i] == 1;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    int off = s_base + __popc(bal & ((1u << lane) - 1u)), tot = 0;
    for (int w = 0; w < 8; ++w) {
      if (w < warp) off += s_w[w];
      tot += s_w[w];
    }
    if (on) idx[(size_t)f * F + off] = i;
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
  if (tid == 0) m_out[f] = 2 * s_base;
}

// column c of H row (feature i, component a): cols 0..12 camera, then pos .. pos+5 (cartesian: 3)
__device__ __forceinline__ int h_col(int t, int pos) { return t < 13 ? t : pos + (t - 13); }
__device__ __forceinline__ double h_val(int t, int a, const double* __restrict__ hc, const double* __restrict__ hf) {
  return t < 13 ? hc[2 * t + a] : hf[2 * (t - 13) + a];  // 2 x 13 / 2 x 6 column-major blocks
}

// G = P H': thread = state row r, marching over the frame's flagged features.  The 13 camera columns of P(r, :) stay in
// registers for the whole march, the six feature columns are read once per feature and serve both measurement rows:
// 6 loads + 2 stores per feature instead of 38 loads.  H blocks of 32 features at a time are staged in shared memory.
constexpr int GF = 32;
__global__ void __launch_bounds__(256)
k_upd_G(const double* __restrict__ P, UpdDims d, const int32_t* __restrict__ idx, const int32_t* __restrict__ mf,
        const int32_t* __restrict__ type, const int32_t* __restrict__ pos, const double* __restrict__ Hcam,
        const double* __restrict__ Hfeat, double* __restrict__ G) {
  __shared__ double s_hc[GF][26];
  __shared__ double s_hf[GF][12];
  __shared__ int s_pos[GF], s_nf[GF];
  const int f = blockIdx.y;
  const int Lall = mf[f] >> 1;
  // gridDim.z blocks share the features of a frame (more blocks when few frames are in flight)
  const int per = (Lall + gridDim.z - 1) / gridDim.z;
  const int jbeg = blockIdx.z * per, L = min(Lall, jbeg + per);
  if (jbeg >= L) return;
  const int r = blockIdx.x * 256 + threadIdx.x;
  const bool live = r < d.n;
  const double* Pf = P + (size_t)f * d.n * d.n;
  double* Gf = G + (size_t)f * d.Mmax * d.n;
  double pc[13];
#pragma unroll
  for (int t = 0; t < 13; ++t) pc[t] = live ? Pf[(size_t)t * d.n + r] : 0.0;
  for (int j0 = jbeg; j0 < L; j0 += GF) {
    const int nj = min(GF, L - j0);
    __syncthreads();
    for (int e = threadIdx.x; e < nj * 38; e += 256) {
      const int jj = e / 38, q = e - jj * 38;
      const size_t fi = (size_t)f * d.F + idx[(size_t)f * d.F + j0 + jj];
      if (q < 26) s_hc[jj][q] = Hcam[fi * 26 + q];
      else s_hf[jj][q - 26] = Hfeat[fi * 12 + (q - 26)];
    }
    for (int jj = threadIdx.x; jj < nj; jj += 256) {
      const size_t fi = (size_t)f * d.F + idx[(size_t)f * d.F + j0 + jj];
      s_pos[jj] = pos[fi];
      s_nf[jj] = type[fi] == 0 ? 6 : 3;
    }
    __syncthreads();
    if (!live) continue;
    for (int jj = 0; jj < nj; ++jj) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int t = 0; t < 13; ++t) {
        a0 = fma(pc[t], s_hc[jj][2 * t], a0);
        a1 = fma(pc[t], s_hc[jj][2 * t + 1], a1);
      }
      const double* pcol = Pf + (size_t)s_pos[jj] * d.n + r;
      const int nf = s_nf[jj];
      for (int u = 0; u < nf; ++u) {
        const double pv = pcol[(size_t)u * d.n];
        a0 = fma(pv, s_hf[jj][2 * u], a0);
        a1 = fma(pv, s_hf[jj][2 * u + 1], a1);
      }
      Gf[(size_t)(2 * (j0 + jj)) * d.n + r] = a0;
      Gf[(size_t)(2 * (j0 + jj) + 1) * d.n + r] = a1;
    }
  }
}

// S(ra, cb) = sum_t H(ra, col_t) G(col_t, cb) + (ra == cb); grid (m tiles, m, frames)
__global__ void __launch_bounds__(128)
k_upd_S(const double* __restrict__ G, UpdDims d, const int32_t* __restrict__ idx, const int32_t* __restrict__ mf,
        const int32_t* __restrict__ type, const int32_t* __restrict__ pos, const double* __restrict__ Hcam,
        const double* __restrict__ Hfeat, double r_diag, double* __restrict__ S) {
  const int f = blockIdx.z, cb = blockIdx.y, m = mf[f];
  const int ra = blockIdx.x * 128 + threadIdx.x;
  if (cb >= m || ra >= m) return;
  const int i = idx[(size_t)f * d.F + (ra >> 1)], a = ra & 1;
  const size_t fi = (size_t)f * d.F + i;
  const double* hc = Hcam + fi * 26;
  const double* hf = Hfeat + fi * 12;
  const int ps = pos[fi], nz = 13 + (type[fi] == 0 ? 6 : 3);
  const double* g = G + ((size_t)f * d.Mmax + cb) * d.n;
  double acc = 0.0;
  for (int t = 0; t < nz; ++t) acc = fma(h_val(t, a, hc, hf), g[h_col(t, ps)], acc);
  S[((size_t)f * d.Mmax + cb) * d.Mmax + ra] = acc + (ra == cb ? r_diag : 0.0);
}

// In-place BLOCKED Gauss-Jordan inversion of S = H P H' + r I (symmetric positive definite: no pivoting needed), block
// size 32.  Per block step kb the matrix is read and written once (the scalar algorithm did that once per COLUMN:
// 11.4 ms per 64 frames at m = 323, bound by the L2 bandwidth of the one SM that owned the frame; blocked: see DESIGN.md):
//   k_inv_diag    one block per frame: Dinv = inv(A_kk) in shared memory
//   k_inv_panel   64-wide slabs over all frames: Cc = A(:, kb) saved; row panel A_kj <- Dinv A_kj (also kept in Rp);
//                 A_kk <- Dinv
//   k_inv_update  64 x 64 tiles over all frames: A_ij <- A_ij - Cc_i Rp_j (i, j outside kb);  A_ik <- -Cc_i Dinv
// A: m x m column-major, ld = Mmax.  Side buffers per frame: Di 32 x 32, Cc Mmax x 32, Rp 32 x Mmax.
constexpr int IB = 32;
// step 1: one block per frame, the 32 x 32 pivot block only
__global__ void __launch_bounds__(1024)
k_inv_diag(double* __restrict__ Sinv, int Mmax, const int32_t* __restrict__ mf, int k0, double* __restrict__ Di) {
  __shared__ double D[IB][IB + 1];
  const int f = blockIdx.x, tid = threadIdx.x;
  const int m = mf[f];
  if (k0 >= m) return;
  const int bs = min(IB, m - k0);
  double* A = Sinv + (size_t)f * Mmax * Mmax;
  const int ti = tid & 31, tj = tid >> 5;  // element (ti, tj) of the diagonal block
  D[ti][tj] = (ti < bs && tj < bs) ? A[(size_t)(k0 + tj) * Mmax + k0 + ti] : (ti == tj ? 1.0 : 0.0);
  __syncthreads();
  for (int k = 0; k < IB; ++k) {  // unpivoted in-place Gauss-Jordan (identity padding beyond bs)
    const double dinv = 1.0 / D[k][k];
    const double rkj = D[k][tj], cik = D[ti][k], old = D[ti][tj];
    __syncthreads();
    double v;
    if (ti == k) v = tj == k ? dinv : rkj * dinv;
    else if (tj == k) v = -cik * dinv;
    else v = fma(-cik, rkj * dinv, old);
    D[ti][tj] = v;
    __syncthreads();
  }
  Di[(size_t)f * IB * IB + tj * IB + ti] = D[ti][tj];
}

// step 2: grid (64-wide slabs, frames): slab t saves Cc(i, :) = A(i, kb) for its rows i and forms the row panel
// Rp(:, j) = Dinv A(kb, j) for its columns j (written back to A as well: only this block touches A(kb, j) here);
// the slab that holds the pivot block writes A_kk = Dinv last.
__global__ void __launch_bounds__(256)
k_inv_panel(double* __restrict__ Sinv, int Mmax, const int32_t* __restrict__ mf, int k0, const double* __restrict__ Di,
            double* __restrict__ Cc, double* __restrict__ Rp) {
  __shared__ double Dv[IB][IB + 1];  // Dv[r][c] = Dinv(r, c)
  __shared__ double Ak[IB][64 + 1];  // Ak[c][jj] = A(k0 + c, j0 + jj), old
  const int f = blockIdx.y, tid = threadIdx.x;
  const int m = mf[f];
  if (k0 >= m) return;
  const int bs = min(IB, m - k0);
  const int t0 = blockIdx.x * 64;
  if (t0 >= m) return;
  double* A = Sinv + (size_t)f * Mmax * Mmax;
  const double* di = Di + (size_t)f * IB * IB;
  double* cc = Cc + (size_t)f * Mmax * IB;
  double* rp = Rp + (size_t)f * IB * Mmax;
  for (int e = tid; e < IB * IB; e += 256) Dv[e & 31][e >> 5] = di[e];
  for (int e = tid; e < IB * 64; e += 256) {  // column panel rows t0 .. t0+63, coalesced along rows
    const int ii = e & 63, c = e >> 6;
    if (c < bs && t0 + ii < m) cc[(size_t)c * Mmax + t0 + ii] = A[(size_t)(k0 + c) * Mmax + t0 + ii];
  }
  for (int e = tid; e < IB * 64; e += 256) {  // old row panel columns t0 .. t0+63, contiguous along c
    const int c = e & 31, jj = e >> 5;
    Ak[c][jj] = (c < bs && t0 + jj < m) ? A[(size_t)(t0 + jj) * Mmax + k0 + c] : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < IB * 64; e += 256) {
    const int r = e & 31, jj = e >> 5;
    const int j = t0 + jj;
    if (r >= bs || j >= m) continue;
    double v;
    if (j >= k0 && j < k0 + bs) {
      v = Dv[r][j - k0];  // A_kk <- Dinv
    } else {
      double acc = 0.0;
#pragma unroll 8
      for (int c = 0; c < IB; ++c) acc = fma(Dv[r][c], Ak[c][jj], acc);
      v = acc;
      rp[(size_t)j * IB + r] = acc;
    }
    A[(size_t)j * Mmax + k0 + r] = v;
  }
}

__global__ void __launch_bounds__(256)
k_inv_update(double* __restrict__ Sinv, int Mmax, const int32_t* __restrict__ mf, int k0,
             const double* __restrict__ Di, const double* __restrict__ Cc, const double* __restrict__ Rp) {
  __shared__ double Cs[IB][64];      // Cs[c][i]
  __shared__ double Bs[IB][64 + 1];  // Bs[c][j]: Rp(c, j), or Dinv(c, j - k0) inside the pivot block's columns
  const int f = blockIdx.z;
  const int m = mf[f];
  if (k0 >= m) return;
  const int bs = min(IB, m - k0);
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  if (i0 >= m || j0 >= m) return;
  double* A = Sinv + (size_t)f * Mmax * Mmax;
  const double* di = Di + (size_t)f * IB * IB;
  const double* cc = Cc + (size_t)f * Mmax * IB;
  const double* rp = Rp + (size_t)f * IB * Mmax;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int e = tid; e < IB * 64; e += 256) {
    const int ii = e & 63, c = e >> 6;
    Cs[c][ii] = (c < bs && i0 + ii < m) ? cc[(size_t)c * Mmax + i0 + ii] : 0.0;
  }
  for (int e = tid; e < IB * 64; e += 256) {
    const int c = e & (IB - 1), jj = e >> 5;
    const int j = j0 + jj;
    double v = 0.0;
    if (c < bs && j < m) v = (j >= k0 && j < k0 + bs) ? di[(j - k0) * IB + c] : rp[(size_t)j * IB + c];
    Bs[c][jj] = v;
  }
  __syncthreads();
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
  for (int c = 0; c < IB; ++c) {
    double av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) av[a] = Cs[c][tx + 16 * a];
#pragma unroll
    for (int b = 0; b < 4; ++b) bv[b] = Bs[c][ty + 16 * b];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int j = j0 + ty + 16 * b;
    if (j >= m) continue;
    const bool jpiv = j >= k0 && j < k0 + bs;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + tx + 16 * a;
      if (i >= m || (i >= k0 && i < k0 + bs)) continue;  // the pivot block's rows are final (k_inv_panel)
      double* p = A + (size_t)j * Mmax + i;
      *p = jpiv ? -acc[a][b] : *p - acc[a][b];
    }
  }
}

// ---- fp64 GEMM, column-major:  C = [D +] A B  (TRANSB 0)  or  C = [D +] A B'  (TRANSB 1) ------------
//   A: M x Kd (lda), B: Kd x N (ldb) or N x Kd (ldb) when transposed; Kd = mf[f] per frame (and N = mf[f] when
//   n_from_m), or kd_fixed when mf is null.
// 64 x 64 tile per 256-thread block, 16-deep k slices through shared memory, 4 x 4 outputs per thread with rows
// tx + 16 i (conflict-free 128-byte smem reads, coalesced stores) and columns ty + 16 j (broadcast reads).
constexpr int GM = 64, GN = 64, GK = 16;
template <int TRANSB>
__global__ void __launch_bounds__(256)
k_dgemm(const double* __restrict__ A, int lda, size_t strideA, const double* __restrict__ B, int ldb, size_t strideB,
        double* __restrict__ C, int ldc, size_t strideC, const double* __restrict__ D, int ldd, size_t strideD, int M,
        int N, const int32_t* __restrict__ mf, int n_from_m, int kd_fixed) {
  __shared__ double As[GK][GM];
  __shared__ double Bs[GK][GN + 1];
  const int f = blockIdx.z;
  const int Kd = mf ? mf[f] : kd_fixed;
  if (n_from_m) N = Kd;
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  if (m0 >= M || n0 >= N) return;
  A += (size_t)f * strideA;
  B += (size_t)f * strideB;
  C += (size_t)f * strideC;
  if (D) D += (size_t)f * strideD;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < Kd; k0 += GK) {
    // A tile: 64 rows x 16 k, contiguous along rows
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int mm = tid & 63, kk = (tid >> 6) + 4 * it;
      const int gr = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gr < M && gk < Kd) ? A[(size_t)gk * lda + gr] : 0.0;
    }
    if (TRANSB) {  // B(k, j) = Bp[k * ldb + j]: contiguous along j
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int jj = tid & 63, kk = (tid >> 6) + 4 * it;
        const int gj = n0 + jj, gk = k0 + kk;
        Bs[kk][jj] = (gj < N && gk < Kd) ? B[(size_t)gk * ldb + gj] : 0.0;
      }
    } else {  // B(k, j) = Bp[j * ldb + k]: contiguous along k
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int kk = tid & 15, jj = (tid >> 4) + 16 * it;
        const int gj = n0 + jj, gk = k0 + kk;
        Bs[kk][jj] = (gj < N && gk < Kd) ? B[(size_t)gj * ldb + gk] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gc = n0 + ty + 16 * j;
    if (gc >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = m0 + tx + 16 * i;
      if (gr >= M) continue;
      const double v = acc[i][j];
      C[(size_t)gc * ldc + gr] = D ? D[(size_t)gc * ldd + gr] + v : v;
    }
  }
}

// P_out = 0.5 (P + P') - G K'   (update.m:37-38 with K S = P H' = G).  G K' is symmetric (= K S K'), so only the tiles
// on and below the diagonal are computed; the symmetrisation of :38 is the epilogue (the mirrored tile is staged through
// shared memory so that both writes are coalesced).  m == 0: P_out = P (update.m:52-53).
__global__ void __launch_bounds__(256)
k_dsyrk_update(const double* __restrict__ G, const double* __restrict__ K, size_t strideGK, const double* __restrict__ Pin,
               double* __restrict__ Pout, int n, const int32_t* __restrict__ mf) {
  __shared__ double smem[64 * 65];
  double(*As)[GM] = reinterpret_cast<double(*)[GM]>(smem);            // [GK][GM]
  double(*Bs)[GN] = reinterpret_cast<double(*)[GN]>(smem + GK * GM);  // [GK][GN]
  double(*Tt)[65] = reinterpret_cast<double(*)[65]>(smem);            // epilogue: [64][65]
  const int f = blockIdx.z;
  const int Kd = mf[f];
  const int i0 = blockIdx.x * GM, j0 = blockIdx.y * GN;
  if (j0 > i0 || i0 >= n) return;
  G += (size_t)f * strideGK;
  K += (size_t)f * strideGK;
  Pin += (size_t)f * n * n;
  Pout += (size_t)f * n * n;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool diag = i0 == j0;
  if (Kd == 0) {  // copy through, both triangles
    for (int e = tid; e < 64 * 64; e += 256) {
      const int ii = e & 63, jj = e >> 6;
      const int i = i0 + ii, j = j0 + jj;
      if (i < n && j < n) {
        Pout[(size_t)j * n + i] = Pin[(size_t)j * n + i];
        if (!diag) Pout[(size_t)i * n + j] = Pin[(size_t)i * n + j];
      }
    }
    return;
  }
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < Kd; k0 += GK) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int mm = tid & 63, kk = (tid >> 6) + 4 * it;
      const int gk = k0 + kk;
      As[kk][mm] = (i0 + mm < n && gk < Kd) ? G[(size_t)gk * n + i0 + mm] : 0.0;
      Bs[kk][mm] = (j0 + mm < n && gk < Kd) ? K[(size_t)gk * n + j0 + mm] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[kk][tx + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][ty + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  // mirrored input tile: Tt[ii][jj] = P(j0 + jj, i0 + ii)
  for (int e = tid; e < 64 * 64; e += 256) {
    const int jj = e & 63, ii = e >> 6;
    Tt[ii][jj] = (i0 + ii < n && j0 + jj < n) ? Pin[(size_t)(i0 + ii) * n + j0 + jj] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int jj = ty + 16 * b, j = j0 + jj;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int ii = tx + 16 * a, i = i0 + ii;
      if (i >= n || j >= n || (diag && ii < jj)) continue;
      const double v = (0.5 * Pin[(size_t)j * n + i] + 0.5 * Tt[ii][jj]) - acc[a][b];
      Pout[(size_t)j * n + i] = v;
      Tt[ii][jj] = v;
    }
  }
  __syncthreads();
  for (int e = tid; e < 64 * 64; e += 256) {  // mirror: P_out(j, i) = P_out(i, j)
    const int jj = e & 63, ii = e >> 6;
    if (i0 + ii >= n || j0 + jj >= n || (diag && ii <= jj)) continue;
    Pout[(size_t)(i0 + ii) * n + j0 + jj] = Tt[ii][jj];
  }
}

// x_out = x + K (z - h) over the flagged features (m == 0: x_out = x)
__global__ void __launch_bounds__(256)
k_upd_x(const double* __restrict__ x, const double* __restrict__ K, UpdDims d, const int32_t* __restrict__ idx,
        const int32_t* __restrict__ mf, const double* __restrict__ z, const double* __restrict__ h,
        double* __restrict__ x_out) {
  extern __shared__ double s_d[];  // z - h, m entries
  const int f = blockIdx.y, m = mf[f];
  for (int j = threadIdx.x; j < m; j += 256) {
    const size_t fi = (size_t)f * d.F + idx[(size_t)f * d.F + (j >> 1)];
    s_d[j] = z[fi * 2 + (j & 1)] - h[fi * 2 + (j & 1)];
  }
  __syncthreads();
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= d.n) return;
  const double* Kf = K + (size_t)f * d.Mmax * d.n;
  double acc = 0.0;
  for (int j = 0; j < m; ++j) acc = fma(Kf[(size_t)j * d.n + r], s_d[j], acc);
  x_out[(size_t)f * d.n + r] = x[(size_t)f * d.n + r] + acc;
}

// rows / columns 4:7 (0-based 3..6) through Jnorm = normJac(x_k_k(4:7)) (update.m:42-46), then q / |q| (:48)
__global__ void __launch_bounds__(256)
k_upd_quat(double* __restrict__ P, double* __restrict__ x, int n, const int32_t* __restrict__ mf) {
  __shared__ double J[16];
  __shared__ double q44[16], tmp[16];
  const int f = blockIdx.x;
  if (mf[f] == 0) return;
  double* Pf = P + (size_t)f * n * n;
  double* xf = x + (size_t)f * n;
  const int tid = threadIdx.x;
  if (tid == 0) {
    const double r = xf[3], a = xf[4], b = xf[5], c = xf[6];
    const double s2 = ((r * r + a * a) + b * b) + c * c;
    const double sc = 1.0 / (s2 * sqrt(s2));  // ^(-3/2)
    const double M[16] = {a * a + b * b + c * c, -r * a, -r * b, -r * c,   // row 0
                          -a * r, r * r + b * b + c * c, -a * b, -a * c,   // row 1
                          -b * r, -b * a, r * r + a * a + c * c, -b * c,   // row 2
                          -c * r, -c * a, -c * b, r * r + a * a + b * b};  // row 3
    for (int i = 0; i < 16; ++i) J[i] = sc * M[i];  // J[4 * row + col]
  }
  if (tid < 16) q44[tid] = Pf[(size_t)(3 + (tid >> 2)) * n + 3 + (tid & 3)];  // q44[4 * col + row]
  __syncthreads();
  // columns 3..6 for the rows outside the quaternion: p(r, 4:7) * Jnorm'  and the mirrored rows Jnorm * p(4:7, c)
  for (int r = tid; r < n; r += 256) {
    if (r >= 3 && r < 7) continue;
    double v[4], o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) v[b] = Pf[(size_t)(3 + b) * n + r];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double acc = 0.0;
#pragma unroll
      for (int b = 0; b < 4; ++b) acc = fma(v[b], J[4 * a + b], acc);
      o[a] = acc;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      Pf[(size_t)(3 + a) * n + r] = o[a];  // column block
      Pf[(size_t)r * n + 3 + a] = o[a];    // row block (P is symmetric after k_upd_sym)
    }
  }
  // 4 x 4 block: (Jnorm * p44) * Jnorm'
  if (tid < 16) {
    const int a = tid >> 2, c = tid & 3;  // tmp(a, c) = sum_b J(a, b) p44(b, c)
    double acc = 0.0;
    for (int b = 0; b < 4; ++b) acc = fma(J[4 * a + b], q44[4 * c + b], acc);
    tmp[4 * a + c] = acc;
  }
  __syncthreads();
  if (tid < 16) {
    const int a = tid >> 2, c = tid & 3;  // out(a, c) = sum_b tmp(a, b) J(c, b)
    double acc = 0.0;
    for (int b = 0; b < 4; ++b) acc = fma(tmp[4 * a + b], J[4 * c + b], acc);
    Pf[(size_t)(3 + c) * n + 3 + a] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    const double nq = sqrt(((xf[3] * xf[3] + xf[4] * xf[4]) + xf[5] * xf[5]) + xf[6] * xf[6]);
    for (int i = 3; i < 7; ++i) xf[i] = xf[i] / nq;
  }
}

// rescue_hi_inliers.m:35-46: one thread per feature
__global__ void __launch_bounds__(128)
k_upd_rescue(const double* __restrict__ P, int n, int F, const int32_t* __restrict__ type,
             const int32_t* __restrict__ pos, const uint8_t* __restrict__ ic, const uint8_t* __restrict__ li,
             const double* __restrict__ z, const double* __restrict__ h, const double* __restrict__ Hcam,
             const double* __restrict__ Hfeat, double chi2, uint8_t* __restrict__ hi) {
  const int f = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
  if (i >= F) return;
  const size_t fi = (size_t)f * F + i;
  if (!(ic[fi] == 1 && li[fi] == 0)) return;  // the reference leaves high_innovation_inlier untouched
  const double* Pf = P + (size_t)f * n * n;
  const double* hc = Hcam + fi * 26;
  const double* hf = Hfeat + fi * 12;
  const int ps = pos[fi], nz = 13 + (type[fi] == 0 ? 6 : 3);
  double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
  for (int t = 0; t < nz; ++t) {  // (H P)(a, col_t) then times H'
    double g0 = 0.0, g1 = 0.0;
    const int ct = h_col(t, ps);
    for (int u = 0; u < nz; ++u) {
      const double p = Pf[(size_t)ct * n + h_col(u, ps)];
      g0 = fma(h_val(u, 0, hc, hf), p, g0);
      g1 = fma(h_val(u, 1, hc, hf), p, g1);
    }
    const double h0 = h_val(t, 0, hc, hf), h1 = h_val(t, 1, hc, hf);
    s00 = fma(g0, h0, s00);
    s01 = fma(g0, h1, s01);
    s10 = fma(g1, h0, s10);
    s11 = fma(g1, h1, s11);
  }
  const double det = s00 * s11 - s01 * s10;
  const double n0 = z[fi * 2] - h[fi * 2], n1 = z[fi * 2 + 1] - h[fi * 2 + 1];
  // nu' inv(S) nu with inv(S) = [s11 -s01; -s10 s00] / det
  const double q = (n0 * (s11 * n0 - s01 * n1) + n1 * (s00 * n1 - s10 * n0)) / det;
  hi[fi] = q < chi2 ? 1 : 0;
}

#define UPD_LIVE()                                                                                         \
  do {                                                                                                     \
    if (!ctx) return PRE3_ERR_ARG;                                                                         \
    if (ctx->device < 0) return fail(ctx, PRE3_ERR_CUDA, "no CUDA device (libpre3 has no CPU fallback)"); \
    PRE3_CUDA(cudaSetDevice(ctx->device));                                                                 \
  } while (0)

size_t upd_ws_per_frame(int n, int F) {
  const size_t M = 2 * (size_t)F;
  return 2 * align_up(8 * (size_t)n * M) + 2 * align_up(8 * M * M) + align_up(4 * (size_t)F) +
         align_up(8 * IB * IB) + 2 * align_up(8 * M * IB) + 1024;
}

int frames_per_chunk(int Fr, int n, int F) {
  const size_t budget = (size_t)2 << 30;
  const size_t per = upd_ws_per_frame(n, F);
  size_t c = std::max<size_t>(1, std::min<size_t>((size_t)Fr, budget / per));
  if (const char* e = getenv("PRE3_UPD_CHUNK")) c = std::max<size_t>(1, std::min<size_t>(c, (size_t)atoi(e)));  // tests
  return (int)c;
}

int update_chunk(pre3_ctx* ctx, int C, int n, int F, const double* dx, const double* dP, const int32_t* dtype,
                 const int32_t* dpos, const uint8_t* dsel, const double* dz, const double* dh, const double* dHcam,
                 const double* dHfeat, double r_diag, double* dx_out, double* dP_out, int32_t* dm_out) {
  Span span__(ctx, T_EKF_UPDATE);
  if (F == 0) {  // nothing can be flagged: every frame is copied through (update.m:50-54)
    PRE3_CUDA(cudaMemcpyAsync(dx_out, dx, 8 * (size_t)C * n, cudaMemcpyDeviceToDevice, ctx->stream));
    PRE3_CUDA(cudaMemcpyAsync(dP_out, dP, 8 * (size_t)C * n * n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (dm_out) PRE3_CUDA(cudaMemsetAsync(dm_out, 0, 4 * (size_t)C, ctx->stream));
    return PRE3_OK;
  }
  const UpdDims d{n, F, 2 * F};
  const size_t M = (size_t)d.Mmax;
  double* G = ws_take<double>(ctx, (size_t)C * n * M);
  double* K = ws_take<double>(ctx, (size_t)C * n * M);
  double* S = ws_take<double>(ctx, (size_t)C * M * M);
  double* Si = ws_take<double>(ctx, (size_t)C * M * M);
  int32_t* idx = ws_take<int32_t>(ctx, (size_t)C * F);
  double* Di = ws_take<double>(ctx, (size_t)C * IB * IB);
  double* Cc = ws_take<double>(ctx, (size_t)C * M * IB);
  double* Rp = ws_take<double>(ctx, (size_t)C * M * IB);
  int32_t* mf = dm_out ? dm_out : ws_take<int32_t>(ctx, C);
  cudaStream_t st = ctx->stream;
  k_upd_index<<<C, 256, 0, st>>>(dsel, F, idx, mf);
  const int gx = (n + 255) / 256;
  const int gz = std::max(1, std::min(16, (4 * ctx->sm_count + gx * C - 1) / (gx * C)));
  k_upd_G<<<dim3(gx, C, gz), 256, 0, st>>>(dP, d, idx, mf, dtype, dpos, dHcam, dHfeat, G);
  k_upd_S<<<dim3((d.Mmax + 127) / 128, d.Mmax, C), 128, 0, st>>>(G, d, idx, mf, dtype, dpos, dHcam, dHfeat, r_diag, S);
  PRE3_CUDA(cudaMemcpyAsync(Si, S, 8 * (size_t)C * M * M, cudaMemcpyDeviceToDevice, st));
  for (int k0 = 0; k0 < d.Mmax; k0 += IB) {  // frames whose m <= k0 return at once
    k_inv_diag<<<C, 1024, 0, st>>>(Si, d.Mmax, mf, k0, Di);
    k_inv_panel<<<dim3((d.Mmax + 63) / 64, C), 256, 0, st>>>(Si, d.Mmax, mf, k0, Di, Cc, Rp);
    k_inv_update<<<dim3((d.Mmax + 63) / 64, (d.Mmax + 63) / 64, C), 256, 0, st>>>(Si, d.Mmax, mf, k0, Di, Cc, Rp);
  }
  const dim3 g1((n + GM - 1) / GM, (d.Mmax + GN - 1) / GN, C), g2((n + GM - 1) / GM, (n + GN - 1) / GN, C);
  // K = G inv(S);  P_out = 0.5 (P + P') - G K'   (K S K' = G K': K S = P H' = G by construction, so the reference's
  // second n x m x m product is not repeated; the difference is rounding x cond(S), far inside the tolerance)
  k_dgemm<0><<<g1, 256, 0, st>>>(G, n, (size_t)n * M, Si, d.Mmax, M * M, K, n, (size_t)n * M, nullptr, 0, 0, n, 0, mf, 1, 0);
  k_dsyrk_update<<<g2, 256, 0, st>>>(G, K, (size_t)n * M, dP, dP_out, n, mf);
  k_upd_x<<<dim3((n + 255) / 256, C), 256, M * sizeof(double), st>>>(dx, K, d, idx, mf, dz, dh, dx_out);
  k_upd_quat<<<C, 256, 0, st>>>(dP_out, dx_out, n, mf);
  count_launch(ctx, 7 + 3 * ((d.Mmax + IB - 1) / IB));
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int update_impl(pre3_ctx* ctx, int Fr, int n, int F, const double* dx, const double* dP, const int32_t* dtype,
                const int32_t* dpos, const uint8_t* dsel, const double* dz, const double* dh, const double* dHcam,
                const double* dHfeat, double r_diag, double* dx_out, double* dP_out, int32_t* dm_out) {
  const int C = frames_per_chunk(Fr, n, F);
  for (int f0 = 0; f0 < Fr; f0 += C) {
    const int c = std::min(C, Fr - f0);
    ctx->ws_off = 0;  // the chunks reuse the arena in stream order
    const size_t fF = (size_t)f0 * F;
    PRE3_TRY(update_chunk(ctx, c, n, F, dx + (size_t)f0 * n, dP + (size_t)f0 * n * n, dtype + fF, dpos + fF, dsel + fF,
                          dz + 2 * fF, dh + 2 * fF, dHcam + 26 * fF, dHfeat + 12 * fF, r_diag, dx_out + (size_t)f0 * n,
                          dP_out + (size_t)f0 * n * n, dm_out ? dm_out + f0 : nullptr));
  }
  return PRE3_OK;
}

// ---- update.m with H and R given as dense matrices (the function's own signature) ------------------
__global__ void k_set_int(int32_t* p, int v) { *p = v; }

__global__ void __launch_bounds__(256)
k_upd_x_dense(const double* __restrict__ x, const double* __restrict__ K, int n, int m, const double* __restrict__ z,
              const double* __restrict__ h, double* __restrict__ x_out) {
  extern __shared__ double s_d[];
  for (int j = threadIdx.x; j < m; j += 256) s_d[j] = z[j] - h[j];
  __syncthreads();
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  double acc = 0.0;
  for (int j = 0; j < m; ++j) acc = fma(K[(size_t)j * n + r], s_d[j], acc);
  x_out[r] = x[r] + acc;
}

size_t dense_ws_bytes(int n, int m) {
  const size_t M = (size_t)std::max(m, 1);
  return 2 * align_up(8 * (size_t)n * M) + 2 * align_up(8 * M * M) + align_up(8 * IB * IB) + 2 * align_up(8 * M * IB) + 4096;
}

int update_dense_impl(pre3_ctx* ctx, int n, int m, const double* dx, const double* dP, const double* dH,
                      const double* dR, const double* dz, const double* dh, double* dx_out, double* dP_out,
                      double* dK_out) {
  Span span__(ctx, T_EKF_UPDATE);
  cudaStream_t st = ctx->stream;
  if (m == 0) {  // update.m:50-54
    PRE3_CUDA(cudaMemcpyAsync(dx_out, dx, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    PRE3_CUDA(cudaMemcpyAsync(dP_out, dP, 8 * (size_t)n * n, cudaMemcpyDeviceToDevice, st));
    return PRE3_OK;
  }
  const size_t M = (size_t)m;
  double* G = ws_take<double>(ctx, (size_t)n * M);
  double* K = dK_out ? dK_out : ws_take<double>(ctx, (size_t)n * M);
  double* S = ws_take<double>(ctx, M * M);
  double* Si = ws_take<double>(ctx, M * M);
  double* Di = ws_take<double>(ctx, IB * IB);
  double* Cc = ws_take<double>(ctx, M * IB);
  double* Rp = ws_take<double>(ctx, M * IB);
  int32_t* mf = ws_take<int32_t>(ctx, 1);
  k_set_int<<<1, 1, 0, st>>>(mf, m);
  const dim3 gnm((n + GM - 1) / GM, (m + GN - 1) / GN, 1), gmm((m + GM - 1) / GM, (m + GN - 1) / GN, 1);
  // G = P H'  (n x n x m);  S = H G + R  (m x n x m)
  k_dgemm<1><<<gnm, 256, 0, st>>>(dP, n, 0, dH, m, 0, G, n, 0, nullptr, 0, 0, n, m, nullptr, 0, n);
  k_dgemm<0><<<gmm, 256, 0, st>>>(dH, m, 0, G, n, 0, S, m, 0, dR, m, 0, m, m, nullptr, 0, n);
  PRE3_CUDA(cudaMemcpyAsync(Si, S, 8 * M * M, cudaMemcpyDeviceToDevice, st));
  for (int k0 = 0; k0 < m; k0 += IB) {
    k_inv_diag<<<1, 1024, 0, st>>>(Si, m, mf, k0, Di);
    k_inv_panel<<<dim3((m + 63) / 64, 1), 256, 0, st>>>(Si, m, mf, k0, Di, Cc, Rp);
    k_inv_update<<<dim3((m + 63) / 64, (m + 63) / 64, 1), 256, 0, st>>>(Si, m, mf, k0, Di, Cc, Rp);
  }
  k_dgemm<0><<<gnm, 256, 0, st>>>(G, n, 0, Si, m, 0, K, n, 0, nullptr, 0, 0, n, m, nullptr, 0, m);
  k_dsyrk_update<<<dim3((n + GM - 1) / GM, (n + GN - 1) / GN, 1), 256, 0, st>>>(G, K, 0, dP, dP_out, n, mf);
  k_upd_x_dense<<<(n + 255) / 256, 256, M * sizeof(double), st>>>(dx, K, n, m, dz, dh, dx_out);
  k_upd_quat<<<1, 256, 0, st>>>(dP_out, dx_out, n, mf);
  count_launch(ctx, 7 + 3 * ((m + IB - 1) / IB));
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int check_upd(pre3_ctx* ctx, int Fr, int n, int F) {
  if (Fr < 0 || n < 13 || F < 0) return fail(ctx, PRE3_ERR_ARG, "bad sizes (n >= 13: the camera states)");
  return PRE3_OK;
}

}  // namespace
}  // namespace pre3

using namespace pre3;

extern "C" {

int pre3_ekf_update_batch_dev(pre3_ctx* ctx, int Fr, int n, int F, const double* dx, const double* dP,
                              const int32_t* dtype, const int32_t* dpos, const uint8_t* dsel, const double* dz,
                              const double* dh, const double* dHcam, const double* dHfeat, double r_diag,
                              double* dx_out, double* dP_out, int32_t* dm_out) {
  UPD_LIVE();
  PRE3_TRY(check_upd(ctx, Fr, n, F));
  if (Fr == 0) return PRE3_OK;
  if (!dx || !dP || !dsel || !dx_out || !dP_out || (F > 0 && (!dtype || !dpos || !dz || !dh || !dHcam || !dHfeat)))
    return fail(ctx, PRE3_ERR_ARG, "null pointer");
  if (dP_out == dP) return fail(ctx, PRE3_ERR_ARG, "p_k_k must not alias the input covariance");
  const int C = frames_per_chunk(Fr, n, F);
  PRE3_TRY(ws_reserve(ctx, (size_t)C * upd_ws_per_frame(n, F) + 8192));
  return update_impl(ctx, Fr, n, F, dx, dP, dtype, dpos, dsel, dz, dh, dHcam, dHfeat, r_diag, dx_out, dP_out, dm_out);
}

int pre3_ekf_update_batch(pre3_ctx* ctx, int Fr, int n, int F, const double* x, const double* P, const int32_t* type,
                          const int32_t* pos, const uint8_t* sel, const double* z, const double* h, const double* Hcam,
                          const double* Hfeat, double r_diag, double* x_out, double* P_out, int32_t* m_out) {
  UPD_LIVE();
  PRE3_TRY(check_upd(ctx, Fr, n, F));
  if (Fr == 0) return PRE3_OK;
  if (!x || !P || !sel || !x_out || !P_out || (F > 0 && (!type || !pos || !z || !h || !Hcam || !Hfeat)))
    return fail(ctx, PRE3_ERR_ARG, "null pointer");
  // staged through cudaMalloc'd buffers (a covariance batch can exceed the arena's sensible size); one sync at the end
  const size_t xb = 8 * (size_t)Fr * n, pb = 8 * (size_t)Fr * n * n, fF = (size_t)Fr * F;
  char* buf = nullptr;
  const size_t total = 2 * align_up(xb) + 2 * align_up(pb) + 2 * align_up(4 * fF) + align_up(fF) + 2 * align_up(16 * fF) +
                       align_up(208 * fF) + align_up(96 * fF) + align_up(4 * (size_t)Fr) + 4096;
  cudaError_t e = cudaMalloc((void**)&buf, total);
  if (e != cudaSuccess) return fail(ctx, PRE3_ERR_ALLOC, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = buf + off;
    off += align_up(bytes);
    return p;
  };
  double* dx = (double*)take(xb);
  double* dxo = (double*)take(xb);
  double* dP = (double*)take(pb);
  double* dPo = (double*)take(pb);
  int32_t* dty = (int32_t*)take(4 * fF);
  int32_t* dps = (int32_t*)take(4 * fF);
  uint8_t* dsl = (uint8_t*)take(fF);
  double* dzz = (double*)take(16 * fF);
  double* dhh = (double*)take(16 * fF);
  double* dHc = (double*)take(208 * fF);
  double* dHf = (double*)take(96 * fF);
  int32_t* dm = (int32_t*)take(4 * (size_t)Fr);
  cudaStream_t st = ctx->stream;
  int rc = PRE3_OK;
  auto up = [&](void* d, const void* hptr, size_t b) {
    if (b && rc == PRE3_OK && cudaMemcpyAsync(d, hptr, b, cudaMemcpyHostToDevice, st) != cudaSuccess)
      rc = fail(ctx, PRE3_ERR_CUDA, "cudaMemcpyAsync (host to device)");
  };
  up(dx, x, xb); up(dP, P, pb); up(dty, type, 4 * fF); up(dps, pos, 4 * fF); up(dsl, sel, fF);
  up(dzz, z, 16 * fF); up(dhh, h, 16 * fF); up(dHc, Hcam, 208 * fF); up(dHf, Hfeat, 96 * fF);
  if (rc == PRE3_OK)
    rc = pre3_ekf_update_batch_dev(ctx, Fr, n, F, dx, dP, dty, dps, dsl, dzz, dhh, dHc, dHf, r_diag, dxo, dPo, dm);
  if (rc == PRE3_OK) {
    cudaMemcpyAsync(x_out, dxo, xb, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(P_out, dPo, pb, cudaMemcpyDeviceToHost, st);
    if (m_out) cudaMemcpyAsync(m_out, dm, 4 * (size_t)Fr, cudaMemcpyDeviceToHost, st);
  }
  const cudaError_t se = cudaStreamSynchronize(st);
  cudaFree(buf);
  if (rc == PRE3_OK && se != cudaSuccess) return fail(ctx, PRE3_ERR_CUDA, cudaGetErrorString(se));
  return rc;
}

int pre3_ekf_update_dense_dev(pre3_ctx* ctx, int n, int m, const double* dx, const double* dP, const double* dH,
                              const double* dR, const double* dz, const double* dh, double* dx_out, double* dP_out,
                              double* dK_out) {
  UPD_LIVE();
  if (n < 7 || m < 0) return fail(ctx, PRE3_ERR_ARG, "bad sizes (n >= 7: the quaternion lives in x(4:7))");
  if (!dx || !dP || !dx_out || !dP_out || (m > 0 && (!dH || !dR || !dz || !dh))) return fail(ctx, PRE3_ERR_ARG, "null pointer");
  if (dP_out == dP) return fail(ctx, PRE3_ERR_ARG, "p_k_k must not alias the input covariance");
  if ((size_t)m * sizeof(double) > 48 * 1024) return fail(ctx, PRE3_ERR_ARG, "more than 6144 stacked rows");
  PRE3_TRY(ws_reserve(ctx, dense_ws_bytes(n, m)));
  return update_dense_impl(ctx, n, m, dx, dP, dH, dR, dz, dh, dx_out, dP_out, dK_out);
}

int pre3_ekf_update_dense(pre3_ctx* ctx, int n, int m, const double* x, const double* P, const double* H,
                          const double* R, const double* z, const double* h, double* x_out, double* P_out,
                          double* K_out) {
  UPD_LIVE();
  if (n < 7 || m < 0) return fail(ctx, PRE3_ERR_ARG, "bad sizes (n >= 7: the quaternion lives in x(4:7))");
  if (!x || !P || !x_out || !P_out || (m > 0 && (!H || !R || !z || !h))) return fail(ctx, PRE3_ERR_ARG, "null pointer");
  if ((size_t)m * sizeof(double) > 48 * 1024) return fail(ctx, PRE3_ERR_ARG, "more than 6144 stacked rows");
  const size_t nb = 8 * (size_t)n, pb = 8 * (size_t)n * n, hb = 8 * (size_t)m * n, rb = 8 * (size_t)m * m, zb = 8 * (size_t)m;
  PRE3_TRY(ws_reserve(ctx, dense_ws_bytes(n, m) + 2 * align_up(nb) + 2 * align_up(pb) + 2 * align_up(hb) + align_up(rb) +
                               2 * align_up(zb) + 8192));
  double* dx = ws_take<double>(ctx, n);
  double* dxo = ws_take<double>(ctx, n);
  double* dP = ws_take<double>(ctx, (size_t)n * n);
  double* dPo = ws_take<double>(ctx, (size_t)n * n);
  double* dH = ws_take<double>(ctx, (size_t)std::max(m, 1) * n);
  double* dK = ws_take<double>(ctx, (size_t)std::max(m, 1) * n);
  double* dR = ws_take<double>(ctx, (size_t)std::max(m, 1) * std::max(m, 1));
  double* dz = ws_take<double>(ctx, std::max(m, 1));
  double* dh = ws_take<double>(ctx, std::max(m, 1));
  cudaStream_t st = ctx->stream;
  PRE3_CUDA(cudaMemcpyAsync(dx, x, nb, cudaMemcpyHostToDevice, st));
  PRE3_CUDA(cudaMemcpyAsync(dP, P, pb, cudaMemcpyHostToDevice, st));
  if (m > 0) {
    PRE3_CUDA(cudaMemcpyAsync(dH, H, hb, cudaMemcpyHostToDevice, st));
    PRE3_CUDA(cudaMemcpyAsync(dR, R, rb, cudaMemcpyHostToDevice, st));
    PRE3_CUDA(cudaMemcpyAsync(dz, z, zb, cudaMemcpyHostToDevice, st));
    PRE3_CUDA(cudaMemcpyAsync(dh, h, zb, cudaMemcpyHostToDevice, st));
  }
  PRE3_TRY(update_dense_impl(ctx, n, m, dx, dP, dH, dR, dz, dh, dxo, dPo, dK));
  PRE3_CUDA(cudaMemcpyAsync(x_out, dxo, nb, cudaMemcpyDeviceToHost, st));
  PRE3_CUDA(cudaMemcpyAsync(P_out, dPo, pb, cudaMemcpyDeviceToHost, st));
  if (K_out && m > 0) PRE3_CUDA(cudaMemcpyAsync(K_out, dK, hb, cudaMemcpyDeviceToHost, st));
  PRE3_CUDA(cudaStreamSynchronize(st));
  return PRE3_OK;
}

int pre3_ekf_rescue_hi_inliers_batch_dev(pre3_ctx* ctx, int Fr, int n, int F, const double* dP_kk,
                                         const int32_t* dtype, const int32_t* dpos, const uint8_t* dic,
                                         const uint8_t* dli, const double* dz, const double* dh, const double* dHcam,
                                         const double* dHfeat, uint8_t* dhi) {
  UPD_LIVE();
  PRE3_TRY(check_upd(ctx, Fr, n, F));
  if (Fr == 0 || F == 0) return PRE3_OK;
  if (!dP_kk || !dtype || !dpos || !dic || !dli || !dz || !dh || !dHcam || !dHfeat || !dhi)
    return fail(ctx, PRE3_ERR_ARG, "null pointer");
  Span span__(ctx, T_EKF_UPDATE);
  k_upd_rescue<<<dim3((F + 127) / 128, Fr), 128, 0, ctx->stream>>>(dP_kk, n, F, dtype, dpos, dic, dli, dz, dh, dHcam,
                                                                  dHfeat, 5.9915, dhi);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

}  // extern "C"
