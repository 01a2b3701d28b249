// ekf.cu -- BASELINE.json config 4 on sm_100a: the 1-point-RANSAC EKF hypothesis path.
//   ransac_hypotheses              M/ransac_hypotheses.m:27-85
//   compute_hypothesis_support_fast M/compute_hypothesis_support_fast.m:27-116
//   (+ m.m, q2r.m, distort_fm_my_version.m, generate_state_vector_pattern.m,
//      select_random_match.m, set_as_most_supported_hypothesis.m)
//
// Kernels
//   k_ekf_index   per frame: IC list, measured-feature list, m (select_random_match.m:47-51).
//   k_ekf_gain    G = P*H' for every individually compatible feature of the frame (n x 2 each):
//                 P is streamed once (HBM bound); every hypothesis then gathers its 2m columns.
//   k_ekf_hyp     one warp per (frame, hypothesis): S = Hi*P*Hi' + R from the P sub-blocks, inv(S)
//                 by a cooperative Gauss-Jordan, innovation, the updated camera states -> a 480-byte
//                 record per hypothesis.
//   k_ekf_score   one block per (frame, hypothesis), one thread per measured feature: the 6 (3)
//                 updated feature states xi = x + K (zi - hi) in the reference's K-form, projection,
//                 distortion, residual; block min (the `min(residuals) + threshold` rule, :70) and count.
//   k_ekf_stop / k_ekf_pick  the reference's sequential loop control (:40-46,:74-80) replayed
//                 over the supports (hypotheses are evaluated in waves; frames whose loop has
//                 ended skip the later waves), first-maximum selection.
//   k_ekf_score<FINAL> recomputes the winner and writes low_innovation_inlier
//                 (set_as_most_supported_hypothesis.m:32-53).
//   k_ekf_support_given  compute_hypothesis_support_fast for states given by the caller.
// fp64 throughout (FP64-pipe bound); arithmetic order = the SPEC comments of
// oracle/pre3_oracle_ekf.c, so supports, masks and selection equal the CPU checker's bit for bit.
// Compile with -fmad=false.
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace pre3 {
namespace ekf {

// ------------------------------------------------------------------------------------------
// sin / cos: SPEC in oracle/pre3_oracle_ekf.c (Cody-Waite reduction + classic minimax kernels)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double ksin(double x, double y) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = x * x, v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  if (y == 0.0) return x + v * (S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

__device__ __forceinline__ double kcos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = x * x;
  const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  const double hz = 0.5 * z, w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * r - x * y));
}

__device__ __forceinline__ void sincos_spec(double x, double& s, double& c) {
  if (!(fabs(x) < 1647099.0)) {
    s = c = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  if (fabs(x) <= 0.78539816339744830962) {
    s = ksin(x, 0.0);
    c = kcos(x, 0.0);
    return;
  }
  const double INVPIO2 = 6.36619772367581382433e-01, P1 = 1.57079632673412561417e+00,
               P2 = 6.07710050630396597660e-11, P2T = 2.02226624879595063154e-21;
  const long long n = __double2ll_rz(x * INVPIO2 + (x >= 0.0 ? 0.5 : -0.5));
  const double fn = (double)n;
  const double r = (x - fn * P1) - fn * P2;
  const double w = fn * P2T;
  const double y0 = r - w;
  const double y1 = (r - y0) - w;
  const double ks = ksin(y0, y1), kc = kcos(y0, y1);
  switch ((int)(n & 3)) {
    case 0: s = ks; c = kc; break;
    case 1: s = kc; c = -ks; break;
    case 2: s = -ks; c = -kc; break;
    default: s = -kc; c = ks; break;
  }
}

// q2r (M/q2r.m:29-36), row-major rotwc
__device__ __forceinline__ void q2r(const double* q, double* R) {
  const double r = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = ((r * r + x * x) - y * y) - z * z;
  R[1] = 2.0 * (x * y - r * z);
  R[2] = 2.0 * (z * x + r * y);
  R[3] = 2.0 * (x * y + r * z);
  R[4] = ((r * r - x * x) + y * y) - z * z;
  R[5] = 2.0 * (y * z - r * x);
  R[6] = 2.0 * (z * x - r * y);
  R[7] = 2.0 * (y * z + r * x);
  R[8] = ((r * r - x * x) - y * y) + z * z;
}

// rotcw*v, pinhole, radial distortion, residual (compute_hypothesis_support_fast.m:55-69,
// distort_fm_my_version.m:52-61)
__device__ __forceinline__ double project_residual(const double* Rw, const double* v, const pre3_cam& cam, double z0,
                                                   double z1) {
  double hc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) hc[i] = (Rw[i] * v[0] + Rw[3 + i] * v[1]) + Rw[6 + i] * v[2];
  const double u = cam.f * (hc[0] / hc[2]) + cam.Cx;
  const double w = cam.f * (hc[1] / hc[2]) + cam.Cy;
  const double xu = (u - cam.Cx) / cam.f, yu = (w - cam.Cy) / cam.f;
  const double ru = sqrt(xu * xu + yu * yu);
  const double ru2 = ru * ru;
  const double D = (1.0 + cam.k1 * ru2) + cam.k2 * (ru2 * ru2);
  const double ud = (xu * D) * cam.f + cam.Cx, vd = (yu * D) * cam.f + cam.Cy;
  const double n0 = z0 - ud, n1 = z1 - vd;
  return sqrt(n0 * n0 + n1 * n1);
}

// residual of one measured feature from its (updated) states s[0..nf) and the camera position
__device__ __forceinline__ double feature_residual(int type, const double* s, const double* rwc, const double* Rw,
                                                   const pre3_cam& cam, double z0, double z1) {
  double v[3];
  if (type == 0) {
    double st, ct, sp, cp;
    sincos_spec(s[3], st, ct);
    sincos_spec(s[4], sp, cp);
    const double mi[3] = {cp * st, -sp, cp * ct};  // M/m.m:32-34
    const double rho = s[5];
#pragma unroll
    for (int a = 0; a < 3; ++a) v[a] = (s[a] - rwc[a]) * rho + mi[a];
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) v[a] = s[a] - rwc[a];
  }
  return project_residual(Rw, v, cam, z0, z1);
}

// ------------------------------------------------------------------------------------------
// block reductions (blockDim.x a multiple of 32, <= 1024)
// ------------------------------------------------------------------------------------------
// min that skips NaN (MATLAB's min); NaN if every value is NaN
__device__ __forceinline__ double nanmin2(double a, double b) {
  if (a != a) return b;
  if (b != b) return a;
  return b < a ? b : a;
}

__device__ __forceinline__ double block_nanmin(double v, double* scratch) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = nanmin2(v, __shfl_xor_sync(0xffffffffu, v, off));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = scratch[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = nanmin2(r, scratch[w]);
  __syncthreads();
  return r;
}

__device__ __forceinline__ int block_sum_int(int v, int* scratch) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += scratch[w];
  __syncthreads();
  return r;
}

// ------------------------------------------------------------------------------------------
// k_ekf_index: one warp per frame
// ------------------------------------------------------------------------------------------
struct FrameIdx {
  int32_t status;  // 0 ok, 1 no IC match, 3 IC feature without z
  int32_t num_ic;
  int32_t n_meas;
  int32_t m;
};

__global__ void __launch_bounds__(128)
k_ekf_index(int Fr, int F, const uint8_t* __restrict__ has_z, const uint8_t* __restrict__ ic, FrameIdx* __restrict__ idx,
            int32_t* __restrict__ ic_list, int32_t* __restrict__ meas) {
  const int frame = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (frame >= Fr) return;
  int nic = 0, nme = 0;
  bool bad = false;
  for (int base = 0; base < F; base += 32) {
    const int i = base + lane;
    const bool c = i < F && ic[(size_t)frame * F + i] != 0;
    const bool z = i < F && has_z[(size_t)frame * F + i] != 0;
    bad = bad || (c && !z);
    const unsigned bc = __ballot_sync(0xffffffffu, c), bz = __ballot_sync(0xffffffffu, z);
    const unsigned lt = (1u << lane) - 1u;
    if (c) ic_list[(size_t)frame * F + nic + __popc(bc & lt)] = i;
    if (z) meas[(size_t)frame * F + nme + __popc(bz & lt)] = i;
    nic += __popc(bc);
    nme += __popc(bz);
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    FrameIdx o;
    o.status = nic == 0 ? 1 : (bad ? 3 : 0);
    o.num_ic = nic;
    o.n_meas = nme;
    o.m = nic > 3 ? 3 : 1;
    idx[frame] = o;
  }
}

// ------------------------------------------------------------------------------------------
// k_ekf_gain: G[f][c][e] = sum_t P(e, k_t) * H_f(c, k_t) over the structural non-zeros of H_f in
// ascending column order (P*Hi', ransac_hypotheses.m:62).  P column-major: P(r,c) = Pc[c*n + r].
// Thread = 2 state rows e; the camera columns of P stay in registers for all features.
// ------------------------------------------------------------------------------------------
constexpr int GAIN_THREADS = 128;

__global__ void __launch_bounds__(GAIN_THREADS)
k_ekf_gain(int n, int F, const double* __restrict__ P, const int32_t* __restrict__ type, const int32_t* __restrict__ pos,
           const double* __restrict__ Hcam, const double* __restrict__ Hfeat, const FrameIdx* __restrict__ idx,
           const int32_t* __restrict__ ic_list, double* __restrict__ G) {
  // blockIdx.z strides over the IC features so that a few frames still fill the machine
  const int frame = blockIdx.y;
  const FrameIdx fi = idx[frame];
  if (fi.status != 0) return;
  const double* Pc = P + (size_t)frame * n * n;
  const int e0 = blockIdx.x * (2 * GAIN_THREADS) + threadIdx.x, e1 = e0 + GAIN_THREADS;
  const bool v0 = e0 < n, v1 = e1 < n;
  double pc0[13], pc1[13];
#pragma unroll
  for (int k = 0; k < 13; ++k) {
    pc0[k] = v0 ? Pc[(size_t)k * n + e0] : 0.0;
    pc1[k] = v1 ? Pc[(size_t)k * n + e1] : 0.0;
  }
  for (int r = blockIdx.z; r < fi.num_ic; r += gridDim.z) {
    const int f = ic_list[(size_t)frame * F + r];
    const size_t ff = (size_t)frame * F + f;
    const int nf = type[ff] == 0 ? 6 : 3, p = pos[ff];
    const double* hc = Hcam + ff * 26;
    const double* hf = Hfeat + ff * 12;
    double pf0[6], pf1[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      pf0[t] = (v0 && t < nf) ? Pc[(size_t)(p + t) * n + e0] : 0.0;
      pf1[t] = (v1 && t < nf) ? Pc[(size_t)(p + t) * n + e1] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      double g0 = pc0[0] * __ldg(hc + c), g1 = pc1[0] * __ldg(hc + c);
#pragma unroll
      for (int t = 1; t < 13; ++t) {
        const double hv = __ldg(hc + 2 * t + c);
        g0 = g0 + pc0[t] * hv;
        g1 = g1 + pc1[t] * hv;
      }
#pragma unroll
      for (int t = 0; t < 6; ++t)
        if (t < nf) {
          const double hv = __ldg(hf + 2 * t + c);
          g0 = g0 + pf0[t] * hv;
          g1 = g1 + pf1[t] * hv;
        }
      double* g = G + (ff * 2 + c) * (size_t)n;
      if (v0) g[e0] = g0;
      if (v1) g[e1] = g1;
    }
  }
}

// ------------------------------------------------------------------------------------------
// seeded match selection: SPEC in oracle/pre3_oracle_ekf.c (orc_ekf_select)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

__device__ __forceinline__ void select_ranks(uint64_t seed, uint32_t frame, uint32_t hyp, int num_ic, int m, int* out) {
  for (int d = 0; d < m; ++d) {
    const uint64_t x = splitmix64(seed ^ ((uint64_t)frame * 0x9E3779B97F4A7C15ULL) ^
                                  ((uint64_t)hyp * 0xD1B54A32D192ED03ULL) ^ ((uint64_t)(d + 1) * 0x8CB92BA72F3D8DD7ULL));
    const uint32_t r = (uint32_t)(x >> 32);
    int t = (int)(((uint64_t)r * (uint64_t)(num_ic - d)) >> 32);
    // skip the ranks already taken, visiting them in ascending order (d <= 2 earlier picks)
    int s0 = 0x7fffffff, s1 = 0x7fffffff;
    if (d >= 1) s0 = out[0];
    if (d >= 2) {
      s1 = out[1];
      if (s1 < s0) {
        const int tmp = s0;
        s0 = s1;
        s1 = tmp;
      }
    }
    if (d >= 1 && t >= s0) ++t;
    if (d >= 2 && t >= s1) ++t;
    out[d] = t;
  }
}

// ------------------------------------------------------------------------------------------
// k_ekf_score
// ------------------------------------------------------------------------------------------
struct ScoreArgs {
  int n, F, H;
  const double* x;      // Fr x n
  const double* P;      // Fr x n x n (column-major)
  const int32_t* type;  // Fr x F
  const int32_t* pos;
  const double* z;      // Fr x F x 2
  const double* h;
  const double* Hcam;   // Fr x F x 26
  const double* Hfeat;  // Fr x F x 12
  const double* R;      // Fr x F x 4
  const int32_t* sel;   // Fr x H x 3 or nullptr
  const FrameIdx* idx;
  const int32_t* ic_list;
  const int32_t* meas;
  const double* G;      // Fr x F x 2 x n
  pre3_cam cam;
  double thr;
  uint64_t seed;
  uint32_t frame_id0;
  int hbeg, hend;
  const int32_t* stop;  // Fr or nullptr
  const int32_t* best;  // FINAL: winner per frame
  int32_t* supports;    // Fr x H
  uint8_t* li;          // FINAL: Fr x F
};

// What the scoring threads need from one hypothesis: written by k_ekf_hyp, read by k_ekf_score.
struct HypRec {
  double Sinv[36];  // inv(Hi*P*Hi' + R), row-major d x d
  double innov[6];  // zi - hi
  double cam[7];    // updated camera position + quaternion
  double Rw[9];     // q2r of the updated quaternion
  double w[6];      // inv(S)*(zi - hi): the FAST form x + G*w of the state update (proposes, never decides)
  int32_t sel[3];   // the selected features
  int32_t d;        // 2 m
};

struct SetupShared {  // per warp
  double H[3][2][19];
  double W[6][32];
  double aug[6][12];
  int sel[3], nf[3], pos[3];
};

// xi[e] = x[e] + sum_j K[e][j]*innov[j],  K[e][j] = sum_c G[e][c]*Sinv[c][j]  (ransac_hypotheses.m:62-63)
__device__ __forceinline__ double updated_state(const double* __restrict__ x, const double* __restrict__ G,
                                                const size_t* goff, const double* Sinv, const double* innov, int d,
                                                int e) {
  double g[6], dx = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c) g[c] = c < d ? __ldg(G + goff[c] + e) : 0.0;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    if (j < d) {
      double k = g[0] * Sinv[j];
#pragma unroll
      for (int c = 1; c < 6; ++c)
        if (c < d) k = k + g[c] * Sinv[c * d + j];
      const double term = k * innov[j];
      dx = j == 0 ? term : dx + term;
    }
  }
  return x[e] + dx;
}

// xi[e] ~ x[e] + sum_c G[e][c]*w[c] with w = inv(S)*innov: the same update re-associated (6 multiply-adds per state instead
// of 42).  Its residual only PROPOSES: k_ekf_score re-evaluates in the literal K-form above every residual that could
// decide the minimum or sit at a threshold.
__device__ __forceinline__ double updated_state_fast(const double* __restrict__ x, const double* __restrict__ G,
                                                     const size_t* goff, const double* w, int d, int e) {
  double dx = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c)
    if (c < d) dx = fma(__ldg(G + goff[c] + e), w[c], dx);
  return x[e] + dx;
}

constexpr int HYP_WARPS = 4;

// One WARP per (frame, hypothesis): the m selected matches, S = Hi*P*Hi' + R from the P sub-blocks,
// inv(S) by a cooperative Gauss-Jordan, the innovation and the updated camera states.
// grid.x covers the hypotheses [hbeg, hend) of every frame (or, with best != nullptr, the winner of
// every frame); rec is indexed [frame * rec_stride + (hyp - hbeg)].
__global__ void __launch_bounds__(HYP_WARPS * 32)
k_ekf_hyp(const ScoreArgs A, int Fr, HypRec* __restrict__ rec, int rec_stride, int use_best) {
  __shared__ SetupShared shs[HYP_WARPS];
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int per_frame = use_best ? 1 : (A.hend - A.hbeg);
  const long long w = (long long)blockIdx.x * HYP_WARPS + wl;
  if (w >= (long long)Fr * per_frame) return;
  const int frame = (int)(w / per_frame), slot = (int)(w % per_frame);
  const FrameIdx fi = A.idx[frame];
  if (fi.status != 0) return;
  int hyp;
  if (use_best) {
    hyp = A.best[frame];
    if (hyp < 0) return;
  } else {
    hyp = A.hbeg + slot;
    if (A.stop && A.stop[frame] >= 0) return;
  }
  SetupShared& sh = shs[wl];
  const int n = A.n, F = A.F, m = fi.m, d = 2 * m;
  const size_t fF = (size_t)frame * F;
  const double* Pc = A.P + (size_t)frame * n * n;
  const double* x = A.x + (size_t)frame * n;
  HypRec* out = rec + (size_t)frame * rec_stride + slot;

  // ---- the m selected matches (select_random_match.m:47-58) ----------------------------------
  {
    int rk[3] = {0, 0, 0};
    if (!A.sel) select_ranks(A.seed, A.frame_id0 + (uint32_t)frame, (uint32_t)hyp, fi.num_ic, m, rk);
    if (lane < m) {
      int f;
      if (A.sel)
        f = min(max(A.sel[((size_t)frame * A.H + hyp) * 3 + lane], 0), F - 1);
      else
        f = A.ic_list[fF + rk[lane]];
      sh.sel[lane] = f;
      sh.nf[lane] = A.type[fF + f] == 0 ? 6 : 3;
      sh.pos[lane] = A.pos[fF + f];
    }
  }
  __syncwarp();
  // ---- H rows of the selected features: 13 camera + nf feature columns --------------------------
  for (int i = lane; i < m * 2 * 19; i += 32) {
    const int a = i / 38, c = (i / 19) & 1, t = i % 19;
    const size_t ff = fF + sh.sel[a];
    double v = 0.0;
    if (t < 13)
      v = A.Hcam[ff * 26 + 2 * t + c];
    else if (t - 13 < sh.nf[a])
      v = A.Hfeat[ff * 12 + 2 * (t - 13) + c];
    sh.H[a][c][t] = v;
  }
  __syncwarp();
  // ---- W = Hi*P at the columns any selected H touches (lane = column) ----------------------------
  {
    int k = -1;
    if (lane < 13) {
      k = lane;
    } else {
      int off = lane - 13;
      for (int b = 0; b < m; ++b) {
        if (off < sh.nf[b]) {
          k = sh.pos[b] + off;
          break;
        }
        off -= sh.nf[b];
      }
    }
    if (k >= 0) {
      for (int a = 0; a < m; ++a) {
        const int nta = 13 + sh.nf[a], pa = sh.pos[a];
        double w0 = 0.0, w1 = 0.0;
        for (int ta = 0; ta < nta; ++ta) {
          const int kk = ta < 13 ? ta : pa + (ta - 13);
          const double pv = Pc[(size_t)k * n + kk];
          const double t0 = sh.H[a][0][ta] * pv, t1 = sh.H[a][1][ta] * pv;
          w0 = ta == 0 ? t0 : w0 + t0;
          w1 = ta == 0 ? t1 : w1 + t1;
        }
        sh.W[2 * a][lane] = w0;
        sh.W[2 * a + 1][lane] = w1;
      }
    }
  }
  __syncwarp();
  // ---- S = W*Hi' + R, augmented with the identity ----------------------------------------------
  for (int i = lane; i < d * d; i += 32) {
    const int r = i / d, sc = i % d, a = r >> 1, ca = r & 1, b = sc >> 1, cb = sc & 1;
    int offb = 0;
    for (int bb = 0; bb < b; ++bb) offb += sh.nf[bb];
    const int ntb = 13 + sh.nf[b];
    double acc = 0.0;
    for (int tb = 0; tb < ntb; ++tb) {
      const int col = tb < 13 ? tb : 13 + offb + (tb - 13);
      const double term = sh.W[r][col] * sh.H[b][cb][tb];
      acc = tb == 0 ? term : acc + term;
    }
    const double rblk = (a == b) ? A.R[(fF + sh.sel[a]) * 4 + 2 * cb + ca] : 0.0;
    sh.aug[r][sc] = acc + rblk;
    sh.aug[r][d + sc] = (r == sc) ? 1.0 : 0.0;
  }
  __syncwarp();
  // ---- inv(S): Gauss-Jordan with partial pivoting, lane = column of [S | I] ------------------------
  for (int col = 0; col < d; ++col) {
    int pr = col;
    double best = fabs(sh.aug[col][col]);
    for (int r = col + 1; r < d; ++r) {
      const double v = fabs(sh.aug[r][col]);
      if (v > best) {
        best = v;
        pr = r;
      }
    }
    __syncwarp();
    if (pr != col && lane < 2 * d) {
      const double t = sh.aug[col][lane];
      sh.aug[col][lane] = sh.aug[pr][lane];
      sh.aug[pr][lane] = t;
    }
    __syncwarp();
    const double piv = sh.aug[col][col];
    double fr[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) fr[r] = r < d ? sh.aug[r][col] : 0.0;
    __syncwarp();
    if (lane < 2 * d) {
      const double v = sh.aug[col][lane] / piv;
      sh.aug[col][lane] = v;
#pragma unroll
      for (int r = 0; r < 6; ++r)
        if (r < d && r != col) sh.aug[r][lane] = sh.aug[r][lane] - fr[r] * v;
    }
    __syncwarp();
  }
  // Sinv / innovation live in the first rows of W from here on (W is no longer needed)
  double* Sinv = &sh.W[0][0];
  double* innov = &sh.W[2][0];
  double* cam = &sh.W[3][0];
  for (int i = lane; i < d * d; i += 32) Sinv[i] = sh.aug[i / d][d + (i % d)];
  if (lane < d) {
    const size_t ff = fF + sh.sel[lane >> 1];
    innov[lane] = A.z[ff * 2 + (lane & 1)] - A.h[ff * 2 + (lane & 1)];
  }
  __syncwarp();
  size_t goff[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) goff[c] = c < d ? ((fF + sh.sel[c >> 1]) * 2 + (c & 1)) * (size_t)n : 0;
  if (lane < 7) cam[lane] = updated_state(x, A.G, goff, Sinv, innov, d, lane);
  __syncwarp();
  for (int i = lane; i < 36; i += 32) out->Sinv[i] = i < d * d ? Sinv[i] : 0.0;
  if (lane < 6) out->innov[lane] = lane < d ? innov[lane] : 0.0;
  if (lane < 6) {
    double wv = 0.0;
    for (int j = 0; j < d; ++j) wv = fma(lane < d ? Sinv[lane * d + j] : 0.0, innov[j], wv);
    out->w[lane] = lane < d ? wv : 0.0;
  }
  if (lane < 7) out->cam[lane] = cam[lane];
  if (lane < 3) out->sel[lane] = lane < m ? sh.sel[lane] : 0;
  if (lane == 0) {
    double Rw[9];
    q2r(&cam[3], Rw);
#pragma unroll
    for (int i = 0; i < 9; ++i) out->Rw[i] = Rw[i];
    out->d = d;
  }
}

struct ScoreShared {
  HypRec rec;
  double red[32];
  size_t goff[6];
  int ired[32];
};

// One block per (frame, hypothesis), one thread per measured feature.
// FAST: the re-associated proposal + literal recheck described below; !FAST: every residual in the literal K-form.
template <bool FINAL, bool FAST>
__global__ void __launch_bounds__(256) k_ekf_score(const ScoreArgs A, const HypRec* __restrict__ rec, int rec_stride) {
  extern __shared__ double s_res[];  // one residual per measured feature
  __shared__ ScoreShared sh;
  const int frame = blockIdx.y;
  const int tid = threadIdx.x;
  const FrameIdx fi = A.idx[frame];
  if (fi.status != 0) return;
  int hyp, slot;
  if (FINAL) {
    hyp = A.best[frame];
    slot = 0;
    if (hyp < 0) return;
  } else {
    slot = blockIdx.x;
    hyp = A.hbeg + slot;
    if (hyp >= A.hend) return;
    if (A.stop && A.stop[frame] >= 0) return;
  }
  const int n = A.n, F = A.F;
  const size_t fF = (size_t)frame * F;
  const double* x = A.x + (size_t)frame * n;
  {
    const double* src = reinterpret_cast<const double*>(rec + (size_t)frame * rec_stride + slot);
    double* dst = reinterpret_cast<double*>(&sh.rec);
    for (int i = tid; i < (int)(sizeof(HypRec) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int d = sh.rec.d;
  if (tid < 6) sh.goff[tid] = tid < d ? ((fF + sh.rec.sel[tid >> 1]) * 2 + (tid & 1)) * (size_t)n : 0;
  __syncthreads();

  // ---- one thread per measured feature ---------------------------------------------------------
  // The literal update xi = x + (P*Hi'*inv(S))*(zi - hi) (ransac_hypotheses.m:61-63) rebuilds a row of K for every
  // state of every feature of every hypothesis: 468 of the ~600 flops of an evaluation.  Here the re-associated form
  // x + G*w (36 multiply-adds per feature) gives a residual r~ that differs from the literal one by rounding only;
  // it PROPOSES, and the literal form is evaluated for every residual that could DECIDE something:
  //   - candidates for the minimum of the inverse-depth residuals (r~ within 2 delta of the smallest r~),
  //   - residuals within 2 delta of the inlier limit (min + threshold, or threshold for cartesian features),
  //   - residuals that are not finite.
  // delta = 1e-6 (1 + |r~|) pixels is ~9 orders of magnitude above the re-association error of well-conditioned
  // features; near-singular projections give huge or non-finite residuals, which take the literal path.  Supports,
  // masks and the selected hypothesis therefore equal those of the literal evaluation (the CPU checker's), at about a
  // quarter of its arithmetic: typically ONE literal evaluation per hypothesis (the minimum itself).
  auto literal = [&](int f, int ty, int p) {
    double s[6];
#pragma unroll
    for (int t = 0; t < 6; ++t)
      s[t] = (t < 3 || ty == 0) ? updated_state(x, A.G, sh.goff, sh.rec.Sinv, sh.rec.innov, d, p + t) : 0.0;
    return feature_residual(ty, s, sh.rec.cam, sh.rec.Rw, A.cam, A.z[(fF + f) * 2], A.z[(fF + f) * 2 + 1]);
  };
  auto band = [](double r) { return 2.0e-6 * (1.0 + fabs(r)); };
  const double QNAN = __longlong_as_double(0x7ff8000000000000LL);
  double lmin = QNAN;
  for (int j = tid; j < fi.n_meas; j += blockDim.x) {
    const int f = A.meas[fF + j];
    const int ty = A.type[fF + f], p = A.pos[fF + f];
    double r;
    if (FAST) {
      double s[6];
#pragma unroll
      for (int t = 0; t < 6; ++t) s[t] = (t < 3 || ty == 0) ? updated_state_fast(x, A.G, sh.goff, sh.rec.w, d, p + t) : 0.0;
      r = feature_residual(ty, s, sh.rec.cam, sh.rec.Rw, A.cam, A.z[(fF + f) * 2], A.z[(fF + f) * 2 + 1]);
      if (!(fabs(r) < 1.0e300)) r = QNAN;  // not finite: decided by the literal form below
    } else {
      r = literal(f, ty, p);
    }
    s_res[j] = r;
    if (ty == 0) lmin = nanmin2(lmin, r);
  }
  const double mn_fast = block_nanmin(lmin, sh.red);
  if (!FAST) {  // every residual is already the literal one
    const double lim0 = mn_fast + A.thr;
    int cnt0 = 0;
    for (int j = tid; j < fi.n_meas; j += blockDim.x) {
      const int f = A.meas[fF + j];
      const bool in = A.type[fF + f] == 0 ? (s_res[j] < lim0) : (s_res[j] < A.thr);
      cnt0 += in ? 1 : 0;
      if (FINAL) A.li[fF + f] = in ? 1 : 0;
    }
    if (!FINAL) {
      const int total0 = block_sum_int(cnt0, sh.ired);
      if (tid == 0) A.supports[(size_t)frame * A.H + hyp] = total0;
    }
    return;
  }
  // exact minimum: literal residuals of the candidates (every other inverse-depth residual is provably larger)
  double emin = QNAN;
  unsigned exact_mask = 0u;  // bit i: the i-th feature of this thread holds a literal residual in s_res
  {
    int it = 0;
    for (int j = tid; j < fi.n_meas; j += blockDim.x, ++it) {
      const int f = A.meas[fF + j];
      const int ty = A.type[fF + f];
      const double r = s_res[j];
      const bool cand = ty == 0 && (r != r || mn_fast != mn_fast || r <= mn_fast + band(mn_fast) + band(r));
      if (cand || r != r) {
        const double re = literal(f, ty, A.pos[fF + f]);
        s_res[j] = re;
        if (it < 32) exact_mask |= 1u << it;
        if (ty == 0) emin = nanmin2(emin, re);
      }
    }
  }
  const double mn = block_nanmin(emin, sh.red);
  const double lim = mn + A.thr;
  int cnt = 0;
  {
    int it = 0;
    for (int j = tid; j < fi.n_meas; j += blockDim.x, ++it) {
      const int f = A.meas[fF + j];
      const int ty = A.type[fF + f];
      const double limit = ty == 0 ? lim : A.thr;
      double r = s_res[j];
      const bool is_exact = it < 32 && ((exact_mask >> it) & 1u);
      if (!is_exact && !(fabs(r - limit) > band(r) + band(limit))) r = literal(f, ty, A.pos[fF + f]);  // borderline (or NaN limit)
      const bool in = r < limit;
      cnt += in ? 1 : 0;
      if (FINAL) A.li[fF + f] = in ? 1 : 0;
    }
  }
  if (!FINAL) {
    const int total = block_sum_int(cnt, sh.ired);
    if (tid == 0) A.supports[(size_t)frame * A.H + hyp] = total;
  }
}

// ------------------------------------------------------------------------------------------
// loop control (ransac_hypotheses.m:40-46, :74-80) replayed over supports[0..E)
// ------------------------------------------------------------------------------------------
struct Replay {
  int ended;  // the loop has ended within the evaluated prefix
  int n_eval, best, max_support;
  double n_hyp;
};

__device__ __forceinline__ Replay replay(const int32_t* sup, int E, int limit, int m, const double* tabrow, int adaptive,
                                         double n_hyp_init) {
  Replay r;
  r.ended = 0;
  r.n_eval = 0;
  r.best = -1;
  r.max_support = 0;
  r.n_hyp = n_hyp_init;
  const int top = E < limit ? E : limit;
  for (int i = 0; i < top; ++i) {
    if (adaptive && r.n_hyp == 0.0) {
      r.ended = 1;
      return r;
    }
    const int s = sup[i];
    r.n_eval = i + 1;
    if (s > r.max_support) {
      r.max_support = s;
      r.best = i;
      r.n_hyp = tabrow[s];
    }
    if (adaptive && r.n_hyp <= (double)m) {  // `i` at :80 is the inner loop's variable = m
      r.ended = 1;
      return r;
    }
  }
  if (top >= limit) r.ended = 1;
  return r;
}

__global__ void __launch_bounds__(128)
k_ekf_stop(int Fr, int F, int H, int E, int limit, int adaptive, double n_hyp_init, const FrameIdx* __restrict__ idx,
           const double* __restrict__ tab, const int32_t* __restrict__ supports, int32_t* __restrict__ stop) {
  const int frame = blockIdx.x * blockDim.x + threadIdx.x;
  if (frame >= Fr) return;
  if (stop[frame] >= 0) return;
  const FrameIdx fi = idx[frame];
  if (fi.status != 0) {
    stop[frame] = 0;
    return;
  }
  const Replay r = replay(supports + (size_t)frame * H, E, limit, fi.m, tab + (size_t)fi.num_ic * (F + 1), adaptive,
                          n_hyp_init);
  if (r.ended) stop[frame] = r.n_eval;
}

__global__ void __launch_bounds__(128)
k_ekf_pick(int Fr, int F, int H, int limit, int adaptive, double n_hyp_init, const FrameIdx* __restrict__ idx,
           const double* __restrict__ tab, int32_t* __restrict__ supports, pre3_ekf_result* __restrict__ res,
           int32_t* __restrict__ best, int32_t* __restrict__ supports_out) {
  const int frame = blockIdx.x * blockDim.x + threadIdx.x;
  if (frame >= Fr) return;
  const FrameIdx fi = idx[frame];
  pre3_ekf_result o;
  o.status = fi.status;
  o.n_evaluated = 0;
  o.best_hyp = -1;
  o.max_support = 0;
  o.num_ic = fi.num_ic;
  o.m = fi.status == 1 ? 0 : fi.m;
  o.n_hyp = n_hyp_init;
  if (fi.status == 0) {
    const Replay r = replay(supports + (size_t)frame * H, limit, limit, fi.m, tab + (size_t)fi.num_ic * (F + 1),
                            adaptive, n_hyp_init);
    o.n_evaluated = r.n_eval;
    o.best_hyp = r.best;
    o.max_support = r.max_support;
    o.n_hyp = r.n_hyp;
  }
  res[frame] = o;
  best[frame] = o.best_hyp;
  if (supports_out)
    for (int i = 0; i < H; ++i)
      supports_out[(size_t)frame * H + i] = i < o.n_evaluated ? supports[(size_t)frame * H + i] : -1;
}

// ------------------------------------------------------------------------------------------
// k_ekf_support_given: compute_hypothesis_support_fast for B states supplied by the caller
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ekf_support_given(const double* __restrict__ xi, int n, pre3_cam cam, const int32_t* __restrict__ ir,
                    const int32_t* __restrict__ ia, const int32_t* __restrict__ irho, const double* __restrict__ z_id,
                    int n_id, const int32_t* __restrict__ ixyz, const double* __restrict__ z_euc, int n_euc, double thr,
                    int32_t* __restrict__ support, uint8_t* __restrict__ li_id, uint8_t* __restrict__ li_euc) {
  extern __shared__ double s_res[];
  __shared__ double s_red[32], s_Rw[9], s_rwc[3];
  __shared__ int s_ired[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* x = xi + (size_t)b * n;
  if (tid == 0) {
    q2r(x + 3, s_Rw);
    s_rwc[0] = x[0];
    s_rwc[1] = x[1];
    s_rwc[2] = x[2];
  }
  __syncthreads();
  double lmin = __longlong_as_double(0x7ff8000000000000LL);
  for (int j = tid; j < n_id + n_euc; j += blockDim.x) {
    double s[6], r;
    if (j < n_id) {
      for (int a = 0; a < 3; ++a) s[a] = x[ir[3 * j + a]];
      s[3] = x[ia[2 * j]];
      s[4] = x[ia[2 * j + 1]];
      s[5] = x[irho[j]];
      r = feature_residual(0, s, s_rwc, s_Rw, cam, z_id[2 * j], z_id[2 * j + 1]);
      lmin = nanmin2(lmin, r);
    } else {
      const int e = j - n_id;
      for (int a = 0; a < 3; ++a) s[a] = x[ixyz[3 * e + a]];
      s[3] = s[4] = s[5] = 0.0;
      r = feature_residual(1, s, s_rwc, s_Rw, cam, z_euc[2 * e], z_euc[2 * e + 1]);
    }
    s_res[j] = r;
  }
  const double mn = block_nanmin(lmin, s_red);
  const double lim = mn + thr;
  int cnt = 0;
  for (int j = tid; j < n_id + n_euc; j += blockDim.x) {
    const bool in = j < n_id ? (s_res[j] < lim) : (s_res[j] < thr);
    cnt += in ? 1 : 0;
    if (j < n_id) {
      if (li_id) li_id[(size_t)b * n_id + j] = in ? 1 : 0;
    } else if (li_euc) {
      li_euc[(size_t)b * n_euc + (j - n_id)] = in ? 1 : 0;
    }
  }
  const int total = block_sum_int(cnt, s_ired);
  if (tid == 0 && support) support[b] = total;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Re-prediction at x_k_k (M/@ekf_filter/rescue_hi_inliers.m:32-33): predict_camera_measurements.m:27-68 followed by
// calculate_derivatives.m:27-59, one thread per feature.  Same specification as oracle/pre3_oracle_ekf.c
// (orc_ekf_predict): inv() = Gauss-Jordan with partial pivoting, sin / cos = sincos_spec, products accumulated left to
// right.  A feature that fails the field-of-view or image-bounds test keeps its previous h; H is produced for every
// feature whose h is not empty, evaluated at that h (calculate_Hi_inverse_depth_my_version.m:29).
// ---------------------------------------------------------------------------------------------
template <int D>
__device__ void inv_spec(const double* S, double* out) {  // row-major D x D
  double a[D][2 * D];
  for (int r = 0; r < D; ++r)
    for (int c = 0; c < D; ++c) {
      a[r][c] = S[r * D + c];
      a[r][D + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int col = 0; col < D; ++col) {
    int pr = col;
    double best = fabs(a[col][col]);
    for (int r = col + 1; r < D; ++r)
      if (fabs(a[r][col]) > best) best = fabs(a[r][col]), pr = r;
    if (pr != col)
      for (int c = 0; c < 2 * D; ++c) {
        const double t = a[col][c];
        a[col][c] = a[pr][c], a[pr][c] = t;
      }
    const double piv = a[col][col];
    for (int c = 0; c < 2 * D; ++c) a[col][c] = a[col][c] / piv;
    for (int r = 0; r < D; ++r) {
      if (r == col) continue;
      const double f = a[r][col];
      for (int c = 0; c < 2 * D; ++c) a[r][c] = a[r][c] - f * a[col][c];
    }
  }
  for (int r = 0; r < D; ++r)
    for (int c = 0; c < D; ++c) out[r * D + c] = a[r][D + c];
}

__device__ void mm_spec(const double* A, int ra, int ca, const double* B, int cb, double* C) {  // row-major
  for (int i = 0; i < ra; ++i)
    for (int j = 0; j < cb; ++j) {
      double s = A[i * ca] * B[j];
      for (int k = 1; k < ca; ++k) s = s + A[i * ca + k] * B[k * cb + j];
      C[i * cb + j] = s;
    }
}

__global__ void __launch_bounds__(64)
k_ekf_predict(const double* __restrict__ X, int n, pre3_cam cam, int nRows, int nCols, int F,
              const int32_t* __restrict__ type, const int32_t* __restrict__ pos, const uint8_t* __restrict__ has_h_in,
              const double* __restrict__ h_in, double* __restrict__ h_out, uint8_t* __restrict__ has_h_out,
              uint8_t* __restrict__ predicted, double* __restrict__ Hcam, double* __restrict__ Hfeat) {
  const int fr = blockIdx.y, i = blockIdx.x * 64 + threadIdx.x;
  if (i >= F) return;
  const double* x = X + (size_t)fr * n;
  const size_t fi = (size_t)fr * F + i;
  double Rw[9], Rrw[9];
  q2r(x + 3, Rw);
  inv_spec<3>(Rw, Rrw);
  const int ty = type[fi];
  const double* y = x + pos[fi];
  double a[3], st = 0, ct = 0, sp = 0, cp = 0, rho = 1.0;
  if (ty == 0) {
    sincos_spec(y[3], st, ct);
    sincos_spec(y[4], sp, cp);
    const double mi[3] = {cp * st, -sp, cp * ct};
    rho = y[5];
    for (int k = 0; k < 3; ++k) a[k] = (y[k] - x[k]) * rho + mi[k];
  } else {
    for (int k = 0; k < 3; ++k) a[k] = y[k] - x[k];
  }
  double hrl[3];
  if (ty == 0) {
    for (int k = 0; k < 3; ++k) hrl[k] = (Rw[k] * a[0] + Rw[3 + k] * a[1]) + Rw[6 + k] * a[2];  // r_wc' (hi_inverse_depth.m:33)
  } else {
    mm_spec(Rrw, 3, 3, a, 1, hrl);                                                                // inv(r_wc) (hi_cartesian.m:33)
  }
  const double PI = 3.14159265358979323846;
  bool ok = true;
  const double ax = atan2(hrl[0], hrl[2]) * 180 / PI, ay = atan2(hrl[1], hrl[2]) * 180 / PI;
  if (ax < -60 || ax > 60 || ay < -60 || ay > 60) ok = false;
  double ud = 0, vd = 0;
  if (ok) {
    const double u = cam.Cx + (hrl[0] / hrl[2]) * cam.f, v = cam.Cy + (hrl[1] / hrl[2]) * cam.f;
    const double xu = (u - cam.Cx) / cam.f, yu = (v - cam.Cy) / cam.f;
    const double ru = sqrt(xu * xu + yu * yu);
    const double ru2 = ru * ru;
    const double Dd = (1.0 + cam.k1 * ru2) + cam.k2 * (ru2 * ru2);
    ud = (xu * Dd) * cam.f + cam.Cx, vd = (yu * Dd) * cam.f + cam.Cy;
    if (!(ud > 0 && ud < nCols && vd > 0 && vd < nRows)) ok = false;
  }
  const bool has = ok || (has_h_in ? has_h_in[fi] != 0 : true);
  predicted[fi] = ok ? 1 : 0;
  has_h_out[fi] = has ? 1 : 0;
  const double uu = ok ? ud : h_in[2 * fi], vv = ok ? vd : h_in[2 * fi + 1];
  h_out[2 * fi] = uu;
  h_out[2 * fi + 1] = vv;
  double* Hc = Hcam + 26 * fi;
  double* Hf = Hfeat + 12 * fi;
  for (int k = 0; k < 26; ++k) Hc[k] = 0.0;
  for (int k = 0; k < 12; ++k) Hf[k] = 0.0;
  if (!has) return;  // calculate_derivatives.m:34
  double hc[3];
  mm_spec(Rrw, 3, 3, a, 1, hc);
  const double f = cam.f;
  const double dhu[6] = {f / hc[2], 0.0, -hc[0] * f / (hc[2] * hc[2]), 0.0, f / hc[2], -hc[1] * f / (hc[2] * hc[2])};
  const double xd = uu - cam.Cx, yd = vv - cam.Cy;
  const double r2 = (xd * xd + yd * yd) / (f * f), r4 = r2 * r2;
  const double k1 = cam.k1, k2 = cam.k2;
  double J[4], Ji[4], Jii[4];
  J[0] = (1 + k1 * r2 + k2 * r4) + (uu - cam.Cx) * (k1 + 2 * k2 * r2) * (2 * (uu - cam.Cx) / (f * f));
  J[3] = (1 + k1 * r2 + k2 * r4) + (vv - cam.Cy) * (k1 + 2 * k2 * r2) * (2 * (vv - cam.Cy) / (f * f));
  J[1] = (uu - cam.Cx) * (k1 + 2 * k2 * r2) * (2 * (vv - cam.Cy) / (f * f));
  J[2] = (vv - cam.Cy) * (k1 + 2 * k2 * r2) * (2 * (uu - cam.Cx) / (f * f));
  inv_spec<2>(J, Ji);
  inv_spec<2>(Ji, Jii);
  double dh[6];
  mm_spec(Jii, 2, 2, dhu, 3, dh);
  double drw[9], Hrw[6];
  for (int k = 0; k < 9; ++k) drw[k] = ty == 0 ? -Rrw[k] * rho : -Rrw[k];
  mm_spec(dh, 2, 3, drw, 3, Hrw);
  const double q0 = x[3], qx = -x[4], qy = -x[5], qz = -x[6];  // qconj (dRq_times_a_by_dq.m:26-103)
  const double dR[4][9] = {{2 * q0, -2 * qz, 2 * qy, 2 * qz, 2 * q0, -2 * qx, -2 * qy, 2 * qx, 2 * q0},
                           {2 * qx, 2 * qy, 2 * qz, 2 * qy, -2 * qx, -2 * q0, 2 * qz, 2 * q0, -2 * qx},
                           {-2 * qy, 2 * qx, 2 * q0, 2 * qx, 2 * qy, 2 * qz, -2 * q0, 2 * qz, -2 * qy},
                           {-2 * qz, -2 * q0, 2 * qx, 2 * q0, -2 * qz, 2 * qy, 2 * qx, 2 * qy, 2 * qz}};
  double dq[12];
  for (int c = 0; c < 4; ++c) {
    double t[3];
    mm_spec(dR[c], 3, 3, a, 1, t);
    for (int r = 0; r < 3; ++r) dq[r * 4 + c] = c == 0 ? t[r] : t[r] * -1.0;  // * dqbar_by_dq = diag([1 -1 -1 -1])
  }
  double Hq[8];
  mm_spec(dh, 2, 3, dq, 4, Hq);
  for (int r = 0; r < 2; ++r) {
    for (int c = 0; c < 3; ++c) Hc[2 * c + r] = Hrw[r * 3 + c];
    for (int c = 0; c < 4; ++c) Hc[2 * (3 + c) + r] = Hq[r * 4 + c];
  }
  if (ty == 0) {
    double dy[18];
    const double dth[3] = {cp * ct, 0.0, -cp * st}, dph[3] = {-sp * st, -cp, -sp * ct};
    const double yr[3] = {y[0] - x[0], y[1] - x[1], y[2] - x[2]};
    double c4[3], c5[3], c6[3];
    mm_spec(Rrw, 3, 3, dth, 1, c4);
    mm_spec(Rrw, 3, 3, dph, 1, c5);
    mm_spec(Rrw, 3, 3, yr, 1, c6);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) dy[r * 6 + c] = rho * Rrw[r * 3 + c];
      dy[r * 6 + 3] = c4[r], dy[r * 6 + 4] = c5[r], dy[r * 6 + 5] = c6[r];
    }
    double Hy[12];
    mm_spec(dh, 2, 3, dy, 6, Hy);
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 6; ++c) Hf[2 * c + r] = Hy[r * 6 + c];
  } else {
    double Hy[6];
    mm_spec(dh, 2, 3, Rrw, 3, Hy);
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 3; ++c) Hf[2 * c + r] = Hy[r * 3 + c];
  }
}

static int wave_ends(const pre3_ekf_opts& o, int32_t* ends, int cap) {
  const int limit = std::min(o.H, o.n_hyp_init);
  if (limit <= 0) return 0;
  int n = 0;
  if (!o.adaptive || limit <= 48) {
    if (n < cap) ends[n] = limit;
    return 1;
  }
  int beg = 0, width = 32;
  while (beg < limit) {
    int end = std::min(limit, beg + width);
    if (limit - end < 32) end = limit;
    if (n < cap) ends[n] = end;
    ++n;
    beg = end;
    width *= 2;
  }
  return n;
}

// n_hyp = ceil(log(1-p)/log(1-(1-epsilon))), epsilon = 1 - support/num_IC (ransac_hypotheses.m:29,77-78)
static double nhyp_rule(int support, int num_ic) {
  const double p = 0.99;
  const double epsilon = 1.0 - ((double)support / (double)num_ic);
  const double a = 1.0 - (1.0 - epsilon);
  const double L = std::log(1.0 - p);
  if (a > 0.0) return std::ceil(L / std::log(a));
  if (a == 0.0) return 0.0;
  // MATLAB: complex log of a negative number; relational operators compare real parts
  const double la = std::log(-a), pi = 3.14159265358979323846;
  return std::ceil((L * la) / (la * la + pi * pi));
}

static int ensure_table(pre3_ctx* ctx, int F) {
  if (ctx->d_ekf_tab && ctx->ekf_tab_F == F) return PRE3_OK;
  std::vector<double> tab((size_t)(F + 1) * (F + 1), 0.0);
  for (int nic = 1; nic <= F; ++nic)
    for (int s = 0; s <= F; ++s) tab[(size_t)nic * (F + 1) + s] = nhyp_rule(s, nic);
  if (ctx->d_ekf_tab) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_ekf_tab);
    ctx->d_ekf_tab = nullptr;
  }
  PRE3_CUDA(cudaMalloc((void**)&ctx->d_ekf_tab, tab.size() * sizeof(double)));
  PRE3_CUDA(cudaMemcpyAsync(ctx->d_ekf_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice,
                            ctx->stream));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->ekf_tab_F = F;
  return PRE3_OK;
}

static int frames_per_chunk(int Fr, int n, int F) {
  const size_t per = (size_t)F * 2 * n * sizeof(double);
  const size_t budget = (size_t)2 << 30;
  return (int)std::max<size_t>(1, std::min<size_t>((size_t)Fr, budget / std::max<size_t>(per, 1)));
}

static size_t ekf_ws_bytes(int C, int n, int F, int H) {
  size_t b = 0;
  b += align_up(sizeof(FrameIdx) * (size_t)C);
  b += 2 * align_up(4 * (size_t)C * F);
  b += align_up(8 * (size_t)C * F * 2 * n);
  b += align_up(4 * (size_t)C * std::max(H, 1));
  b += 2 * align_up(4 * (size_t)C);
  b += align_up(sizeof(HypRec) * (size_t)C * std::max(H, 1));  // hypothesis records of the widest wave (<= H)
  return b + 8192;
}

// all device pointers; the workspace must hold ekf_ws_bytes(chunk) beyond ctx->ws_off
static int ekf_impl(pre3_ctx* ctx, int Fr, int n, int F, const double* dx, const double* dP, double std_z,
                    const pre3_cam& cam, const int32_t* dtype, const int32_t* dpos, const uint8_t* dhas_z,
                    const uint8_t* dic, const double* dz, const double* dh, const double* dHcam, const double* dHfeat,
                    const double* dR, const int32_t* dsel, const pre3_ekf_opts& o, uint32_t frame_id0, uint8_t* dli,
                    pre3_ekf_result* dres, int32_t* dsupports) {
  PRE3_TRY(ensure_table(ctx, F));
  const int H = o.H, limit = std::min(o.H, o.n_hyp_init);
  const int C = frames_per_chunk(Fr, n, F);
  const size_t ws_mark = ctx->ws_off;
  const int bs = std::min(256, std::max(64, (F + 31) / 32 * 32));
  const size_t smem = sizeof(double) * (size_t)F;
  if (smem > 40000) return fail(ctx, PRE3_ERR_ARG, "too many features per frame (F <= 5000)");
  int32_t ends[40];
  const int nw = wave_ends(o, ends, 40);
  for (int c0 = 0; c0 < Fr; c0 += C) {
    const int nc = std::min(C, Fr - c0);
    ctx->ws_off = ws_mark;
    FrameIdx* idx = ws_take<FrameIdx>(ctx, nc);
    int32_t* ic_list = ws_take<int32_t>(ctx, (size_t)nc * F);
    int32_t* meas = ws_take<int32_t>(ctx, (size_t)nc * F);
    double* G = ws_take<double>(ctx, (size_t)nc * F * 2 * n);
    int32_t* sup = ws_take<int32_t>(ctx, (size_t)nc * std::max(H, 1));
    int32_t* stop = ws_take<int32_t>(ctx, nc);
    int32_t* best = ws_take<int32_t>(ctx, nc);
    int maxw = 1;
    for (int w = 0, beg = 0; w < nw; beg = ends[w], ++w) maxw = std::max(maxw, ends[w] - beg);
    HypRec* rec = ws_take<HypRec>(ctx, (size_t)nc * maxw);
    const size_t fo = (size_t)c0;
    {
      Span span__(ctx, T_EKF_SELECT);
      k_ekf_index<<<(nc + 3) / 4, 128, 0, ctx->stream>>>(nc, F, dhas_z + fo * F, dic + fo * F, idx, ic_list, meas);
      count_launch(ctx);
    }
    {
      Span span__(ctx, T_EKF_GAIN);
      const int gx = (n + 2 * GAIN_THREADS - 1) / (2 * GAIN_THREADS);
      const int want = 8 * ctx->sm_count;  // resident blocks to hide the dependent column loads
      const int gz = std::max(1, std::min(std::min(F, 64), (want + gx * nc - 1) / (gx * nc)));
      const dim3 grid(gx, nc, gz);
      k_ekf_gain<<<grid, GAIN_THREADS, 0, ctx->stream>>>(n, F, dP + fo * n * n, dtype + fo * F, dpos + fo * F,
                                                         dHcam + fo * F * 26, dHfeat + fo * F * 12, idx, ic_list, G);
      count_launch(ctx);
    }
    // PRE3_EKF_FAST=1: re-associated proposal + literal recheck in k_ekf_score (same supports / masks; measured slower at the
    // config-4 shape: the block's latency chain grows by one literal evaluation and two reductions)
    static const bool ekf_fast = getenv("PRE3_EKF_FAST") && atoi(getenv("PRE3_EKF_FAST")) != 0;
    PRE3_CUDA(cudaMemsetAsync(sup, 0, 4 * (size_t)nc * std::max(H, 1), ctx->stream));
    PRE3_CUDA(cudaMemsetAsync(stop, 0xFF, 4 * (size_t)nc, ctx->stream));
    ScoreArgs a;
    a.n = n;
    a.F = F;
    a.H = H;
    a.x = dx + fo * n;
    a.P = dP + fo * n * n;
    a.type = dtype + fo * F;
    a.pos = dpos + fo * F;
    a.z = dz + fo * F * 2;
    a.h = dh + fo * F * 2;
    a.Hcam = dHcam + fo * F * 26;
    a.Hfeat = dHfeat + fo * F * 12;
    a.R = dR + fo * F * 4;
    a.sel = dsel ? dsel + fo * H * 3 : nullptr;
    a.idx = idx;
    a.ic_list = ic_list;
    a.meas = meas;
    a.G = G;
    a.cam = cam;
    a.thr = std_z;
    a.seed = o.seed;
    a.frame_id0 = frame_id0 + (uint32_t)c0;
    a.stop = nw > 1 ? stop : nullptr;
    a.best = best;
    a.supports = sup;
    a.li = dli + fo * F;
    int beg = 0;
    for (int w = 0; w < nw; ++w) {
      const int end = ends[w];
      {
        Span span__(ctx, T_EKF_SCORE);
        a.hbeg = beg;
        a.hend = end;
        const long long warps = (long long)nc * (end - beg);
        k_ekf_hyp<<<(unsigned)((warps + HYP_WARPS - 1) / HYP_WARPS), HYP_WARPS * 32, 0, ctx->stream>>>(a, nc, rec, maxw, 0);
        if (ekf_fast) k_ekf_score<false, true><<<dim3(end - beg, nc), bs, smem, ctx->stream>>>(a, rec, maxw);
        else k_ekf_score<false, false><<<dim3(end - beg, nc), bs, smem, ctx->stream>>>(a, rec, maxw);
        count_launch(ctx, 2);
      }
      if (end < limit) {
        Span span__(ctx, T_EKF_SELECT);
        k_ekf_stop<<<(nc + 127) / 128, 128, 0, ctx->stream>>>(nc, F, H, end, limit, o.adaptive, (double)o.n_hyp_init,
                                                              idx, ctx->d_ekf_tab, sup, stop);
        count_launch(ctx);
      }
      beg = end;
    }
    {
      Span span__(ctx, T_EKF_SELECT);
      k_ekf_pick<<<(nc + 127) / 128, 128, 0, ctx->stream>>>(nc, F, H, limit, o.adaptive, (double)o.n_hyp_init, idx,
                                                            ctx->d_ekf_tab, sup, dres + fo, best,
                                                            dsupports ? dsupports + fo * H : nullptr);
      count_launch(ctx);
    }
    {
      Span span__(ctx, T_EKF_SCORE);
      a.hbeg = 0;
      a.hend = 1;
      k_ekf_hyp<<<(nc + HYP_WARPS - 1) / HYP_WARPS, HYP_WARPS * 32, 0, ctx->stream>>>(a, nc, rec, maxw, 1);
      if (ekf_fast) k_ekf_score<true, true><<<dim3(1, nc), bs, smem, ctx->stream>>>(a, rec, maxw);
      else k_ekf_score<true, false><<<dim3(1, nc), bs, smem, ctx->stream>>>(a, rec, maxw);
      count_launch(ctx, 2);
    }
    PRE3_CUDA(cudaGetLastError());
  }
  return PRE3_OK;
}

static int check_ekf_args(pre3_ctx* ctx, int Fr, int n, int F, const pre3_cam* cam, const pre3_ekf_opts* o) {
  if (!cam || !o) return fail(ctx, PRE3_ERR_ARG, "cam / options missing");
  if (Fr < 0 || n < 13 || F < 0) return fail(ctx, PRE3_ERR_ARG, "bad sizes (the state holds at least the 13 camera entries)");
  if (o->H < 0 || o->n_hyp_init < 0) return fail(ctx, PRE3_ERR_ARG, "H and n_hyp_init must be >= 0");
  if ((long long)Fr * std::max(o->H, 1) >= (1ll << 31)) return fail(ctx, PRE3_ERR_ARG, "frames x H must stay below 2^31");
  return PRE3_OK;
}

}  // namespace ekf
}  // namespace pre3

using namespace pre3;
using namespace pre3::ekf;

#define EKF_LIVE()                                                                                         \
  do {                                                                                                     \
    if (!ctx) return PRE3_ERR_ARG;                                                                         \
    if (ctx->device < 0) return fail(ctx, PRE3_ERR_CUDA, "no CUDA device (libpre3 has no CPU fallback)");  \
    PRE3_CUDA(cudaSetDevice(ctx->device));                                                                 \
  } while (0)

extern "C" {

int pre3_ekf_eval_schedule(const pre3_ekf_opts* opts, int32_t* ends, int cap) {
  if (!opts || (cap > 0 && !ends)) return PRE3_ERR_ARG;
  return wave_ends(*opts, ends, cap);
}

int pre3_ransac_hypotheses_batch_dev(pre3_ctx* ctx, int Fr, int n, int F, const double* dx, const double* dP,
                                     double std_z, const pre3_cam* cam, const int32_t* dtype, const int32_t* dpos,
                                     const uint8_t* dhas_z, const uint8_t* dic, const double* dz, const double* dh,
                                     const double* dHcam, const double* dHfeat, const double* dR, const int32_t* dsel,
                                     const pre3_ekf_opts* opts, uint32_t frame_id0, uint8_t* dli_inlier,
                                     pre3_ekf_result* dres, int32_t* dsupports) {
  EKF_LIVE();
  PRE3_TRY(check_ekf_args(ctx, Fr, n, F, cam, opts));
  if (!dres || !dli_inlier) return fail(ctx, PRE3_ERR_ARG, "result pointers missing");
  if (Fr == 0) return PRE3_OK;
  if (!dx || !dP || !dtype || !dpos || !dhas_z || !dic || !dz || !dh || !dHcam || !dHfeat || !dR)
    return fail(ctx, PRE3_ERR_ARG, "null pointer");
  PRE3_TRY(ws_reserve(ctx, ekf_ws_bytes(frames_per_chunk(Fr, n, F), n, F, opts->H)));
  return ekf_impl(ctx, Fr, n, F, dx, dP, std_z, *cam, dtype, dpos, dhas_z, dic, dz, dh, dHcam, dHfeat, dR, dsel, *opts,
                  frame_id0, dli_inlier, dres, dsupports);
}

// Host buffers: frames are staged in chunks (P alone is n*n*8 bytes per frame).
int pre3_ransac_hypotheses_batch(pre3_ctx* ctx, int Fr, int n, int F, const double* x, const double* P, double std_z,
                                 const pre3_cam* cam, const int32_t* type, const int32_t* pos, const uint8_t* has_z,
                                 const uint8_t* ic, const double* z, const double* h, const double* Hcam,
                                 const double* Hfeat, const double* R, const int32_t* sel, const pre3_ekf_opts* opts,
                                 uint32_t frame_id0, uint8_t* li_inlier, pre3_ekf_result* res, int32_t* supports) {
  EKF_LIVE();
  PRE3_TRY(check_ekf_args(ctx, Fr, n, F, cam, opts));
  if (!res || !li_inlier) return fail(ctx, PRE3_ERR_ARG, "result pointers missing");
  if (Fr == 0) return PRE3_OK;
  if (!x || !P || !type || !pos || !has_z || !ic || !z || !h || !Hcam || !Hfeat || !R)
    return fail(ctx, PRE3_ERR_ARG, "null pointer");
  const int H = opts->H;
  const size_t per_frame = 8 * ((size_t)n * n + n + (size_t)F * (2 + 2 + 26 + 12 + 4)) + (size_t)F * (4 + 4 + 1 + 1 + 1) +
                           (sel ? 12 * (size_t)H : 0) + sizeof(pre3_ekf_result) + (supports ? 4 * (size_t)H : 0);
  const int C = (int)std::max<size_t>(1, std::min<size_t>((size_t)Fr, ((size_t)512 << 20) / per_frame));
  PRE3_TRY(ws_reserve(ctx, ekf_ws_bytes(frames_per_chunk(C, n, F), n, F, H) + per_frame * C + 64 * 256 + 8192));
  for (int c0 = 0; c0 < Fr; c0 += C) {
    const int nc = std::min(C, Fr - c0);
    const size_t fo = (size_t)c0;
    ctx->ws_off = 0;
    double* dx = ws_take<double>(ctx, (size_t)nc * n);
    double* dP = ws_take<double>(ctx, (size_t)nc * n * n);
    int32_t* dtype = ws_take<int32_t>(ctx, (size_t)nc * F);
    int32_t* dpos = ws_take<int32_t>(ctx, (size_t)nc * F);
    uint8_t* dhz = ws_take<uint8_t>(ctx, (size_t)nc * F);
    uint8_t* dic = ws_take<uint8_t>(ctx, (size_t)nc * F);
    uint8_t* dli = ws_take<uint8_t>(ctx, (size_t)nc * F);
    double* dz = ws_take<double>(ctx, (size_t)nc * F * 2);
    double* dh = ws_take<double>(ctx, (size_t)nc * F * 2);
    double* dHc = ws_take<double>(ctx, (size_t)nc * F * 26);
    double* dHf = ws_take<double>(ctx, (size_t)nc * F * 12);
    double* dR = ws_take<double>(ctx, (size_t)nc * F * 4);
    int32_t* dsel = sel ? ws_take<int32_t>(ctx, (size_t)nc * H * 3) : nullptr;
    pre3_ekf_result* dres = ws_take<pre3_ekf_result>(ctx, nc);
    int32_t* dsup = supports ? ws_take<int32_t>(ctx, (size_t)nc * std::max(H, 1)) : nullptr;
    auto up = [&](void* d, const void* s, size_t bytes) {
      return bytes ? cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaSuccess;
    };
    PRE3_CUDA(up(dx, x + fo * n, 8 * (size_t)nc * n));
    PRE3_CUDA(up(dP, P + fo * n * n, 8 * (size_t)nc * n * n));
    PRE3_CUDA(up(dtype, type + fo * F, 4 * (size_t)nc * F));
    PRE3_CUDA(up(dpos, pos + fo * F, 4 * (size_t)nc * F));
    PRE3_CUDA(up(dhz, has_z + fo * F, (size_t)nc * F));
    PRE3_CUDA(up(dic, ic + fo * F, (size_t)nc * F));
    PRE3_CUDA(up(dli, li_inlier + fo * F, (size_t)nc * F));
    PRE3_CUDA(up(dz, z + fo * F * 2, 16 * (size_t)nc * F));
    PRE3_CUDA(up(dh, h + fo * F * 2, 16 * (size_t)nc * F));
    PRE3_CUDA(up(dHc, Hcam + fo * F * 26, 8 * 26 * (size_t)nc * F));
    PRE3_CUDA(up(dHf, Hfeat + fo * F * 12, 8 * 12 * (size_t)nc * F));
    PRE3_CUDA(up(dR, R + fo * F * 4, 32 * (size_t)nc * F));
    if (dsel) PRE3_CUDA(up(dsel, sel + fo * H * 3, 12 * (size_t)nc * H));
    PRE3_TRY(ekf_impl(ctx, nc, n, F, dx, dP, std_z, *cam, dtype, dpos, dhz, dic, dz, dh, dHc, dHf, dR, dsel, *opts,
                      frame_id0 + (uint32_t)c0, dli, dres, dsup));
    PRE3_CUDA(cudaMemcpyAsync(li_inlier + fo * F, dli, (size_t)nc * F, cudaMemcpyDeviceToHost, ctx->stream));
    PRE3_CUDA(cudaMemcpyAsync(res + fo, dres, sizeof(pre3_ekf_result) * (size_t)nc, cudaMemcpyDeviceToHost, ctx->stream));
    if (dsup)
      PRE3_CUDA(cudaMemcpyAsync(supports + fo * H, dsup, 4 * (size_t)nc * H, cudaMemcpyDeviceToHost, ctx->stream));
    PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return PRE3_OK;
}

int pre3_ekf_support(pre3_ctx* ctx, const double* xi, int n, int B, const pre3_cam* cam, const double* pattern,
                     const double* z_id, int n_id, const double* z_euc, int n_euc, double threshold, int32_t* support,
                     uint8_t* li_id, uint8_t* li_euc) {
  EKF_LIVE();
  if (!xi || !cam || !pattern) return fail(ctx, PRE3_ERR_ARG, "null pointer");
  if (n < 7 || B < 0 || n_id < 0 || n_euc < 0) return fail(ctx, PRE3_ERR_ARG, "bad sizes");
  if ((n_id > 0 && !z_id) || (n_euc > 0 && !z_euc)) return fail(ctx, PRE3_ERR_ARG, "measurements missing");
  if (B == 0) return PRE3_OK;
  // logical pattern columns -> index lists in state order (xi(logical(pattern(:,c))), :35-37,:83)
  std::vector<int32_t> lists[4];
  for (int c = 0; c < 4; ++c)
    for (int e = 0; e < n; ++e)
      if (pattern[(size_t)c * n + e] != 0.0) lists[c].push_back(e);
  if ((int)lists[0].size() != 3 * n_id || (int)lists[1].size() != 2 * n_id || (int)lists[2].size() != n_id ||
      (int)lists[3].size() != 3 * n_euc)
    return fail(ctx, PRE3_ERR_ARG, "state_vector_pattern does not match the number of measurements (reshape would fail)");
  const int nt = n_id + n_euc;
  PRE3_TRY(ws_reserve(ctx, align_up(8 * (size_t)n * B) + 4 * align_up(4 * (size_t)(3 * nt + 4)) +
                               2 * align_up(16 * (size_t)(nt + 1)) + align_up(4 * (size_t)B) +
                               2 * align_up((size_t)B * (nt + 1)) + 8192));
  double* dxi = ws_take<double>(ctx, (size_t)n * B);
  int32_t* dl[4];
  for (int c = 0; c < 4; ++c) dl[c] = ws_take<int32_t>(ctx, lists[c].size() + 1);
  double* dzi = ws_take<double>(ctx, 2 * (size_t)n_id + 2);
  double* dze = ws_take<double>(ctx, 2 * (size_t)n_euc + 2);
  int32_t* dsup = ws_take<int32_t>(ctx, B);
  uint8_t* dli = ws_take<uint8_t>(ctx, (size_t)B * n_id + 1);
  uint8_t* dle = ws_take<uint8_t>(ctx, (size_t)B * n_euc + 1);
  PRE3_CUDA(cudaMemcpyAsync(dxi, xi, 8 * (size_t)n * B, cudaMemcpyHostToDevice, ctx->stream));
  for (int c = 0; c < 4; ++c)
    if (!lists[c].empty())
      PRE3_CUDA(cudaMemcpyAsync(dl[c], lists[c].data(), 4 * lists[c].size(), cudaMemcpyHostToDevice, ctx->stream));
  if (n_id) PRE3_CUDA(cudaMemcpyAsync(dzi, z_id, 16 * (size_t)n_id, cudaMemcpyHostToDevice, ctx->stream));
  if (n_euc) PRE3_CUDA(cudaMemcpyAsync(dze, z_euc, 16 * (size_t)n_euc, cudaMemcpyHostToDevice, ctx->stream));
  const int bs = std::min(256, std::max(64, (nt + 31) / 32 * 32));
  if (8 * (size_t)nt > 40000) return fail(ctx, PRE3_ERR_ARG, "too many measurements (<= 5000)");
  {
    Span span__(ctx, T_EKF_SCORE);
    k_ekf_support_given<<<B, bs, 8 * (size_t)std::max(nt, 1), ctx->stream>>>(dxi, n, *cam, dl[0], dl[1], dl[2], dzi, n_id,
                                                                            dl[3], dze, n_euc, threshold, dsup, dli, dle);
    count_launch(ctx);
  }
  PRE3_CUDA(cudaGetLastError());
  if (support) PRE3_CUDA(cudaMemcpyAsync(support, dsup, 4 * (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
  if (li_id && n_id) PRE3_CUDA(cudaMemcpyAsync(li_id, dli, (size_t)B * n_id, cudaMemcpyDeviceToHost, ctx->stream));
  if (li_euc && n_euc) PRE3_CUDA(cudaMemcpyAsync(li_euc, dle, (size_t)B * n_euc, cudaMemcpyDeviceToHost, ctx->stream));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));  // the index vectors go out of scope
  return PRE3_OK;
}

int pre3_ekf_predict_measurements_batch_dev(pre3_ctx* ctx, int Fr, int n, int F, const double* dx, const pre3_cam* cam,
                                            int nRows, int nCols, const int32_t* dtype, const int32_t* dpos,
                                            const uint8_t* dhas_h, const double* dh_in, double* dh_out,
                                            uint8_t* dhas_h_out, uint8_t* dpredicted, double* dHcam, double* dHfeat) {
  EKF_LIVE();
  if (Fr < 0 || n < 13 || F < 0 || !cam) return fail(ctx, PRE3_ERR_ARG, "bad sizes");
  if (Fr == 0 || F == 0) return PRE3_OK;
  if (!dx || !dtype || !dpos || !dh_in || !dh_out || !dhas_h_out || !dpredicted || !dHcam || !dHfeat)
    return fail(ctx, PRE3_ERR_ARG, "null pointer");
  Span span__(ctx, T_EKF_UPDATE);
  k_ekf_predict<<<dim3((F + 63) / 64, Fr), 64, 0, ctx->stream>>>(dx, n, *cam, nRows, nCols, F, dtype, dpos, dhas_h, dh_in,
                                                                  dh_out, dhas_h_out, dpredicted, dHcam, dHfeat);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int pre3_ekf_predict_measurements_batch(pre3_ctx* ctx, int Fr, int n, int F, const double* x, const pre3_cam* cam,
                                        int nRows, int nCols, const int32_t* type, const int32_t* pos,
                                        const uint8_t* has_h, const double* h_in, double* h_out, uint8_t* has_h_out,
                                        uint8_t* predicted, double* Hcam, double* Hfeat) {
  EKF_LIVE();
  if (Fr < 0 || n < 13 || F < 0 || !cam) return fail(ctx, PRE3_ERR_ARG, "bad sizes");
  if (Fr == 0 || F == 0) return PRE3_OK;
  if (!x || !type || !pos || !h_in || !h_out || !has_h_out || !predicted || !Hcam || !Hfeat)
    return fail(ctx, PRE3_ERR_ARG, "null pointer");
  const size_t ff = (size_t)Fr * F;
  PRE3_TRY(ws_reserve(ctx, align_up(8 * (size_t)Fr * n) + 2 * align_up(4 * ff) + 3 * align_up(ff) + 2 * align_up(16 * ff) +
                               align_up(208 * ff) + align_up(96 * ff) + 4096));
  double* dx = ws_take<double>(ctx, (size_t)Fr * n);
  int32_t* dty = ws_take<int32_t>(ctx, ff);
  int32_t* dps = ws_take<int32_t>(ctx, ff);
  uint8_t* dhh = has_h ? ws_take<uint8_t>(ctx, ff) : nullptr;
  uint8_t* dho = ws_take<uint8_t>(ctx, ff);
  uint8_t* dpr = ws_take<uint8_t>(ctx, ff);
  double* dhi = ws_take<double>(ctx, 2 * ff);
  double* dhO = ws_take<double>(ctx, 2 * ff);
  double* dHc = ws_take<double>(ctx, 26 * ff);
  double* dHf = ws_take<double>(ctx, 12 * ff);
  PRE3_CUDA(cudaMemcpyAsync(dx, x, 8 * (size_t)Fr * n, cudaMemcpyHostToDevice, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(dty, type, 4 * ff, cudaMemcpyHostToDevice, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(dps, pos, 4 * ff, cudaMemcpyHostToDevice, ctx->stream));
  if (dhh) PRE3_CUDA(cudaMemcpyAsync(dhh, has_h, ff, cudaMemcpyHostToDevice, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(dhi, h_in, 16 * ff, cudaMemcpyHostToDevice, ctx->stream));
  PRE3_TRY(pre3_ekf_predict_measurements_batch_dev(ctx, Fr, n, F, dx, cam, nRows, nCols, dty, dps, dhh, dhi, dhO, dho, dpr,
                                                   dHc, dHf));
  PRE3_CUDA(cudaMemcpyAsync(h_out, dhO, 16 * ff, cudaMemcpyDeviceToHost, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(has_h_out, dho, ff, cudaMemcpyDeviceToHost, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(predicted, dpr, ff, cudaMemcpyDeviceToHost, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(Hcam, dHc, 208 * ff, cudaMemcpyDeviceToHost, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(Hfeat, dHf, 96 * ff, cudaMemcpyDeviceToHost, ctx->stream));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

}  // extern "C"
