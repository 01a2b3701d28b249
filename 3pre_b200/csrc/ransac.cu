// ransac.cu -- stages 2-4 of the 3PRE frame-to-frame path on sm_100a:
//   k_prep    per pair: threshold override (RANSAC_CALC_VER2.m:69-72), fp32 float4 copies of
//             the correspondences, magnitude bounds for the fp32 error band.
//   k_eval    one hypothesis per thread: sample set -> fp64 minimal fit (find_transform_matrix
//             or Horn) -> fp32 support scoring over float4 correspondences tiled through
//             shared memory -> fp64 recheck of threshold-borderline residuals, so the
//             cardinalities are those of the fp64 reference (RANSAC_CALC_VER2.m:96-125).
//   k_select  one block per pair: the reference's sequential loop control (adaptive stop :86,
//             :137-140; skip on state -1 :97-99) as a prefix scan, selection (max cardinality,
//             min ErrorSum, first index :165-175), winner mask, fp64 least-squares refit :186.
// Compile with -fmad=false (see fit.cuh); the fp32 scorer uses explicit fmaf().
#include "fit.cuh"
#include "ransac.cuh"

#include <cmath>
#include <cstdlib>
#include <new>

namespace pre3 {

constexpr int EV_THREADS = 128;
constexpr int EV_TILE = 512;
constexpr int EV_LIST = 1024;  // borderline list entries per block (hypothesis id in 8 bits: blocks of <= 256)
constexpr int SEL_THREADS = 256;
constexpr int EVP_THREADS = 64;  // k_eval_pairloop: sample sets per chunk = threads per block
constexpr int EVP_TILE = 384;    // a whole SR4000 pair (~300 matches) in one tile: staged once per block
constexpr int EVP_LIST = 256;
constexpr int MAX_K = 8;
constexpr int FIN_CHUNK = 2048;   // k_finish: correspondences per compaction round of the ordered ErrorSum

// ------------------------------------------------------------------------------------------
// seeded sample sets: SPEC in oracle/pre3_oracle.c (orc_sample_set), the stand-in for
// get_rand.m:43-48 (ascending k-subset).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

template <int KN>
__device__ __forceinline__ void sample_set(uint64_t seed, uint32_t pair, uint32_t hyp, int N, int k_rt, int* out) {
  const int k = KN > 0 ? KN : k_rt;
#pragma unroll
  for (int d = 0; d < (KN > 0 ? KN : MAX_K); ++d) {
    if (d >= k) break;
    const int j = N - k + d;
    const uint64_t x = splitmix64(seed ^ ((uint64_t)pair * 0x9E3779B97F4A7C15ULL) ^
                                  ((uint64_t)hyp * 0xD1B54A32D192ED03ULL) ^
                                  ((uint64_t)(d + 1) * 0x8CB92BA72F3D8DD7ULL));
    const uint32_t r = (uint32_t)(x >> 32);
    const int t = (int)(((uint64_t)r * (uint64_t)(j + 1)) >> 32);
    bool dup = false;
#pragma unroll
    for (int i = 0; i < (KN > 0 ? KN : MAX_K); ++i)
      if (i < d && out[i] == t) dup = true;
    int pick = dup ? j : t;
    // sorted insert with static indexing: bubble the new value down
    out[d] = pick;
#pragma unroll
    for (int i = (KN > 0 ? KN : MAX_K) - 1; i > 0; --i)
      if (i <= d && out[i - 1] > out[i]) {
        const int tmp = out[i - 1];
        out[i - 1] = out[i];
        out[i] = tmp;
      }
  }
}

// ------------------------------------------------------------------------------------------
// k_prep
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_prep(const double* __restrict__ Ya, const double* __restrict__ Yb,
                                              const int32_t* __restrict__ n_corr, int Nmax, int method,
                                              double thr_opt, const double* __restrict__ thr_override,
                                              int tab_triangular, const int32_t* __restrict__ rowoff,
                                              PairMeta* __restrict__ meta, float4* __restrict__ Ya4,
                                              float4* __restrict__ Yb4) {
  const int p = blockIdx.x;
  int N = n_corr ? n_corr[p] : Nmax;
  N = max(0, min(N, Nmax));
  const double* ya = Ya + (size_t)p * Nmax * 3;
  const double* yb = Yb + (size_t)p * Nmax * 3;
  float y1 = 0.f, xm = 0.f;
  double bz = INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double ax = ya[3 * i], ay = ya[3 * i + 1], az = ya[3 * i + 2];
    const double bx = yb[3 * i], by = yb[3 * i + 1], bzz = yb[3 * i + 2];
    Ya4[(size_t)p * Nmax + i] = make_float4(__double2float_rn(ax), __double2float_rn(ay), __double2float_rn(az), 0.f);
    Yb4[(size_t)p * Nmax + i] = make_float4(__double2float_rn(bx), __double2float_rn(by), __double2float_rn(bzz), 0.f);
    y1 = fmaxf(y1, __double2float_ru(fabs(bx) + fabs(by) + fabs(bzz)));
    xm = fmaxf(xm, __double2float_ru(fmax(fabs(ax), fmax(fabs(ay), fabs(az)))));
    if (bzz < bz || (bzz == bz && i < bi)) {
      bz = bzz;
      bi = i;
    }
  }
  __shared__ float s_y1[256], s_xm[256];
  __shared__ double s_bz[256];
  __shared__ int s_bi[256];
  s_y1[threadIdx.x] = y1;
  s_xm[threadIdx.x] = xm;
  s_bz[threadIdx.x] = bz;
  s_bi[threadIdx.x] = bi;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      const int o = threadIdx.x + off;
      s_y1[threadIdx.x] = fmaxf(s_y1[threadIdx.x], s_y1[o]);
      s_xm[threadIdx.x] = fmaxf(s_xm[threadIdx.x], s_xm[o]);
      if (s_bz[o] < s_bz[threadIdx.x] || (s_bz[o] == s_bz[threadIdx.x] && s_bi[o] < s_bi[threadIdx.x])) {
        s_bz[threadIdx.x] = s_bz[o];
        s_bi[threadIdx.x] = s_bi[o];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double thr = thr_opt;
    if (method == PRE3_METHOD_SVD) {
      if (N > 0) {
        const int j = s_bi[0] == 0x7fffffff ? 0 : s_bi[0];
        const double x = yb[3 * j], y = yb[3 * j + 1], z = yb[3 * j + 2];
        thr = 0.01 * sqrt((x * x + y * y) + z * z);
      } else {
        thr = 0.0;
      }
    }
    if (thr_override) thr = *thr_override;
    PairMeta m;
    m.thr = thr;
    m.thr2 = __double2float_rn(thr * thr);
    m.y1max = s_y1[0];
    m.xmax = s_xm[0];
    m.N = N;
    m.pad = tab_triangular ? (int32_t)(((long long)N * (N + 1)) / 2) : (rowoff ? rowoff[N] : 0);
    meta[p] = m;
  }
}

// ------------------------------------------------------------------------------------------
// k_eval
// ------------------------------------------------------------------------------------------
// MODE 0: find_transform_matrix fit, MODE 1: Horn fit, MODE 2: (R,t) given (column-major R),
// MODE 3: the code_from_dr_ye variant (ransac_dr_ye.m:59-70): find_transform_matrix fit on the sample in the
//         order given, EVERY hypothesis scored (rot = H, trans = 0 when the fit fails), and the threshold
//         meta.thr bounds the SQUARED distance (meta.thr2 = fl32(thr)).
template <int MODE>
__device__ __forceinline__ bool exact_inlier(const double* R, const double* t, const double* ya, const double* yb,
                                             double thr) {
  if (MODE == 3) return residual_sq(R, t, ya, yb) < thr;
  return residual_norm(R, t, ya, yb) < thr;
}

// Packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2: two lanes of work per issue slot).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

// thread = hypothesis: sample set -> fp64 minimal fit -> fp32 support scoring with PACKED arithmetic: every FFMA2 /
// FADD2 / FMUL2 scores TWO correspondences against the thread's hypothesis (its 12 coefficients sit duplicated in
// register pairs).  The correspondences are tiled through shared memory coordinate by coordinate (bx[], by[], bz[],
// -ax[], -ay[], -az[]), so one broadcast LDS.128 brings one coordinate of four correspondences: per 4 evals 6 LDS.128 +
// 32 packed FMA-pipe instructions + ~9 for the count (sign bit of r^2 - thr^2) and the borderline test (one
// rarely-taken branch).  Measured against the scalar loop of round 1 and three other mappings in tools/evalbench.cu
// (B200, N = 20 000): 1.76 vs 1.22 T evals/s (2 hypotheses x 1 match 1.55, 4 hypotheses x 1 match 1.68).
struct EvalHyp {     // what the scorer keeps of one hypothesis
  float2 cf[12];     // R (row-major) and t, each duplicated
  float2 nthr;       // -thr^2 (twice); +1 when the hypothesis is not scored
  float delta;       // certified fp32 error band around thr^2; -1 when not scored
  int state;
  bool scored, exact_me;
};

// sample set -> fit (registers) -> fp64 (R, t) into sRt_row[12], fp32 copies + error band into hy
template <int K, int MODE>
__device__ __forceinline__ void eval_fit(const PairMeta& m, const double* __restrict__ ya, const double* __restrict__ yb,
                                         bool valid, int p, int h, const int32_t* __restrict__ samples, uint64_t seed,
                                         uint32_t pair_id0, long long h0, int H, const double* __restrict__ Rin,
                                         const double* __restrict__ Tin, double* __restrict__ sRt_row, EvalHyp& hy) {
  const int N = m.N;
  hy.state = 0;
  hy.exact_me = false;
  Rigid fit;
#pragma unroll
  for (int i = 0; i < 9; ++i) fit.R[i] = 0.0;
  fit.t[0] = fit.t[1] = fit.t[2] = 0.0;
  if (valid) {
    if (MODE == 2) {
      const double* r = Rin + ((size_t)p * H + h) * 9;
      const double* t = Tin + ((size_t)p * H + h) * 3;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) fit.R[3 * rr + cc] = r[3 * cc + rr];
      fit.t[0] = t[0];
      fit.t[1] = t[1];
      fit.t[2] = t[2];
      hy.state = 1;
    } else {
      int idx[K > 0 ? K : 1];
      if (samples) {
        const int32_t* s = samples + ((size_t)p * H + h) * K;
#pragma unroll
        for (int i = 0; i < K; ++i) idx[i] = min(max(s[i], 0), N - 1);
      } else {
        sample_set<K>(seed, pair_id0 + (uint32_t)p, (uint32_t)(h0 + h), N, K, idx);
      }
      double pa[K > 0 ? K : 1][3], pb[K > 0 ? K : 1][3];
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          pa[i][r] = ya[3 * idx[i] + r];
          pb[i][r] = yb[3 * idx[i] + r];
        }
      auto get = [&](int i, double* a, double* b) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          a[r] = pa[i][r];
          b[r] = pb[i][r];
        }
      };
      if (MODE == 0 || MODE == 3)
        hy.state = fit_kabsch<K>(K, get, fit);
      else
        hy.state = fit_horn<K>(K, get, fit);
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) sRt_row[i] = fit.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) sRt_row[9 + i] = fit.t[i];

  hy.scored = valid && !(MODE == 0 && hy.state == -1);
  // fp32 copies and the certified error band
  float c32[12];
#pragma unroll
  for (int i = 0; i < 9; ++i) c32[i] = __double2float_rn(fit.R[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i) c32[9 + i] = __double2float_rn(fit.t[i]);
  float thr2 = m.thr2;
  {
    float rmax = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) rmax = fmaxf(rmax, __double2float_ru(fabs(fit.R[i])));
    const float tmax = fmaxf(fabsf(c32[9]), fmaxf(fabsf(c32[10]), fabsf(c32[11]))) * 1.0000002f;
    // |e32 - e| <= 12 u (Rmax*|yb|_1 + |t| + |ya|): 2u input rounding per product term, u per
    // fma/sub rounding, generous constant; u = 2^-24.
    const float bound = rmax * m.y1max + tmax + m.xmax;
    const float eps = 1.001f * 12.0f * 5.9604645e-8f * bound;
    const float thrf = MODE == 3 ? __fsqrt_ru(__double2float_ru(m.thr)) : __double2float_ru(m.thr);
    // |r2_32 - r2| <= eps (2 sqrt(3) r + 3 eps) + 4u r2 near r = thr; doubled for slack.
    hy.delta = 1.01f * (2.0f * eps * (3.4641018f * thrf + 3.0f * eps) + 8.0f * 5.9604645e-8f * thr2);
    // a garbage fit (state 0: rot = H) with non-finite or huge entries could produce inf - inf = NaN in the fp32
    // scorer, whose sign bit means nothing: such a hypothesis is counted in fp64 instead
    hy.exact_me = hy.scored && !(bound < 1.0e18f);
  }
  if (!hy.scored || hy.exact_me) {  // can neither count nor be borderline in the fp32 pass
    thr2 = -1.0f;
    hy.delta = -1.0f;
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) hy.cf[i] = make_float2(c32[i], c32[i]);
  hy.nthr = make_float2(-thr2, -thr2);
}

// one tile of correspondences -> shared memory, coordinate by coordinate; padded to `cap` with a point no hypothesis
// can reach (never an inlier, never borderline)
template <int NT, int TL>
__device__ __forceinline__ void eval_stage(float (*sM)[TL], const float4* __restrict__ A4,
                                           const float4* __restrict__ B4, int tn, int cap) {
  for (int i = threadIdx.x; i < cap; i += NT) {
    float4 a = make_float4(1.0e9f, 1.0e9f, 1.0e9f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < tn) {
      a = A4[i];
      b = B4[i];
    }
    sM[0][i] = b.x, sM[1][i] = b.y, sM[2][i] = b.z, sM[3][i] = -a.x, sM[4][i] = -a.y, sM[5][i] = -a.z;
  }
}

// the hot loop: support of one hypothesis over the staged tile (tn correspondences, global index base + i)
template <int TL, int LIST>
__device__ __forceinline__ int eval_score_tile(const float (*sM)[TL], int tn, int base, const EvalHyp& hy,
                                               int hyp_local, uint32_t* sList, int* sListN) {
  int cnt = 0;
  const int tn4 = (tn + 3) & ~3;
  // the operands of trip i + 1 are loaded before trip i is computed (the kernel runs at 16-20 warps per SM: the ~30
  // cycles of a shared-memory load are not hidden by other warps alone)
  float4 nbx = *reinterpret_cast<const float4*>(&sM[0][0]), nby = *reinterpret_cast<const float4*>(&sM[1][0]);
  float4 nbz = *reinterpret_cast<const float4*>(&sM[2][0]), nnx = *reinterpret_cast<const float4*>(&sM[3][0]);
  float4 nny = *reinterpret_cast<const float4*>(&sM[4][0]), nnz = *reinterpret_cast<const float4*>(&sM[5][0]);
  for (int i = 0; i < tn4; i += 4) {
    const float4 bx = nbx, by = nby, bz = nbz, nx = nnx, ny = nny, nz = nnz;
    if (i + 4 < tn4) {
      nbx = *reinterpret_cast<const float4*>(&sM[0][i + 4]);
      nby = *reinterpret_cast<const float4*>(&sM[1][i + 4]);
      nbz = *reinterpret_cast<const float4*>(&sM[2][i + 4]);
      nnx = *reinterpret_cast<const float4*>(&sM[3][i + 4]);
      nny = *reinterpret_cast<const float4*>(&sM[4][i + 4]);
      nnz = *reinterpret_cast<const float4*>(&sM[5][i + 4]);
    }
    float2 d[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const float2 x = g ? make_float2(bx.z, bx.w) : make_float2(bx.x, bx.y);
      const float2 y = g ? make_float2(by.z, by.w) : make_float2(by.x, by.y);
      const float2 z = g ? make_float2(bz.z, bz.w) : make_float2(bz.x, bz.y);
      const float2 px = g ? make_float2(nx.z, nx.w) : make_float2(nx.x, nx.y);
      const float2 py = g ? make_float2(ny.z, ny.w) : make_float2(ny.x, ny.y);
      const float2 pz = g ? make_float2(nz.z, nz.w) : make_float2(nz.x, nz.y);
      const float2 ex = fadd2(ffma2(hy.cf[0], x, ffma2(hy.cf[1], y, ffma2(hy.cf[2], z, hy.cf[9]))), px);
      const float2 ey = fadd2(ffma2(hy.cf[3], x, ffma2(hy.cf[4], y, ffma2(hy.cf[5], z, hy.cf[10]))), py);
      const float2 ez = fadd2(ffma2(hy.cf[6], x, ffma2(hy.cf[7], y, ffma2(hy.cf[8], z, hy.cf[11]))), pz);
      // r^2 - thr^2 with -thr^2 as the innermost addend: one packed instruction fewer per two evals than squaring first
      // and subtracting last; three roundings of at most u max(r^2, thr^2) each, inside the 4u r^2 term of hy.delta
      d[g] = ffma2(ex, ex, ffma2(ey, ey, ffma2(ez, ez, hy.nthr)));
    }
    cnt += (int)(__float_as_uint(d[0].x) >> 31) + (int)(__float_as_uint(d[0].y) >> 31) +
           (int)(__float_as_uint(d[1].x) >> 31) + (int)(__float_as_uint(d[1].y) >> 31);
    const float mn = fminf(fminf(fabsf(d[0].x), fabsf(d[0].y)), fminf(fabsf(d[1].x), fabsf(d[1].y)));
    if (mn <= hy.delta) {  // a threshold-borderline residual among the 4: fp64 recheck after the loop
      const float dd[4] = {d[0].x, d[0].y, d[1].x, d[1].y};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (fabsf(dd[e]) <= hy.delta) {
          const int slot = atomicAdd(sListN, 1);
          if (slot < LIST)
            sList[slot] = ((uint32_t)hyp_local << 24) | ((dd[e] < 0.f ? 1u : 0u) << 23) | (uint32_t)(base + i + e);
        }
    }
  }
  return cnt;
}

// fp64 recheck of the queued borderline residuals (counts become those of the fp64 reference); hypotheses flagged
// exact_me (or every scored one when the list overflowed) are recounted in fp64.  Call with all NT threads.
template <int MODE, int NT, int LIST>
__device__ __forceinline__ void eval_recheck(const PairMeta& m, const double* __restrict__ ya,
                                             const double* __restrict__ yb, const double* sRt, int* sCnt,
                                             const uint32_t* sList, const int* sListN, bool scored, bool exact_me) {
  const int tid = threadIdx.x;
  __syncthreads();
  const int nl = *sListN;
  if (nl > LIST) exact_me = scored;  // too many borderline residuals (degenerate scale)
  if (nl <= LIST) {
    for (int it = tid; it < nl; it += NT) {
      const uint32_t item = sList[it];
      const int hl = item >> 24;
      const bool in32 = (item >> 23) & 1u;
      const int mi = item & 0x7FFFFFu;
      const bool in64 = exact_inlier<MODE>(&sRt[hl * 12], &sRt[hl * 12 + 9], ya + 3 * mi, yb + 3 * mi, m.thr);
      if (in64 != in32) atomicAdd(&sCnt[hl], in64 ? 1 : -1);
    }
  }
  __syncthreads();
  if (exact_me) {
    int c = 0;
    for (int i = 0; i < m.N; ++i)
      c += exact_inlier<MODE>(&sRt[tid * 12], &sRt[tid * 12 + 9], ya + 3 * i, yb + 3 * i, m.thr) ? 1 : 0;
    sCnt[tid] = c;
  }
}

// grid (sample-set blocks, pairs): the fixed-H paths, single large pairs, small adaptive batches (waves)
template <int K, int MODE>
__global__ void __launch_bounds__(EV_THREADS, 6)
k_eval(const PairMeta* __restrict__ meta, const double* __restrict__ Ya, const double* __restrict__ Yb,
       const float4* __restrict__ Ya4, const float4* __restrict__ Yb4, int Nmax,
       const int32_t* __restrict__ samples, uint64_t seed, uint32_t pair_id0, long long h0, int H,
       int hbeg, int hend, const int32_t* __restrict__ stop,
       const double* __restrict__ Rin, const double* __restrict__ Tin, int32_t* __restrict__ counts,
       int8_t* __restrict__ states) {
  __shared__ __align__(16) float sM[6][EV_TILE];
  __shared__ double sRt[EV_THREADS * 12];
  __shared__ int sCnt[EV_THREADS];
  __shared__ uint32_t sList[EV_LIST];
  __shared__ int sListN;

  const int p = blockIdx.y;
  const int tid = threadIdx.x;
  // the reference's loop already ended inside the sample sets evaluated by earlier waves
  if (stop && stop[p] >= 0) return;
  const int h = hbeg + blockIdx.x * EV_THREADS + tid;
  const PairMeta m = meta[p];
  const int N = m.N;
  const double* ya = Ya + (size_t)p * Nmax * 3;
  const double* yb = Yb + (size_t)p * Nmax * 3;
  if (tid == 0) sListN = 0;

  const bool valid = (h < hend) && (MODE == 2 || N >= K) && N > 0;
  EvalHyp hy;
  eval_fit<K, MODE>(m, ya, yb, valid, p, h, samples, seed, pair_id0, h0, H, Rin, Tin, &sRt[tid * 12], hy);
  int cnt = 0;
  for (int base = 0; base < N; base += EV_TILE) {
    const int tn = min(EV_TILE, N - base);
    __syncthreads();
    eval_stage<EV_THREADS, EV_TILE>(sM, Ya4 + (size_t)p * Nmax + base, Yb4 + (size_t)p * Nmax + base, tn, (tn + 3) & ~3);
    __syncthreads();
    cnt += eval_score_tile<EV_TILE, EV_LIST>(sM, tn, base, hy, tid, sList, &sListN);
  }
  sCnt[tid] = cnt;
  eval_recheck<MODE, EV_THREADS, EV_LIST>(m, ya, yb, sRt, sCnt, sList, &sListN, hy.scored, hy.exact_me);
  if (h < hend) {
    counts[(size_t)p * H + h] = hy.scored ? sCnt[tid] : -1;
    if (states) states[(size_t)p * H + h] = (int8_t)hy.state;
  }
}

// ------------------------------------------------------------------------------------------
// k_stop: one warp per pair.  Replays the reference's loop control (RANSAC_CALC_VER2.m:86,
// :97-99, :137-140) over the first E evaluated sample sets; if the loop condition fails at some
// s < E the pair needs no further hypotheses: stop[p] = s (k_select recomputes the same value).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_stop(const PairMeta* __restrict__ meta, const int32_t* __restrict__ counts, const int8_t* __restrict__ states,
       int P, int H, int E, int method, int max_iteration, const int32_t* __restrict__ tab,
       int32_t* __restrict__ stop) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= P) return;
  if (stop[p] >= 0) return;
  const PairMeta m = meta[p];
  const int N = m.N;
  const int32_t* cnt = counts + (size_t)p * H;
  const int8_t* sts = states + (size_t)p * H;
  const int32_t* trow = tab + m.pad;
  const int ipt = (E + 31) / 32;
  const int lo = min(E, lane * ipt), hi = min(E, lo + ipt);
  int lc = 0, lm = 0;
  for (int s = lo; s < hi; ++s) {
    const bool rec = !(method == PRE3_METHOD_SVD && sts[s] == -1);
    if (rec) {
      ++lc;
      lm = max(lm, cnt[s]);
    }
  }
  // exclusive prefix (sum of recorded, max of cardinalities) over the lanes
  int pc = lc, pm = lm;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int c = __shfl_up_sync(0xffffffffu, pc, off);
    const int x = __shfl_up_sync(0xffffffffu, pm, off);
    if (lane >= off) {
      pc += c;
      pm = max(pm, x);
    }
  }
  pc = __shfl_up_sync(0xffffffffu, pc, 1);
  pm = __shfl_up_sync(0xffffffffu, pm, 1);
  if (lane == 0) {
    pc = 0;
    pm = 0;
  }
  int my_stop = 0x7fffffff;
  for (int s = lo; s < hi; ++s) {
    int nit = max_iteration;
    if (pm >= 5) nit = min(trow[min(pm, N)], max_iteration);
    if (!(1 + pc < nit)) {
      my_stop = s;
      break;
    }
    const bool rec = !(method == PRE3_METHOD_SVD && sts[s] == -1);
    if (rec) {
      ++pc;
      pm = max(pm, cnt[s]);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) my_stop = min(my_stop, __shfl_xor_sync(0xffffffffu, my_stop, off));
  if (lane == 0 && my_stop != 0x7fffffff) stop[p] = my_stop;
}

// ------------------------------------------------------------------------------------------
// block helpers for k_select
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int tid = threadIdx.x;
  __syncthreads();
  scratch[tid] = v;
  __syncthreads();
  for (int off = SEL_THREADS / 2; off > 0; off >>= 1) {
    if (tid < off) scratch[tid] = scratch[tid] + scratch[tid + off];
    __syncthreads();
  }
  const double r = scratch[0];
  __syncthreads();
  return r;
}

struct SelShared {
  double scratch[SEL_THREADS];
  int pre_cnt[SEL_THREADS];
  int pre_max[SEL_THREADS];
  int i0[SEL_THREADS];
  double tie_es[SEL_THREADS / 32];
  int tie_s[SEL_THREADS / 32];
  double Rt[12];
  int stop, maxc, winner, first_rec, first_nonmax, n_iter;
};

__device__ __forceinline__ void load_sample(const int32_t* samples, uint64_t seed, uint32_t pair, long long h0,
                                            int s, int H, int p, int N, int k, int* idx) {
  if (samples) {
    const int32_t* sp = samples + ((size_t)p * H + s) * k;
    for (int i = 0; i < k; ++i) idx[i] = min(max(sp[i], 0), N - 1);
  } else {
    sample_set<0>(seed, pair, (uint32_t)(h0 + s), N, k, idx);
  }
}

__device__ __forceinline__ int fit_sample(int method, const double* ya, const double* yb, const int* idx, int k,
                                          Rigid& out) {
  auto get = [&](int i, double* a, double* b) {
    const int j = idx[i];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      a[r] = ya[3 * j + r];
      b[r] = yb[3 * j + r];
    }
  };
  if (method == PRE3_METHOD_SVD) return fit_kabsch<0>(k, get, out);
  return fit_horn<0>(k, get, out);
}

// Least-squares refit over the masked correspondences by the whole block; thread 0 gets the
// result.  Sums are formed per thread over a strided subset and then tree-reduced in a fixed
// order (deterministic; differs from the reference's sequential sum only in rounding).
__device__ __forceinline__ int block_refit(int method, const double* ya, const double* yb,
                                           const uint8_t* mask, int N, double* scratch, Rigid& out,
                                           double threshold = 0.00000000001) {
  const int tid = threadIdx.x;
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  double ns = 0.0;
  for (int i = tid; i < N; i += SEL_THREADS)
    if (mask[i]) {
      ns += 1.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        c1[r] += ya[3 * i + r];
        c2[r] += yb[3 * i + r];
      }
    }
  ns = block_sum(ns, scratch);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c1[r] = block_sum(c1[r], scratch) / ns;
    c2[r] = block_sum(c2[r], scratch) / ns;
  }
  if (method == PRE3_METHOD_SVD) {
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < N; i += SEL_THREADS)
      if (mask[i]) {
        double q1[3], q2[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          q1[r] = ya[3 * i + r] - c1[r];
          q2[r] = yb[3 * i + r] - c2[r];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) H[3 * r + c] = H[3 * r + c] + q2[r] * q1[c];
      }
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = block_sum(H[i], scratch);
    int st = 0;
    if (tid == 0) st = kabsch_from_H(H, c1, c2, out, threshold);
    return st;
  } else {
    double M[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = 0.0;
    for (int i = tid; i < N; i += SEL_THREADS)
      if (mask[i]) {
        double an[3], bn[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          an[r] = yb[3 * i + r] - c2[r];  // A = Yb (current), centroid Ca = c2
          bn[r] = ya[3 * i + r] - c1[r];  // B = Ya (previous), centroid Cb = c1
        }
        horn_accumulate(M, an, bn);
      }
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = block_sum(M[i], scratch);
    if (tid == 0) horn_from_M(M, c2, c1, out);
    return 1;
  }
}

__device__ __forceinline__ void store_colmajor(double* dst, const double* Rrow) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[3 * c + r] = Rrow[3 * r + c];
}

// warp helpers for the selection kernels
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// sequential (index-order) sum of the residual norms of the inliers of (R,t): ErrorSum of
// RANSAC_CALC_VER2.m:135; also writes the mask when asked.  Every lane returns the sum.
__device__ __forceinline__ double warp_errsum(const double* R, const double* t, const double* ya, const double* yb,
                                              int N, double thr, uint8_t* mask, int* count_out) {
  const int lane = threadIdx.x & 31;
  double es = 0.0;
  int c = 0;
  for (int ib = 0; ib < N; ib += 32) {
    const int i = ib + lane;
    double nr = 0.0;
    bool in = false;
    if (i < N) {
      nr = residual_norm(R, t, ya + 3 * i, yb + 3 * i);
      in = nr < thr;
      if (mask) mask[i] = in ? 1 : 0;
    }
    unsigned inb = __ballot_sync(0xffffffffu, in);
    c += __popc(inb);
    while (inb) {
      const int li = __ffs(inb) - 1;
      inb &= inb - 1;
      es = es + __shfl_sync(0xffffffffu, nr, li);
    }
  }
  if (count_out) *count_out = c;
  return es;
}

__device__ __forceinline__ int warp_refit(int method, const double* ya, const double* yb, const uint8_t* mask, int N,
                                          Rigid& out) {
  const int lane = threadIdx.x & 31;
  double c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
  double ns = 0.0;
  for (int i = lane; i < N; i += 32)
    if (mask[i]) {
      ns += 1.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        c1[r] += ya[3 * i + r];
        c2[r] += yb[3 * i + r];
      }
    }
  ns = warp_sum(ns);
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c1[r] = warp_sum(c1[r]) / ns;
    c2[r] = warp_sum(c2[r]) / ns;
  }
  if (method == PRE3_METHOD_SVD) {
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = lane; i < N; i += 32)
      if (mask[i]) {
        double q1[3], q2[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          q1[r] = ya[3 * i + r] - c1[r];
          q2[r] = yb[3 * i + r] - c2[r];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) H[3 * r + c] = H[3 * r + c] + q2[r] * q1[c];
      }
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = warp_sum(H[i]);
    return kabsch_from_H(H, c1, c2, out);  // every lane computes the same result
  } else {
    double M[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = 0.0;
    for (int i = lane; i < N; i += 32)
      if (mask[i]) {
        double an[3], bn[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          an[r] = yb[3 * i + r] - c2[r];
          bn[r] = ya[3 * i + r] - c1[r];
        }
        horn_accumulate(M, an, bn);
      }
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = warp_sum(M[i]);
    horn_from_M(M, c2, c1, out);
    return 1;
  }
}

// ------------------------------------------------------------------------------------------
// Selection (RANSAC_CALC_VER2.m:86-175 and the refit :186) in three kernels:
//   k_sel_scan   one block per pair (one warp for batches of pairs, 1024 threads for a single
//                large pair): the reference's sequential loop control (:86, :97-99, :137-140) as
//                a tiled prefix scan that stops at the first tile in which the loop ends; max
//                cardinality; the list of hypotheses tied at that cardinality.
//   k_sel_tie    one THREAD per tied hypothesis: minimal fit again + ErrorSum (:135), summed
//                strictly in index order (non-inliers add an exact +0.0).
//   k_sel_final  one warp per pair: [C,I] = min(eee1) (:165-175: min ErrorSum among the ties,
//                10000 for every other recorded hypothesis, first index wins), winner mask,
//                least-squares refit on the support set.
// ------------------------------------------------------------------------------------------
struct SelInfo {
  int32_t status;        // 0 ok; 1 fewer than k correspondences; 2 nothing recorded
  int32_t S_end;         // sample sets consumed
  int32_t maxc;          // max cardinality over the recorded hypotheses
  int32_t n_iter;        // recorded hypotheses (length(M))
  int32_t first_nonmax;  // first recorded hypothesis that is not a tie (0x7fffffff: none)
  int32_t n_ties;
};

__device__ __forceinline__ void result_init(pre3_pair_result* out, int status, int n_consumed, int N, double thr) {
  out->status = status;
  out->state = 0;
  out->best_fit = 0;
  out->best_sample = -1;
  out->best_iter = 0;
  out->n_iter = 0;
  out->n_consumed = n_consumed;
  out->n_matches = N;
  out->thr = thr;
  out->error_sum = 0.0;
  for (int i = 0; i < 9; ++i) out->R[i] = out->R_hyp[i] = 0.0;
  for (int i = 0; i < 3; ++i) out->T[i] = out->T_hyp[i] = 0.0;
}

// Block-wide EXCLUSIVE scan of (sum, max) over one value per thread (identity 0, 0); also
// returns the block totals.  NT threads, NT a multiple of 32 (NT == 32: pure warp shuffles).
template <int NT>
__device__ __forceinline__ void block_exscan_cm(int c, int m, int& ex_c, int& ex_m, int& tot_c, int& tot_m,
                                                int* s_c, int* s_m) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int cc = __shfl_up_sync(0xffffffffu, c, off);
    const int mm = __shfl_up_sync(0xffffffffu, m, off);
    if (lane >= off) {
      c += cc;
      m = max(m, mm);
    }
  }
  ex_c = __shfl_up_sync(0xffffffffu, c, 1);
  ex_m = __shfl_up_sync(0xffffffffu, m, 1);
  if (lane == 0) {
    ex_c = 0;
    ex_m = 0;
  }
  if (NT == 32) {
    tot_c = __shfl_sync(0xffffffffu, c, 31);
    tot_m = __shfl_sync(0xffffffffu, m, 31);
    return;
  }
  __syncthreads();  // s_c / s_m may still be read from the previous call
  if (lane == 31) {
    s_c[warp] = c;
    s_m[warp] = m;
  }
  __syncthreads();
  int bc = 0, bm = 0, tc = 0, tm = 0;
  for (int w = 0; w < NT / 32; ++w) {
    const int wc = s_c[w], wm = s_m[w];
    if (w < warp) {
      bc += wc;
      bm = max(bm, wm);
    }
    tc += wc;
    tm = max(tm, wm);
  }
  ex_c += bc;
  ex_m = max(ex_m, bm);
  tot_c = tc;
  tot_m = tm;
}

template <int NT>
__device__ __forceinline__ int block_min_int(int v, int* s_v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
  if (NT == 32) return v;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_v[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = 0x7fffffff;
  for (int w = 0; w < NT / 32; ++w) r = min(r, s_v[w]);
  return r;
}

template <int NT>
__global__ void __launch_bounds__(NT)
k_sel_scan(const PairMeta* __restrict__ meta, int Nmax, int H, int k, int method, int max_iteration, int adaptive,
           const int32_t* __restrict__ tab, const int32_t* __restrict__ counts, const int8_t* __restrict__ states,
           pre3_pair_result* __restrict__ res, uint8_t* __restrict__ masks, int mask_stride,
           int32_t* __restrict__ counts_out, int8_t* __restrict__ states_out, SelInfo* __restrict__ info,
           int2* __restrict__ ties, int* __restrict__ tie_total, int32_t* __restrict__ pair_ties) {
  __shared__ int s_a[32], s_b[32], s_base, s_pc, s_pm;
  const int p = blockIdx.x;
  const int tid = threadIdx.x;
  const PairMeta m = meta[p];
  const int N = m.N;
  const int32_t* cnt = counts + (size_t)p * H;
  const int8_t* sts = states + (size_t)p * H;
  uint8_t* mask = masks ? masks + (size_t)p * mask_stride : nullptr;
  SelInfo si;
  si.status = 0;
  si.S_end = 0;
  si.maxc = 0;
  si.n_iter = 0;
  si.first_nonmax = 0x7fffffff;
  si.n_ties = 0;
  if (N < k || N <= 0) {  // get_rand(k, N) errors in the reference (get_rand.m:39-41)
    if (tid == 0) {
      result_init(res + p, 1, 0, N, m.thr);
      si.status = 1;
      info[p] = si;
    }
    for (int i = tid; i < H; i += NT) {
      if (counts_out) counts_out[(size_t)p * H + i] = -1;
      if (states_out) states_out[(size_t)p * H + i] = 0;
    }
    if (mask)
      for (int i = tid; i < mask_stride; i += NT) mask[i] = 0;
    return;
  }
  const int32_t* trow = tab ? tab + m.pad : nullptr;

  // ---- 1. loop control, tile by tile; also yields max cardinality and length(M) at the stop ----
  int S_end = H, maxc = 0, n_iter = 0;
  {
    int car_c = 0, car_m = 0;  // recorded count / max cardinality before the current tile
    bool stopped = false;
    for (int base = 0; base < H; base += NT) {
      const int s = base + tid;
      const bool rec = s < H && !(method == PRE3_METHOD_SVD && sts[s] == -1);
      int ec, em, tot_c, tot_m;
      block_exscan_cm<NT>(rec ? 1 : 0, rec ? cnt[s] : 0, ec, em, tot_c, tot_m, s_a, s_b);
      const int pc = car_c + ec, pm = max(car_m, em);
      int nit = max_iteration;
      if (adaptive && pm >= 5) nit = min(trow[min(pm, N)], max_iteration);
      const bool stop_here = s < H && !(1 + pc < nit);
      const int first = block_min_int<NT>(stop_here ? s : 0x7fffffff, s_a);
      if (first != 0x7fffffff) {
        // values at the stop index = recorded count / max cardinality over s < first
        if (s == first) {
          s_pc = pc;
          s_pm = pm;
        }
        __syncthreads();
        n_iter = s_pc;
        maxc = s_pm;
        S_end = first;
        stopped = true;
        break;
      }
      car_c += tot_c;
      car_m = max(car_m, tot_m);
      if (NT != 32) __syncthreads();
    }
    if (!stopped) {
      n_iter = car_c;
      maxc = car_m;
    }
  }
  if (counts_out || states_out)
    for (int s = tid; s < H; s += NT) {
      const bool consumed = s < S_end;
      const bool rec = consumed && !(method == PRE3_METHOD_SVD && sts[s] == -1);
      if (counts_out) counts_out[(size_t)p * H + s] = rec ? cnt[s] : -1;
      if (states_out) states_out[(size_t)p * H + s] = consumed ? sts[s] : (int8_t)0;
    }
  si.S_end = S_end;
  si.maxc = maxc;
  si.n_iter = n_iter;
  if (n_iter < 1) {  // nothing recorded: M(1).ErrorSum = [] in the reference -> error
    if (tid == 0) {
      result_init(res + p, 2, S_end, N, m.thr);
      si.status = 2;
      info[p] = si;
    }
    if (mask)
      for (int i = tid; i < mask_stride; i += NT) mask[i] = 0;
    return;
  }

  // ---- 2. ties at max cardinality -> global list; first recorded non-tie ------------------------
  int first_nonmax = 0x7fffffff, n_ties = 0;
  for (int base = 0; base < S_end; base += NT) {
    const int s = base + tid;
    const bool rec = s < S_end && !(method == PRE3_METHOD_SVD && sts[s] == -1);
    const bool is_tie = rec && maxc > 0 && cnt[s] == maxc;
    if (rec && !is_tie) first_nonmax = min(first_nonmax, s);
    const unsigned bal = __ballot_sync(0xffffffffu, is_tie);
    int tile_ties;
    int my_off = __popc(bal & ((1u << (tid & 31)) - 1u));
    if (NT == 32) {
      tile_ties = __popc(bal);
    } else {
      __syncthreads();
      if ((tid & 31) == 0) s_a[tid >> 5] = __popc(bal);
      __syncthreads();
      tile_ties = 0;
      for (int w = 0; w < NT / 32; ++w) {
        if (w < (tid >> 5)) my_off += s_a[w];
        tile_ties += s_a[w];
      }
    }
    if (tile_ties > 0) {
      if (tid == 0) s_base = atomicAdd(tie_total, tile_ties);
      if (NT == 32) __syncwarp(); else __syncthreads();
      const int tb = s_base;
      if (is_tie) {
        ties[tb + my_off] = make_int2(p, s);
        pair_ties[(size_t)p * H + n_ties + my_off] = s;  // this pair's ties in ascending order
      }
      if (NT == 32) __syncwarp(); else __syncthreads();
    }
    n_ties += tile_ties;
  }
  first_nonmax = block_min_int<NT>(first_nonmax, s_a);
  if (tid == 0) {
    si.first_nonmax = first_nonmax;
    si.n_ties = n_ties;
    info[p] = si;
  }
}

// The fit of sample set s of pair p if k_eval_pairloop kept it (fc == nullptr: no cache in this call), looked up by a
// GROUP of G lanes (a power of two; all lanes of the warp call it, the lanes of a group with the same p and s): lane `sub`
// of the group tests entries sub, sub + G, ... -- their loads are independent, where a one-thread loop walks up to 24
// cache lines one after the other (ncu source page: 6-10 % of the selection kernels' samples) -- and the first matching
// entry in list order is taken.
template <int G>
__device__ __forceinline__ bool fit_cached_group(const FitCacheEntry* __restrict__ fc, const int32_t* __restrict__ fcn,
                                                 int p, int s, int sub, Rigid& f) {
  if (!fc) return false;
  const int n = min(fcn[p], FIT_CACHE_CAP);
  const FitCacheEntry* e = fc + (size_t)p * FIT_CACHE_CAP;
  int found = 0x7fffffff;
#pragma unroll
  for (int r = 0; r < (FIT_CACHE_CAP + G - 1) / G; ++r) {
    const int j = r * G + sub;
    if (j < n && e[j].h == s) found = min(found, j);
  }
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, off, G));
  if (found == 0x7fffffff) return false;
#pragma unroll
  for (int i = 0; i < 9; ++i) f.R[i] = e[found].Rt[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) f.t[i] = e[found].Rt[9 + i];
  return true;
}

constexpr int TIE_THREADS = 128;

// sum of v over the 32 lanes strictly in lane order, continued from acc (every lane returns it)
__device__ __forceinline__ double warp_ordered_add(double acc, double v) {
#pragma unroll
  for (int l = 0; l < 32; ++l) acc = acc + __shfl_sync(0xffffffffu, v, l);
  return acc;
}

// One WARP per tied hypothesis: minimal fit again (every lane, same result), the residual norms of
// 32 correspondences at a time, summed strictly in index order (non-inliers add an exact +0.0).
__global__ void __launch_bounds__(TIE_THREADS)
k_sel_tie(const PairMeta* __restrict__ meta, const double* __restrict__ Ya, const double* __restrict__ Yb, int Nmax,
          const int32_t* __restrict__ samples, uint64_t seed, uint32_t pair_id0, long long h0, int H, int k,
          int method, const int2* __restrict__ ties, const int* __restrict__ tie_total, int cap,
          double* __restrict__ es_out) {
  const int total = min(*tie_total, cap);
  const int lane = threadIdx.x & 31;
  const int wpb = TIE_THREADS / 32;
  for (int t = blockIdx.x * wpb + (threadIdx.x >> 5); t < total; t += gridDim.x * wpb) {
    const int2 ps = ties[t];
    const int p = ps.x, s = ps.y;
    const PairMeta m = meta[p];
    const int N = m.N;
    const double* ya = Ya + (size_t)p * Nmax * 3;
    const double* yb = Yb + (size_t)p * Nmax * 3;
    int idx[MAX_K];
    load_sample(samples, seed, pair_id0 + (uint32_t)p, h0, s, H, p, N, k, idx);
    Rigid f;
    fit_sample(method, ya, yb, idx, k, f);
    double es = 0.0;
    for (int ib = 0; ib < N; ib += 32) {
      const int i = ib + lane;
      double v = 0.0;
      if (i < N) {
        const double nr = residual_norm(f.R, f.t, ya + 3 * i, yb + 3 * i);
        v = nr < m.thr ? nr : 0.0;  // sum(normResidu(inliers)) in index order (:135)
      }
      es = warp_ordered_add(es, v);
    }
    if (lane == 0) es_out[(size_t)p * H + s] = es;
  }
}

// FOUR lanes per tied hypothesis (small N): the lanes of a quad take every fourth correspondence and the
// ordered sum is rebuilt with width-4 shuffles (four adds per four correspondences, in index order).  Four
// times the warps of the thread-per-tie form for the same ties: the kernel is latency bound (sqrt / dependent
// adds), not throughput bound.
template <int G>
__global__ void __launch_bounds__(TIE_THREADS, G >= 16 ? 3 : 4)
k_sel_tie_quad(const PairMeta* __restrict__ meta, const double* __restrict__ Ya, const double* __restrict__ Yb,
               int Nmax, const int32_t* __restrict__ samples, uint64_t seed, uint32_t pair_id0, long long h0, int H,
               int k, int method, const int2* __restrict__ ties, const int* __restrict__ tie_total, int cap,
               double* __restrict__ es_out, const FitCacheEntry* __restrict__ fcache,
               const int32_t* __restrict__ fcache_n) {
  const int total = min(*tie_total, cap);
  const int sub = threadIdx.x & (G - 1);
  const int quads = (gridDim.x * TIE_THREADS) / G;
  const int rounds = (total + quads - 1) / quads;  // every thread runs the same number of rounds (full-warp shuffles)
  for (int rd = 0; rd < rounds; ++rd) {
    const int t = rd * quads + ((blockIdx.x * TIE_THREADS + threadIdx.x) / G);
    const bool live = t < total;
    if (!__any_sync(0xffffffffu, live)) continue;  // warp-uniform: nothing to do for this warp in this round
    const int2 ps = ties[live ? t : 0];
    const int p = ps.x, s = ps.y;
    const PairMeta m = meta[p];
    const int N = live ? m.N : 0;
    const double* ya = Ya + (size_t)p * Nmax * 3;
    const double* yb = Yb + (size_t)p * Nmax * 3;
    Rigid f;
    if (!fit_cached_group<G>(fcache, fcache_n, p, s, sub, f)) {
      int idx[MAX_K];
      load_sample(samples, seed, pair_id0 + (uint32_t)p, h0, s, H, p, max(m.N, 1), k, idx);
      fit_sample(method, ya, yb, idx, k, f);
    }
    const int Nw = __reduce_max_sync(0xffffffffu, N);
    double es = 0.0;
    for (int ib = 0; ib < Nw; ib += 2 * G) {
      double v0 = 0.0, v1 = 0.0;
      const int i0 = ib + sub, i1 = ib + G + sub;
      {  // the lines the walk reaches three trips from now are asked into L1 (no registers held)
        const int ipf = min(i0 + 6 * G, max(N, 1) - 1);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(ya + 3 * ipf));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(yb + 3 * ipf));
      }
      if (i0 < N) {
        const double nr = residual_norm(f.R, f.t, ya + 3 * i0, yb + 3 * i0);
        v0 = nr < m.thr ? nr : 0.0;  // sum(normResidu(inliers)) in index order (:135); + 0.0 is exact
      }
      if (i1 < N) {
        const double nr = residual_norm(f.R, f.t, ya + 3 * i1, yb + 3 * i1);
        v1 = nr < m.thr ? nr : 0.0;
      }
#pragma unroll
      for (int l = 0; l < G; ++l) es = es + __shfl_sync(0xffffffffu, v0, l, G);
#pragma unroll
      for (int l = 0; l < G; ++l) es = es + __shfl_sync(0xffffffffu, v1, l, G);
    }
    if (live && sub == 0) es_out[(size_t)p * H + s] = es;
  }
}

// NT threads per pair: 32 (one warp, batches of pairs) or 512 (a few large pairs).
// FUSED (ablation, see launch_select): the ErrorSum of the pair's tied hypotheses is computed here, four lanes per tie
// (the work of k_sel_tie_quad); with NT > 32 the pair's correspondences are staged in 48 Nmax bytes of shared memory.
template <int NT, bool FUSED>
__global__ void __launch_bounds__(NT)
k_sel_final(const PairMeta* __restrict__ meta, const double* __restrict__ Ya, const double* __restrict__ Yb, int P,
            int Nmax, const int32_t* __restrict__ samples, uint64_t seed, uint32_t pair_id0, long long h0, int H, int k,
            int method, const int32_t* __restrict__ counts, const int8_t* __restrict__ states,
            const SelInfo* __restrict__ info, const int32_t* __restrict__ pair_ties, const double* __restrict__ es_in,
            pre3_pair_result* __restrict__ res, uint8_t* __restrict__ masks, int mask_stride,
            uint8_t* __restrict__ mask_scratch, const FitCacheEntry* __restrict__ fcache,
            const int32_t* __restrict__ fcache_n) {
  __shared__ double s_es[32];
  __shared__ int s_i[32];
  const int p = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (p >= P) return;
  const SelInfo si = info[p];
  if (si.status != 0) return;  // result already written by k_sel_scan
  const PairMeta m = meta[p];
  const int N = m.N;
  const double* ya = Ya + (size_t)p * Nmax * 3;
  const double* yb = Yb + (size_t)p * Nmax * 3;
  extern __shared__ double s_pts[];
  if (FUSED && NT > 32) {
    double* sa = s_pts;
    double* sb = s_pts + 3 * (size_t)Nmax;
    for (int i = tid; i < 3 * N; i += NT) {
      sa[i] = ya[i];
      sb[i] = yb[i];
    }
    ya = sa;
    yb = sb;
    __syncthreads();
  }
  const int8_t* sts = states + (size_t)p * H;
  uint8_t* mask = masks ? masks + (size_t)p * mask_stride : mask_scratch + (size_t)p * Nmax;
  auto recorded = [&](int s) { return !(method == PRE3_METHOD_SVD && sts[s] == -1); };

  // ---- min ErrorSum over this pair's ties (listed in ascending s), first index on equal sums --------
  double best_es = INFINITY;
  int best_s = 0x7fffffff;
  if (FUSED) {
    constexpr int G = 4;  // lanes per tie (k_sel_tie_quad: 4 measured best)
    const int sub = tid & (G - 1), qid = tid / G, quads = NT / G;
    for (int j0 = 0; j0 < si.n_ties; j0 += quads) {
      const int j = j0 + qid;
      const bool live = j < si.n_ties;
      if (!__any_sync(0xffffffffu, live)) continue;  // warp-uniform
      const int s = pair_ties[(size_t)p * H + (live ? j : 0)];
      int idx[MAX_K];
      load_sample(samples, seed, pair_id0 + (uint32_t)p, h0, s, H, p, max(N, 1), k, idx);
      Rigid f;
      fit_sample(method, ya, yb, idx, k, f);
      double es = 0.0;
      for (int ib = 0; ib < N; ib += 2 * G) {
        double v0 = 0.0, v1 = 0.0;
        const int i0 = ib + sub, i1 = ib + G + sub;
        if (i0 < N) {
          const double nr = residual_norm(f.R, f.t, ya + 3 * i0, yb + 3 * i0);
          v0 = nr < m.thr ? nr : 0.0;  // sum(normResidu(inliers)) in index order (:135); + 0.0 is exact
        }
        if (i1 < N) {
          const double nr = residual_norm(f.R, f.t, ya + 3 * i1, yb + 3 * i1);
          v1 = nr < m.thr ? nr : 0.0;
        }
#pragma unroll
        for (int l = 0; l < G; ++l) es = es + __shfl_sync(0xffffffffu, v0, l, G);
#pragma unroll
        for (int l = 0; l < G; ++l) es = es + __shfl_sync(0xffffffffu, v1, l, G);
      }
      if (live && es < best_es) {  // ascending s within a quad: strict < keeps the first
        best_es = es;
        best_s = s;
      }
    }
  } else {
    for (int j = tid; j < si.n_ties; j += NT) {
      const int s = pair_ties[(size_t)p * H + j];
      const double es = es_in[(size_t)p * H + s];
      if (es < best_es) {  // ascending s within a thread: strict < keeps the first
        best_es = es;
        best_s = s;
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double oe = __shfl_xor_sync(0xffffffffu, best_es, off);
    const int os = __shfl_xor_sync(0xffffffffu, best_s, off);
    if (oe < best_es || (oe == best_es && os < best_s)) {
      best_es = oe;
      best_s = os;
    }
  }
  if (NT > 32) {
    if (lane == 0) {
      s_es[warp] = best_es;
      s_i[warp] = best_s;
    }
    __syncthreads();
    best_es = s_es[0];
    best_s = s_i[0];
    for (int w = 1; w < NT / 32; ++w)
      if (s_es[w] < best_es || (s_es[w] == best_es && s_i[w] < best_s)) {
        best_es = s_es[w];
        best_s = s_i[w];
      }
    __syncthreads();
  }
  int win = best_s;
  const int fn = si.first_nonmax;
  if (best_s == 0x7fffffff) {
    win = fn;  // maxc == 0: every entry of eee1 is 10000 -> I = first recorded
  } else if (fn != 0x7fffffff && (10000.0 < best_es || (10000.0 == best_es && fn < best_s))) {
    win = fn;
  }
  // BestFitIdx: recorded hypotheses up to and including the winner
  int win_iter = 0;
  for (int s = tid; s <= win; s += NT) win_iter += recorded(s) ? 1 : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) win_iter += __shfl_xor_sync(0xffffffffu, win_iter, off);
  if (NT > 32) {
    if (lane == 0) s_i[warp] = win_iter;
    __syncthreads();
    win_iter = 0;
    for (int w = 0; w < NT / 32; ++w) win_iter += s_i[w];
    __syncthreads();
  }

  // ---- winner: hypothesis, mask, ErrorSum; refit on the support set (:186) --------------------------
  Rigid f;
  if (!fit_cached_group<32>(fcache, fcache_n, p, win, lane, f)) {
    int idx[MAX_K];
    load_sample(samples, seed, pair_id0 + (uint32_t)p, h0, win, H, p, N, k, idx);
    fit_sample(method, ya, yb, idx, k, f);  // every thread computes the same fit
  }
  for (int i = tid; i < N; i += 4 * NT) {  // four correspondences per trip: their loads are in flight together
    double ra[4][3], rb[4][3];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int iu = min(i + u * NT, N - 1);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        ra[u][r] = ya[3 * iu + r];
        rb[u][r] = yb[3 * iu + r];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * NT < N) mask[i + u * NT] = residual_norm(f.R, f.t, ra[u], rb[u]) < m.thr ? 1 : 0;
  }
  if (masks)
    for (int i = N + tid; i < mask_stride; i += NT) mask[i] = 0;
  if (NT > 32) __syncthreads(); else __syncwarp();
  if (warp != 0) return;
  double es = best_es;
  if (win != best_s) es = warp_errsum(f.R, f.t, ya, yb, N, m.thr, nullptr, nullptr);
  Rigid rf;
  const int st = warp_refit(method, ya, yb, mask, N, rf);
  if (lane == 0) {
    pre3_pair_result* out = res + p;
    out->status = 0;
    out->state = st;
    out->best_fit = si.maxc;
    out->best_sample = win;
    out->best_iter = win_iter;
    out->n_iter = si.n_iter;
    out->n_consumed = si.S_end;
    out->n_matches = N;
    out->thr = m.thr;
    out->error_sum = es;
    store_colmajor(out->R, rf.R);
    for (int i = 0; i < 3; ++i) out->T[i] = rf.t[i];
    store_colmajor(out->R_hyp, f.R);
    for (int i = 0; i < 3; ++i) out->T_hyp[i] = f.t[i];
  }
}
// ------------------------------------------------------------------------------------------
// Selection of one pair by the block that evaluated it (64 threads): the work of k_sel_scan / k_sel_tie / k_sel_final
// (RANSAC_CALC_VER2.m:165-175 and the refit :186) without leaving the kernel.  The loop control already gave the stop
// index S_end, length(M) = n_iter and the max cardinality maxc.  The selection is latency bound (a handful of fp64
// fits and index-ordered sums per pair); inside the evaluation kernel it overlaps with the scoring of the other pairs
// resident on the SM.
//   ties at max cardinality are visited in ascending order, 16 at a time: FOUR lanes per tie rebuild its ErrorSum
//   (:135, summed strictly in index order; non-inliers add an exact +0.0) with width-4 shuffles;
//   [C, I] = min(eee1): min ErrorSum among the ties, 10000 for every other recorded hypothesis, first index wins;
//   winner: hypothesis again, mask, ErrorSum, least-squares refit on the support set by warp 0.
// scratch: sRt (>= 16 doubles), sCnt / sState (>= 64 ints) of the evaluation phase.
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void pairloop_select(const PairMeta& m, const double* __restrict__ ya,
                                                const double* __restrict__ yb, const int32_t* __restrict__ samples,
                                                uint64_t seed, uint32_t pair_id0, int H, int k, int method, int p,
                                                const int32_t* __restrict__ cnt, const int8_t* __restrict__ sts,
                                                int S_end, int n_iter, int maxc, pre3_pair_result* __restrict__ out,
                                                uint8_t* __restrict__ mask, int mask_stride, double* sRt, int* sA,
                                                int* sB) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = m.N;
  __syncthreads();  // counts / states of the last chunk are in global memory, the scratch arrays are free
  if (n_iter < 1) {  // nothing recorded: M(1).ErrorSum = [] in the reference -> error
    if (tid == 0) result_init(out, 2, S_end, N, m.thr);
    for (int i = tid; i < mask_stride; i += EVP_THREADS) mask[i] = 0;
    return;
  }
  auto recorded = [&](int s) { return !(method == PRE3_METHOD_SVD && sts[s] == -1); };
  // ---- ties in ascending order, 16 per round; running (min ErrorSum, first index) -----------------------------
  double best_es = INFINITY;
  int best_s = 0x7fffffff, first_nonmax = 0x7fffffff;
  int* sTie = sA;          // ties gathered for the current round
  int* sWarpN = sB;        // per-warp tie counts of a 64-wide scan tile
  int nt = 0;              // ties waiting in sTie (same value in every thread)
  const int sub = tid & 3, quad = tid >> 2;
  auto flush = [&](int n) {  // ErrorSum of sTie[0..n), n <= 16: every thread takes part (full-warp shuffles)
    const bool live = quad < n;
    const int s = live ? sTie[quad] : 0;
    Rigid f;
    if (live) {  // (no shuffles inside: idle quads skip the fit)
      int idx[MAX_K];
      load_sample(samples, seed, pair_id0 + (uint32_t)p, 0, s, H, p, N, k, idx);
      fit_sample(method, ya, yb, idx, k, f);
    }
    double es = 0.0;
    for (int ib = 0; ib < N; ib += 8) {
      double v0 = 0.0, v1 = 0.0;
      const int i0 = ib + sub, i1 = ib + 4 + sub;
      if (live && i0 < N) {
        const double nr = residual_norm(f.R, f.t, ya + 3 * i0, yb + 3 * i0);
        v0 = nr < m.thr ? nr : 0.0;  // sum(normResidu(inliers)) in index order (:135); + 0.0 is exact
      }
      if (live && i1 < N) {
        const double nr = residual_norm(f.R, f.t, ya + 3 * i1, yb + 3 * i1);
        v1 = nr < m.thr ? nr : 0.0;
      }
#pragma unroll
      for (int l = 0; l < 4; ++l) es = es + __shfl_sync(0xffffffffu, v0, l, 4);
#pragma unroll
      for (int l = 0; l < 4; ++l) es = es + __shfl_sync(0xffffffffu, v1, l, 4);
    }
    // min over the round (ascending s inside the round: strict < keeps the first), then against the running best
    double e = live ? es : INFINITY;
    int si = live ? s : 0x7fffffff;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double oe = __shfl_xor_sync(0xffffffffu, e, off);
      const int os = __shfl_xor_sync(0xffffffffu, si, off);
      if (oe < e || (oe == e && os < si)) {
        e = oe;
        si = os;
      }
    }
    __syncthreads();  // sTie has been read by every quad
    double* sE = sRt;  // [2] per-warp minima
    if (lane == 0) {
      sE[warp] = e;
      sB[32 + warp] = si;
    }
    __syncthreads();
    const double e0 = sE[0], e1 = sE[1];
    const int i0 = sB[32], i1 = sB[33];
    double re = e0;
    int rs = i0;
    if (e1 < re || (e1 == re && i1 < rs)) {
      re = e1;
      rs = i1;
    }
    if (re < best_es || (re == best_es && rs < best_s)) {  // earlier rounds hold lower indices
      best_es = re;
      best_s = rs;
    }
    __syncthreads();  // the scratch is rewritten by the next scan tile
  };
  for (int base = 0; base < S_end; base += EVP_THREADS) {
    const int s = base + tid;
    const bool rec = s < S_end && recorded(s);
    const bool is_tie = rec && maxc > 0 && cnt[s] == maxc;
    if (rec && !is_tie) first_nonmax = min(first_nonmax, s);
    const unsigned bal = __ballot_sync(0xffffffffu, is_tie);
    if (lane == 0) sWarpN[warp] = __popc(bal);
    __syncthreads();
    const int n0 = sWarpN[0], n1 = sWarpN[1];
    __syncthreads();
    // append this tile's ties (ascending) to the round buffer; flush whenever 16 are waiting
    int off_in_tile = __popc(bal & ((1u << lane) - 1u)) + (warp ? n0 : 0);
    int remaining = n0 + n1, consumed = 0;
    while (remaining > 0) {
      const int take = min(16 - nt, remaining);
      if (is_tie && off_in_tile >= consumed && off_in_tile < consumed + take) sTie[nt + off_in_tile - consumed] = s;
      __syncthreads();
      nt += take;
      consumed += take;
      remaining -= take;
      if (nt == 16) {
        flush(16);
        nt = 0;
      }
    }
  }
  if (nt > 0) flush(nt);
  // first recorded non-tie over the block
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) first_nonmax = min(first_nonmax, __shfl_xor_sync(0xffffffffu, first_nonmax, off));
  if (lane == 0) sB[40 + warp] = first_nonmax;
  __syncthreads();
  const int fn = min(sB[40], sB[41]);
  int win = best_s;
  if (best_s == 0x7fffffff) {
    win = fn;  // maxc == 0: every entry of eee1 is 10000 -> I = first recorded
  } else if (fn != 0x7fffffff && (10000.0 < best_es || (10000.0 == best_es && fn < best_s))) {
    win = fn;
  }
  // BestFitIdx: recorded hypotheses up to and including the winner
  int win_iter = 0;
  for (int s = tid; s <= win; s += EVP_THREADS) win_iter += recorded(s) ? 1 : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) win_iter += __shfl_xor_sync(0xffffffffu, win_iter, off);
  if (lane == 0) sB[44 + warp] = win_iter;
  __syncthreads();
  win_iter = sB[44] + sB[45];
  // ---- winner: hypothesis, mask, ErrorSum; refit on the support set (:186) --------------------------
  for (int i = N + tid; i < mask_stride; i += EVP_THREADS) mask[i] = 0;
  if (warp != 0) return;
  int idx[MAX_K];
  load_sample(samples, seed, pair_id0 + (uint32_t)p, 0, win, H, p, N, k, idx);
  Rigid f;
  fit_sample(method, ya, yb, idx, k, f);  // every lane computes the same fit
  for (int i = lane; i < N; i += 32) mask[i] = residual_norm(f.R, f.t, ya + 3 * i, yb + 3 * i) < m.thr ? 1 : 0;
  __syncwarp();
  double es = best_es;
  if (win != best_s) es = warp_errsum(f.R, f.t, ya, yb, N, m.thr, nullptr, nullptr);
  Rigid rf;
  const int st = warp_refit(method, ya, yb, mask, N, rf);
  if (lane == 0) {
    out->status = 0;
    out->state = st;
    out->best_fit = maxc;
    out->best_sample = win;
    out->best_iter = win_iter;
    out->n_iter = n_iter;
    out->n_consumed = S_end;
    out->n_matches = N;
    out->thr = m.thr;
    out->error_sum = es;
    store_colmajor(out->R, rf.R);
    for (int i = 0; i < 3; ++i) out->T[i] = rf.t[i];
    store_colmajor(out->R_hyp, f.R);
    for (int i = 0; i < 3; ++i) out->T_hyp[i] = f.t[i];
  }
}

// One block per PAIR, adaptive stop: the block walks the pair's sample sets in chunks of EVP_THREADS, and after every
// chunk replays the reference's loop control (RANSAC_CALC_VER2.m:86, :97-99, :137-140) over the chunk's cardinalities
// with the running (recorded hypotheses, max cardinality) carried from the chunks before; it ends with the chunk in
// which the reference's loop ends.  No waves, no per-wave stop kernel, and the work follows the reference's own
// iteration count to within one chunk (the wave schedule of round 1 evaluated 1.34x the sample sets the loop needs).
// Entries beyond the last chunk are never written: k_sel_scan finds the same stop index and reads nothing past it.
// SELECT: the block also runs the selection of its pair (pairloop_select below) -- no selection kernels.
// NT: threads per block = sample sets per chunk (64; 128 / 256 when a GPU holds fewer pairs than it has block slots:
// fewer, wider chunks shorten a pair's critical path at the price of evaluating sets the reference loop would not reach)
template <int K, int MODE, bool SELECT, int MINB = 8, int NT = EVP_THREADS>
__global__ void __launch_bounds__(NT, MINB)
k_eval_pairloop(const PairMeta* __restrict__ meta, const double* __restrict__ Ya, const double* __restrict__ Yb,
                const float4* __restrict__ Ya4, const float4* __restrict__ Yb4, int Nmax,
                const int32_t* __restrict__ samples, uint64_t seed, uint32_t pair_id0, int H, int method,
                int max_iteration, const int32_t* __restrict__ tab, int32_t* __restrict__ counts,
                int8_t* __restrict__ states, int32_t* __restrict__ evaluated, pre3_pair_result* __restrict__ res,
                uint8_t* __restrict__ masks, int mask_stride, uint8_t* __restrict__ mask_scratch,
                FitCacheEntry* __restrict__ fcache, int32_t* __restrict__ fcache_n) {
  __shared__ __align__(16) float sM[6][EVP_TILE];
  static_assert(!SELECT || NT == EVP_THREADS, "fused selection is built for 64 threads");
  constexpr int LIST = EVP_LIST * (NT / EVP_THREADS);
  __shared__ double sRt[NT * 12];
  __shared__ int sCnt[NT];
  __shared__ int sState[NT];
  __shared__ uint32_t sList[LIST];
  __shared__ int sListN, sStop, sSel[3], sMaxNow, sCacheN;

  const int p = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const PairMeta m = meta[p];
  const int N = m.N;
  const double* ya = Ya + (size_t)p * Nmax * 3;
  const double* yb = Yb + (size_t)p * Nmax * 3;
  const int32_t* trow = tab + m.pad;
  if (N < K || N <= 0) {  // get_rand(k, N) errors in the reference (get_rand.m:39-41): nothing to evaluate
    if (tid == 0 && evaluated) evaluated[p] = 0;
    if (tid == 0 && fcache_n) fcache_n[p] = 0;
    if (SELECT) {
      if (tid == 0) result_init(res + p, 1, 0, N, m.thr);
      if (masks)
        for (int i = tid; i < mask_stride; i += NT) masks[(size_t)p * mask_stride + i] = 0;
    }
    return;
  }
  const bool one_tile = N <= EVP_TILE;
  if (one_tile) eval_stage<NT, EVP_TILE>(sM, Ya4 + (size_t)p * Nmax, Yb4 + (size_t)p * Nmax, N, (N + 3) & ~3);
  int car_c = 0, car_m = 0;  // recorded hypotheses / max cardinality over the chunks before
  int hdone = 0;
  if (tid == 0) sCacheN = 0;
  for (int hbeg = 0; hbeg < H; hbeg += NT) {
    const int h = hbeg + tid;
    if (tid == 0) sListN = 0;
    const bool valid = h < H;
    EvalHyp hy;
    eval_fit<K, MODE>(m, ya, yb, valid, p, h, samples, seed, pair_id0, 0, H, nullptr, nullptr, &sRt[tid * 12], hy);
    int cnt = 0;
    if (one_tile) {
      __syncthreads();  // staged tile / list counter visible
      cnt = eval_score_tile<EVP_TILE, LIST>(sM, N, 0, hy, tid, sList, &sListN);
    } else {
      for (int base = 0; base < N; base += EVP_TILE) {
        const int tn = min(EVP_TILE, N - base);
        __syncthreads();
        eval_stage<NT, EVP_TILE>(sM, Ya4 + (size_t)p * Nmax + base, Yb4 + (size_t)p * Nmax + base, tn, (tn + 3) & ~3);
        __syncthreads();
        cnt += eval_score_tile<EVP_TILE, LIST>(sM, tn, base, hy, tid, sList, &sListN);
      }
    }
    sCnt[tid] = cnt;
    eval_recheck<MODE, NT, LIST>(m, ya, yb, sRt, sCnt, sList, &sListN, hy.scored, hy.exact_me);
    const int c = hy.scored ? sCnt[tid] : -1;
    if (h < H) {
      counts[(size_t)p * H + h] = c;
      states[(size_t)p * H + h] = (int8_t)hy.state;
    }
    sState[tid] = (h < H && !(method == PRE3_METHOD_SVD && hy.state == -1)) ? 1 : 0;  // recorded by the reference's loop
    sCnt[tid] = c;
    __syncthreads();
    // loop control over this chunk: warp 0, NT / 32 consecutive sample sets per lane, in order
    if (tid < 32) {
      constexpr int E = NT / 32;
      int rr[E], cc[E];
      int pc = 0, pm = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        rr[e] = sState[E * lane + e];
        cc[e] = rr[e] ? sCnt[E * lane + e] : 0;
        pc += rr[e];
        pm = max(pm, cc[e]);
      }
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int c2 = __shfl_up_sync(0xffffffffu, pc, off);
        const int mm = __shfl_up_sync(0xffffffffu, pm, off);
        if (lane >= off) {
          pc += c2;
          pm = max(pm, mm);
        }
      }
      const int tot_c = __shfl_sync(0xffffffffu, pc, 31), tot_m = __shfl_sync(0xffffffffu, pm, 31);
      pc = __shfl_up_sync(0xffffffffu, pc, 1);
      pm = __shfl_up_sync(0xffffffffu, pm, 1);
      if (lane == 0) {
        pc = 0;
        pm = 0;
      }
      pc += car_c;
      pm = max(pm, car_m);
      int my_stop = 0x7fffffff, stop_pc = 0, stop_pm = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int sl = E * lane + e;
        if (hbeg + sl < H && my_stop == 0x7fffffff) {
          int nit = max_iteration;
          if (pm >= 5) nit = min(trow[min(pm, N)], max_iteration);
          if (!(1 + pc < nit)) {
            my_stop = hbeg + sl;  // the reference's loop condition fails here: recorded count / max cardinality so far
            stop_pc = pc;
            stop_pm = pm;
          }
          if (rr[e]) {
            ++pc;
            pm = max(pm, cc[e]);
          }
        }
      }
      int first = my_stop;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, off));
      car_c += tot_c;
      car_m = max(car_m, tot_m);
      if (first != 0x7fffffff) {
        if (my_stop == first) {
          sSel[0] = first;
          sSel[1] = stop_pc;
          sSel[2] = stop_pm;
        }
      } else if (lane == 0 && hbeg + NT >= H) {  // ran through every sample set
        sSel[0] = H;
        sSel[1] = car_c;
        sSel[2] = car_m;
      }
      if (lane == 0) {
        sStop = first;
        sMaxNow = car_m;
      }
    }
    __syncthreads();
    // a hypothesis that holds the running maximum may tie at the final one: keep its fit for the selection
    if (fcache && sState[tid] && sMaxNow > 0 && sCnt[tid] == sMaxNow) {
      const int slot = atomicAdd(&sCacheN, 1);
      if (slot < FIT_CACHE_CAP) {
        FitCacheEntry* e = fcache + (size_t)p * FIT_CACHE_CAP + slot;
#pragma unroll
        for (int i = 0; i < 12; ++i) e->Rt[i] = sRt[tid * 12 + i];
        e->h = hbeg + tid;
        e->c = sMaxNow;
      }
    }
    hdone = min(H, hbeg + NT);
    if (sStop != 0x7fffffff) break;
  }
  if (tid == 0 && evaluated) evaluated[p] = hdone;
  if (fcache_n) {
    __syncthreads();
    if (tid == 0) fcache_n[p] = min(sCacheN, FIT_CACHE_CAP);
  }
  if constexpr (SELECT)
    pairloop_select(m, ya, yb, samples, seed, pair_id0, H, K, method, p, counts + (size_t)p * H, states + (size_t)p * H,
                    sSel[0], sSel[1], sSel[2], res + p,
                    masks ? masks + (size_t)p * mask_stride : mask_scratch + (size_t)p * Nmax, masks ? mask_stride : 0,
                    sRt, sCnt, sState);
}

// ------------------------------------------------------------------------------------------
// stage-wise kernels
// ------------------------------------------------------------------------------------------
__global__ void k_fit_only(const double* __restrict__ Ya, const double* __restrict__ Yb, int N,
                           const int32_t* __restrict__ samples, int k, int H, int method,
                           double* __restrict__ R, double* __restrict__ T, int32_t* __restrict__ state) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  int idx[MAX_K];
  for (int i = 0; i < k; ++i) idx[i] = min(max(samples[(size_t)h * k + i], 0), N - 1);
  Rigid f;
  const int st = fit_sample(method, Ya, Yb, idx, k, f);
  store_colmajor(R + (size_t)h * 9, f.R);
  for (int i = 0; i < 3; ++i) T[(size_t)h * 3 + i] = f.t[i];
  state[h] = st;
}

// exact fp64 scoring of given hypotheses: one warp per hypothesis; errsum in index order
__global__ void __launch_bounds__(256) k_score_exact(const double* __restrict__ R, const double* __restrict__ T,
                                                      int H, const double* __restrict__ Ya,
                                                      const double* __restrict__ Yb, int N, double thr,
                                                      int32_t* __restrict__ count, double* __restrict__ errsum,
                                                      uint8_t* __restrict__ mask) {
  const int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (h >= H) return;
  double Rr[9], t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Rr[3 * r + c] = R[(size_t)h * 9 + 3 * c + r];
  for (int i = 0; i < 3; ++i) t[i] = T[(size_t)h * 3 + i];
  double es = 0.0;
  int c = 0;
  for (int ib = 0; ib < N; ib += 32) {
    const int i = ib + lane;
    double nr = 0.0;
    bool in = false;
    if (i < N) {
      nr = residual_norm(Rr, t, Ya + 3 * i, Yb + 3 * i);
      in = nr < thr;
      if (mask) mask[(size_t)h * N + i] = in ? 1 : 0;
    }
    unsigned inb = __ballot_sync(0xffffffffu, in);
    c += __popc(inb);
    while (inb) {
      const int li = __ffs(inb) - 1;
      inb &= inb - 1;
      es = es + __shfl_sync(0xffffffffu, nr, li);
    }
  }
  if (lane == 0) {
    if (count) count[h] = c;
    if (errsum) errsum[h] = es;
  }
}

// full-set fits for the find_transform_matrix / absoluteOrientationQuaternion entry points:
// one thread, reference summation order.  out: R(9, column-major), T(3), state, s, err.
__global__ void k_fit_all(const double* __restrict__ p1, const double* __restrict__ p2, int n, int method,
                          int do_scale, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Rigid f;
  int st;
  double s = 1.0, err = 0.0;
  if (method == PRE3_METHOD_SVD) {
    // [rot,trans,state] = find_transform_matrix(pset1 = p1, pset2 = p2)
    auto get = [&](int i, double* a, double* b) {
      for (int r = 0; r < 3; ++r) {
        a[r] = p1[3 * i + r];
        b[r] = p2[3 * i + r];
      }
    };
    st = fit_kabsch<0>(n, get, f);
  } else {
    // [s,R,T,err] = absoluteOrientationQuaternion(A = p1, B = p2, doScale): B ~ s*R*A + T.
    // fit_horn treats its "yb" slot as A and its "ya" slot as B.
    auto get = [&](int i, double* a, double* b) {
      for (int r = 0; r < 3; ++r) {
        a[r] = p2[3 * i + r];
        b[r] = p1[3 * i + r];
      }
    };
    st = fit_horn<0>(n, get, f);
    double Ca[3] = {0, 0, 0}, Cb[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int r = 0; r < 3; ++r) {
        Ca[r] += p1[3 * i + r];
        Cb[r] += p2[3 * i + r];
      }
    for (int r = 0; r < 3; ++r) {
      Ca[r] = Ca[r] / (double)n;
      Cb[r] = Cb[r] / (double)n;
    }
    if (do_scale) {  // absoluteOrientationQuaternion.m:106-112
      double sa = 0.0, sb = 0.0;
      for (int i = 0; i < n; ++i) {
        double an[3], bn[3], ran[3];
        for (int r = 0; r < 3; ++r) {
          an[r] = p1[3 * i + r] - Ca[r];
          bn[r] = p2[3 * i + r] - Cb[r];
        }
        for (int r = 0; r < 3; ++r) ran[r] = (f.R[3 * r] * an[0] + f.R[3 * r + 1] * an[1]) + f.R[3 * r + 2] * an[2];
        sa = sa + ((bn[0] * ran[0] + bn[1] * ran[1]) + bn[2] * ran[2]);
        sb = sb + ((bn[0] * bn[0] + bn[1] * bn[1]) + bn[2] * bn[2]);
      }
      s = sb / sa;
      for (int r = 0; r < 3; ++r)  // T = Cb - s*R*Ca (:118)
        f.t[r] = Cb[r] - ((s * f.R[3 * r] * Ca[0] + s * f.R[3 * r + 1] * Ca[1]) + s * f.R[3 * r + 2] * Ca[2]);
    }
    for (int i = 0; i < n; ++i) {  // :121-127
      double d[3];
      for (int r = 0; r < 3; ++r)
        d[r] = p2[3 * i + r] - (((s * f.R[3 * r] * p1[3 * i] + s * f.R[3 * r + 1] * p1[3 * i + 1]) +
                                 s * f.R[3 * r + 2] * p1[3 * i + 2]) + f.t[r]);
      err = err + sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    }
  }
  store_colmajor(out, f.R);
  for (int i = 0; i < 3; ++i) out[9 + i] = f.t[i];
  out[12] = (double)st;
  out[13] = s;
  out[14] = err;
}

// local best of a hypothesis block for the multi-GPU split: key = (count<<32) | ~id
__global__ void __launch_bounds__(256) k_block_best(const int32_t* __restrict__ counts,
                                                     const int8_t* __restrict__ states, int H, long long h0,
                                                     int method, unsigned long long* __restrict__ key) {
  unsigned long long best = 0ull;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < H; s += gridDim.x * blockDim.x) {
    const bool rec = !(method == PRE3_METHOD_SVD && states[s] == -1);
    if (rec && counts[s] >= 0) {
      const unsigned long long kk = ((unsigned long long)(uint32_t)counts[s] << 32) |
                                    (unsigned long long)(0xFFFFFFFFu - (uint32_t)(h0 + s));
      best = kk > best ? kk : best;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
    best = o > best ? o : best;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(key, best);
}

// ErrorSum of one hypothesis (the local winner) for the reference-exact all-gather mode
__global__ void k_key_errsum(const PairMeta* __restrict__ meta, const double* __restrict__ Ya,
                             const double* __restrict__ Yb, const int32_t* __restrict__ samples, uint64_t seed,
                             uint32_t pair_id0, long long h0, int H, int k, int method,
                             const unsigned long long* __restrict__ key, double* __restrict__ errsum) {
  const int lane = threadIdx.x & 31;
  const PairMeta m = meta[0];
  const unsigned long long kk = *key;
  if (kk == 0ull) {
    if (lane == 0) *errsum = INFINITY;
    return;
  }
  const long long gid = (long long)(0xFFFFFFFFu - (uint32_t)(kk & 0xFFFFFFFFull));
  const int s = (int)(gid - h0);
  int idx[MAX_K];
  load_sample(samples, seed, pair_id0, h0, s, H, 0, m.N, k, idx);
  Rigid f;
  fit_sample(method, Ya, Yb, idx, k, f);
  double es = 0.0;
  for (int ib = 0; ib < m.N; ib += 32) {
    const int i = ib + lane;
    double nr = 0.0;
    bool in = false;
    if (i < m.N) {
      nr = residual_norm(f.R, f.t, Ya + 3 * i, Yb + 3 * i);
      in = nr < m.thr;
    }
    unsigned inb = __ballot_sync(0xffffffffu, in);
    while (inb) {
      const int li = __ffs(inb) - 1;
      inb &= inb - 1;
      es = es + __shfl_sync(0xffffffffu, nr, li);
    }
  }
  if (lane == 0) *errsum = es;
}

// winner known (global id): hypothesis, mask, ErrorSum, refit -- the tail of k_select
__global__ void __launch_bounds__(SEL_THREADS)
k_finish(const PairMeta* __restrict__ meta, const double* __restrict__ Ya, const double* __restrict__ Yb,
         const int32_t* __restrict__ sample_of_winner, uint64_t seed, uint32_t pair_id0, long long winner_id,
         int k, int method, pre3_pair_result* __restrict__ res, uint8_t* __restrict__ mask,
         const unsigned long long* __restrict__ dkey, long long h0, int Hloc, const int32_t* __restrict__ block_samples) {
  __shared__ SelShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PairMeta m = meta[0];
  const int N = m.N;
  if (dkey) {
    // winner = the exchanged key (max count, lowest id); only the rank whose block holds it does the work, the others
    // zero their record and mask (a SUM all-reduce then delivers the owner's)
    const unsigned long long kk = *dkey;
    winner_id = (long long)(0xFFFFFFFFu - (uint32_t)(kk & 0xFFFFFFFFull));
    if (kk == 0ull || winner_id < h0 || winner_id >= h0 + Hloc) {
      unsigned char* rb = reinterpret_cast<unsigned char*>(res);
      for (int i = tid; i < (int)sizeof(pre3_pair_result); i += SEL_THREADS) rb[i] = 0;
      if (mask)
        for (int i = tid; i < N; i += SEL_THREADS) mask[i] = 0;
      return;
    }
    sample_of_winner = block_samples ? block_samples + (size_t)(winner_id - h0) * k : nullptr;
  }
  pre3_pair_result out;
  if (tid == 0) {
    int idx[MAX_K];
    if (sample_of_winner)
      for (int i = 0; i < k; ++i) idx[i] = min(max(sample_of_winner[i], 0), N - 1);
    else
      sample_set<0>(seed, pair_id0, (uint32_t)winner_id, N, k, idx);
    Rigid f;
    fit_sample(method, Ya, Yb, idx, k, f);
    for (int i = 0; i < 9; ++i) sh.Rt[i] = f.R[i];
    for (int i = 0; i < 3; ++i) sh.Rt[9 + i] = f.t[i];
  }
  __syncthreads();
  int c = 0;
  for (int i = tid; i < N; i += SEL_THREADS) {
    const bool in = residual_norm(&sh.Rt[0], &sh.Rt[9], Ya + 3 * i, Yb + 3 * i) < m.thr;
    mask[i] = in ? 1 : 0;
    c += in ? 1 : 0;
  }
  const double cs = block_sum((double)c, sh.scratch);
  // ErrorSum = sum(normResidu(inliers)) strictly in index order (:135).  A chain of ~N/2 dependent fp64 adds cannot be
  // parallelised without changing the rounding, but it can be FED in parallel: per chunk of FIN_CHUNK correspondences
  // the block compacts the inliers' residual norms, in order, into shared memory (each thread owns 8 consecutive
  // correspondences; block scan of the counts) and one thread runs the add chain over the compacted list (8000 inliers:
  // 0.16 ms with the ballot / shuffle walk of one warp, ~0.04 ms this way).
  double es = 0.0;
  {
    __shared__ double s_nr[FIN_CHUNK];
    __shared__ int s_wsum[SEL_THREADS / 32], s_total;
    for (int base = 0; base < N; base += FIN_CHUNK) {
      double nr[FIN_CHUNK / SEL_THREADS];
      int cntl = 0;
#pragma unroll
      for (int j = 0; j < FIN_CHUNK / SEL_THREADS; ++j) {
        const int i = base + (FIN_CHUNK / SEL_THREADS) * tid + j;
        nr[j] = -1.0;
        if (i < N && mask[i]) {
          nr[j] = residual_norm(&sh.Rt[0], &sh.Rt[9], Ya + 3 * i, Yb + 3 * i);
          ++cntl;
        }
      }
      int incl = cntl;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int woff = 0;
      for (int w = 0; w < warp; ++w) woff += s_wsum[w];
      if (tid == SEL_THREADS - 1) s_total = woff + incl;
      int pos = woff + incl - cntl;
#pragma unroll
      for (int j = 0; j < FIN_CHUNK / SEL_THREADS; ++j)
        if (nr[j] >= 0.0) s_nr[pos++] = nr[j];
      __syncthreads();
      if (tid == 0) {
        const int tot = s_total;
        for (int q = 0; q < tot; ++q) es = es + s_nr[q];
      }
      __syncthreads();
    }
  }
  Rigid rf;
  const int st = block_refit(method, Ya, Yb, mask, N, sh.scratch, rf);
  if (tid == 0) {
    out.status = 0;
    out.state = st;
    out.best_fit = (int)cs;
    out.best_sample = (int32_t)winner_id;
    out.best_iter = 0;
    out.n_iter = 0;
    out.n_consumed = 0;
    out.n_matches = N;
    out.thr = m.thr;
    out.error_sum = es;
    store_colmajor(out.R, rf.R);
    for (int i = 0; i < 3; ++i) out.T[i] = rf.t[i];
    store_colmajor(out.R_hyp, &sh.Rt[0]);
    for (int i = 0; i < 3; ++i) out.T_hyp[i] = sh.Rt[9 + i];
    res[0] = out;
  }
}

// ---- hypothesis-block split without host round trips (BASELINE config 5) ---------------------------------------
// Every rank: k_threshold -> k_prep(threshold from device memory) -> k_eval(block) -> local key; ONE collective on the
// same stream (all-reduce MAX of the 8-byte key, or all-gather of 16 bytes per rank); then k_finish / k_split_keep read
// the exchanged keys from device memory and only the owner's record survives (the others write zeros, so a SUM
// all-reduce delivers it).  Nothing is read back before the caller wants the record.
__global__ void k_split_pack(const pre3_pair_result* __restrict__ res, long long h0, unsigned long long* __restrict__ key2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (res->status == 0) {
    key2[0] = ((unsigned long long)(uint32_t)res->best_fit << 32) |
              (unsigned long long)(0xFFFFFFFFu - (uint32_t)(h0 + res->best_sample));
    key2[1] = (unsigned long long)__double_as_longlong(res->error_sum);
  } else {
    key2[0] = ~0ull;  // no recorded hypothesis in this block
    key2[1] = (unsigned long long)__double_as_longlong(INFINITY);
  }
}

// reference rule over the ranks' local winners (RANSAC_CALC_VER2.m:165-175): max cardinality, then min ErrorSum, then
// the lowest hypothesis id; the owner keeps its record (best_sample made global), every other rank zeroes its own.
__global__ void __launch_bounds__(256)
k_split_keep(const unsigned long long* __restrict__ gathered, int ws, int rank, long long h0,
             pre3_pair_result* __restrict__ res, uint8_t* __restrict__ mask, int N) {
  int owner = -1;
  unsigned long long bk = 0ull;
  double be = INFINITY;
  for (int r = 0; r < ws; ++r) {
    const unsigned long long k = gathered[2 * r];
    if (k == ~0ull) continue;
    const double e = __longlong_as_double((long long)gathered[2 * r + 1]);
    const uint32_t c = (uint32_t)(k >> 32), bc = (uint32_t)(bk >> 32);
    const uint32_t id = 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull), bid = 0xFFFFFFFFu - (uint32_t)(bk & 0xFFFFFFFFull);
    if (owner < 0 || c > bc || (c == bc && (e < be || (e == be && id < bid)))) {
      owner = r;
      bk = k;
      be = e;
    }
  }
  if (owner == rank) {
    if (threadIdx.x == 0 && blockIdx.x == 0) res->best_sample = (int32_t)(0xFFFFFFFFu - (uint32_t)(bk & 0xFFFFFFFFull));
    return;
  }
  unsigned char* rb = reinterpret_cast<unsigned char*>(res);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)sizeof(pre3_pair_result); i += gridDim.x * blockDim.x) rb[i] = 0;
  if (mask)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) mask[i] = 0;
}

__global__ void __launch_bounds__(256) k_threshold(const double* __restrict__ Yb, int N, double* __restrict__ thr) {
  __shared__ double s_bz[256];
  __shared__ int s_bi[256];
  double bz = INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < N; i += 256) {
    const double z = Yb[3 * i + 2];
    if (z < bz || (z == bz && i < bi)) {
      bz = z;
      bi = i;
    }
  }
  s_bz[threadIdx.x] = bz;
  s_bi[threadIdx.x] = bi;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      const int o = threadIdx.x + off;
      if (s_bz[o] < s_bz[threadIdx.x] || (s_bz[o] == s_bz[threadIdx.x] && s_bi[o] < s_bi[threadIdx.x])) {
        s_bz[threadIdx.x] = s_bz[o];
        s_bi[threadIdx.x] = s_bi[o];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (N <= 0) {
      *thr = 0.0;
    } else {
      const int j = s_bi[0] == 0x7fffffff ? 0 : s_bi[0];
      const double x = Yb[3 * j], y = Yb[3 * j + 1], z = Yb[3 * j + 2];
      *thr = 0.01 * sqrt((x * x + y * y) + z * z);
    }
  }
}


// ------------------------------------------------------------------------------------------
// The code_from_dr_ye variant: M/code_from_dr_ye/vodometry_dr_ye.m:147-220 with
// ransac_dr_ye.m as the loop body.  Specification: oracle/pre3_oracle_dr_ye.c.
//   k_dy_prep    per pair: dist / threshold on the squared distance (ransac_dr_ye.m:20-23,70), fp32 copies
//   k_dy_sample  per hypothesis: the reference's 4-match sampler (:28-48) on a seeded uniform stream
//   k_eval<4,3>  fit + support of every hypothesis (:59-71)
//   k_dy_select  per pair: first maximum of cnum (vodometry_dr_ye.m:184), support set, refit with the
//                1e-14 variant (:211), mean / std of the residual norms (:212-215), nIterationRansac (:216)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dy_prep(const double* __restrict__ Ya, const double* __restrict__ Yb,
                                                 const int32_t* __restrict__ n_corr, int Nmax, int tab_triangular,
                                                 const int32_t* __restrict__ rowoff, PairMeta* __restrict__ meta,
                                                 float4* __restrict__ Ya4, float4* __restrict__ Yb4) {
  const int p = blockIdx.x;
  int N = n_corr ? n_corr[p] : Nmax;
  N = max(0, min(N, Nmax));
  const double* ya = Ya + (size_t)p * Nmax * 3;
  const double* yb = Yb + (size_t)p * Nmax * 3;
  float y1 = 0.f, xm = 0.f;
  double bz = INFINITY;
  int any = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double ax = ya[3 * i], ay = ya[3 * i + 1], az = ya[3 * i + 2];
    const double bx = yb[3 * i], by = yb[3 * i + 1], bzz = yb[3 * i + 2];
    Ya4[(size_t)p * Nmax + i] = make_float4(__double2float_rn(ax), __double2float_rn(ay), __double2float_rn(az), 0.f);
    Yb4[(size_t)p * Nmax + i] = make_float4(__double2float_rn(bx), __double2float_rn(by), __double2float_rn(bzz), 0.f);
    y1 = fmaxf(y1, __double2float_ru(fabs(bx) + fabs(by) + fabs(bzz)));
    xm = fmaxf(xm, __double2float_ru(fmax(fabs(ax), fmax(fabs(ay), fabs(az)))));
    const double nrm = sqrt((bzz * bzz + by * by) + bx * bx);  // :20
    if (nrm > 0.4) {  // :21
      if (!any || bzz < bz) bz = bzz;
      any = 1;
    }
  }
  __shared__ float s_y1[256], s_xm[256];
  __shared__ double s_bz[256];
  __shared__ int s_i[256];
  s_y1[threadIdx.x] = y1;
  s_xm[threadIdx.x] = xm;
  s_bz[threadIdx.x] = bz;
  s_i[threadIdx.x] = any;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) {
      const int o = threadIdx.x + off;
      s_y1[threadIdx.x] = fmaxf(s_y1[threadIdx.x], s_y1[o]);
      s_xm[threadIdx.x] = fmaxf(s_xm[threadIdx.x], s_xm[o]);
      if (s_i[o] && (!s_i[threadIdx.x] || s_bz[o] < s_bz[threadIdx.x])) s_bz[threadIdx.x] = s_bz[o];
      s_i[threadIdx.x] |= s_i[o];
    }
    __syncthreads();
  }
  const double minZ = s_bz[0];
  const int have = s_i[0];
  y1 = s_y1[0];
  xm = s_xm[0];
  __syncthreads();
  // pmZ = find(pset2(3,:) == minZ) over ALL points; the first one (:22-23)
  int first = 0x7fffffff;
  if (have)
    for (int i = threadIdx.x; i < N; i += blockDim.x)
      if (yb[3 * i + 2] == minZ) {
        first = i;
        break;
      }
  s_i[threadIdx.x] = first;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) s_i[threadIdx.x] = min(s_i[threadIdx.x], s_i[threadIdx.x + off]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    PairMeta m;
    const int j = s_i[0];
    if (have && j != 0x7fffffff) {
      const double x = yb[3 * j], y = yb[3 * j + 1], z = yb[3 * j + 2];
      m.thr = 0.001 * sqrt((x * x + y * y) + z * z);  // :23, :70
    } else {
      m.thr = -1.0;  // no point farther than 0.4 m: the reference errors at :22 (status 5)
    }
    m.thr2 = __double2float_rn(m.thr);
    m.y1max = y1;
    m.xmax = xm;
    m.N = N;
    m.pad = tab_triangular ? (int32_t)(((long long)N * (N + 1)) / 2) : (rowoff ? rowoff[N] : 0);
    meta[p] = m;
  }
}

constexpr int DY_MAX_DRAWS = 1000;

__device__ __forceinline__ int dy_draw(uint64_t seed, uint32_t pair, uint32_t hyp, int j, int pnum) {
  const uint64_t x = splitmix64(seed ^ ((uint64_t)pair * 0x9E3779B97F4A7C15ULL) ^
                                ((uint64_t)hyp * 0xD1B54A32D192ED03ULL) ^
                                ((uint64_t)(j + 1) * 0x8CB92BA72F3D8DD7ULL));
  const double u = (double)(x >> 11) * 0x1.0p-53;
  return (int)round((double)(pnum - 1) * u + 1.0) - 1;  // round((pnum-1)*rand+1), 0-based
}

// one thread per (pair, hypothesis): num_rs(1..4) of ransac_dr_ye.m:28-48, 0-based, in draw order
__global__ void __launch_bounds__(256) k_dy_sample(const PairMeta* __restrict__ meta,
                                                   const int32_t* __restrict__ match, int Nmax, uint64_t seed,
                                                   uint32_t pair_id0, int H, int32_t* __restrict__ samples) {
  const int p = blockIdx.y;
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const int N = meta[p].N;
  int4 out = make_int4(0, 0, 0, 0);
  if (N >= 4) {
    const int32_t* mt = match ? match + 2 * (size_t)p * Nmax : nullptr;
    auto M1 = [&](int i) { return mt ? mt[2 * i] : i; };
    auto M2 = [&](int i) { return mt ? mt[2 * i + 1] : i; };
    const uint32_t pr = pair_id0 + (uint32_t)p;
    int j = 0;
    int n1 = dy_draw(seed, pr, h, j++, N), n2 = dy_draw(seed, pr, h, j++, N);
    int n3 = dy_draw(seed, pr, h, j++, N), n4 = dy_draw(seed, pr, h, j++, N);
    // :33-35: all three flags come from the FIRST draws
    bool d1 = M1(n1) == M1(n2) || M2(n1) == M2(n2);
    bool d2 = M1(n1) == M1(n3) || M1(n2) == M1(n3) || M2(n1) == M2(n3) || M2(n2) == M2(n3);
    bool d3 = M1(n1) == M1(n4) || M1(n2) == M2(n4) || M1(n3) == M1(n4) || M2(n1) == M1(n4) || M2(n2) == M2(n4) ||
              M2(n3) == M2(n4);
    while ((n2 == n1 || d1) && j < DY_MAX_DRAWS) {
      n2 = dy_draw(seed, pr, h, j++, N);
      d1 = M1(n1) == M1(n2) || M2(n1) == M2(n2);
    }
    while ((n3 == n1 || n3 == n2 || d2) && j < DY_MAX_DRAWS) {
      n3 = dy_draw(seed, pr, h, j++, N);
      d2 = M1(n1) == M1(n3) || M1(n2) == M1(n3) || M2(n1) == M2(n3) || M2(n2) == M2(n3);
    }
    while ((n4 == n1 || n4 == n2 || n4 == n3 || d3) && j < DY_MAX_DRAWS) {
      n4 = dy_draw(seed, pr, h, j++, N);
      d3 = M1(n1) == M1(n4) || M1(n2) == M2(n4) || M1(n3) == M1(n4) || M2(n1) == M1(n4) || M2(n2) == M2(n4) ||
           M2(n3) == M2(n4);
    }
    out = make_int4(n1, n2, n3, n4);
  }
  reinterpret_cast<int4*>(samples)[(size_t)p * H + h] = out;
}

__device__ __forceinline__ int dy_rst(int pnum, int max_iteration) {  // min(700, nchoosek(pnum,4)) (:162)
  if (pnum < 4) return 0;
  const double c = ((double)pnum * (pnum - 1) / 2.0) * ((double)(pnum - 2) * (pnum - 3) / 12.0);
  return c < (double)max_iteration ? (int)c : max_iteration;
}

struct DySelShared {
  double scratch[SEL_THREADS];
  unsigned long long key[SEL_THREADS / 32];
  double Rt[12];
  int st;
};

__global__ void __launch_bounds__(SEL_THREADS)
k_dy_select(const PairMeta* __restrict__ meta, const double* __restrict__ Ya, const double* __restrict__ Yb, int Nmax,
            const int32_t* __restrict__ samples, int H, int max_iteration, const int32_t* __restrict__ tab,
            const int32_t* __restrict__ counts, pre3_pair_result* __restrict__ res, uint8_t* __restrict__ masks,
            uint8_t* __restrict__ mask_scratch, pre3_dr_ye_stat* __restrict__ stat,
            int32_t* __restrict__ counts_out) {
  __shared__ DySelShared sh;
  const int p = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PairMeta m = meta[p];
  const int N = m.N;
  const double* ya = Ya + (size_t)p * Nmax * 3;
  const double* yb = Yb + (size_t)p * Nmax * 3;
  uint8_t* mask = masks ? masks + (size_t)p * Nmax : mask_scratch + (size_t)p * Nmax;
  const int32_t* cnt = counts + (size_t)p * H;
  const int rst = dy_rst(N, max_iteration);
  const int L = min(rst, H);
  int status = 0;
  if (N < 4) status = 1;
  else if (m.thr < 0.0) status = 5;
  if (counts_out)
    for (int s = tid; s < H; s += SEL_THREADS) counts_out[(size_t)p * H + s] = (status == 0 && s < L) ? cnt[s] : -1;
  if (masks)
    for (int i = tid; i < Nmax; i += SEL_THREADS) mask[i] = 0;
  pre3_dr_ye_stat sd;
  sd.error_mean = sd.error_std = 0.0;
  sd.dist = status == 0 ? m.thr / 0.001 : 0.0;
  sd.n_iteration_ransac = 0;
  sd.n_loops = status == 0 ? L : 0;
  if (status != 0) {
    if (tid == 0) {
      result_init(res + p, status, 0, N, status == 1 ? 0.0 : m.thr);
      if (stat) stat[p] = sd;
    }
    return;
  }
  // first maximum of tmp_cnum(1..L): max over the key (count << 32 | ~index)
  unsigned long long key = 0ull;
  for (int s = tid; s < L; s += SEL_THREADS) {
    const unsigned long long k = ((unsigned long long)(uint32_t)cnt[s] << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)s);
    key = k > key ? k : key;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
    key = o > key ? o : key;
  }
  if (lane == 0) sh.key[warp] = key;
  __syncthreads();
  key = sh.key[0];
  for (int w = 1; w < SEL_THREADS / 32; ++w) key = sh.key[w] > key ? sh.key[w] : key;
  const int maxc = L > 0 ? (int)(key >> 32) : 0;
  const int win = L > 0 ? (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu)) : -1;
  // nIterations after the last update (:175-178) depends on the final maxCNUM only; never updated if it stays 0
  int nit = rst;
  if (maxc > 0) nit = min(rst, tab[m.pad + min(maxc, N)]);
  sd.n_iteration_ransac = nit;
  if (maxc < 3) {  // :187-194 "no consensus found, ransac fails"
    if (tid == 0) {
      result_init(res + p, 4, L, N, m.thr);
      res[p].best_fit = maxc;
      res[p].best_sample = win;
      res[p].best_iter = win + 1;
      res[p].n_iter = nit;
      if (stat) stat[p] = sd;
    }
    return;
  }
  // winner: fit again (every thread, same result), support set
  int idx[4];
  {
    const int32_t* sp = samples + ((size_t)p * H + win) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) idx[i] = min(max(sp[i], 0), N - 1);
  }
  Rigid f;
  {
    auto get = [&](int i, double* a, double* b) {
      const int j = idx[i];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        a[r] = ya[3 * j + r];
        b[r] = yb[3 * j + r];
      }
    };
    fit_kabsch<4>(4, get, f);
  }
  for (int i = tid; i < N; i += SEL_THREADS) mask[i] = residual_sq(f.R, f.t, ya + 3 * i, yb + 3 * i) < m.thr ? 1 : 0;
  __syncthreads();
  // refit with find_transform_matrix_dr_ye (threshold 1e-14, :211)
  Rigid rf;
  const int st = block_refit(PRE3_METHOD_SVD, ya, yb, mask, N, sh.scratch, rf, 0.00000000000001);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) sh.Rt[i] = rf.R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) sh.Rt[9 + i] = rf.t[i];
    sh.st = st;
  }
  __syncthreads();
  // ErrorRANSAC_Norm over the support set: mean, std (n - 1) (:212-215)
  double sum = 0.0;
  for (int i = tid; i < N; i += SEL_THREADS)
    if (mask[i]) sum += residual_norm(sh.Rt, sh.Rt + 9, ya + 3 * i, yb + 3 * i);
  sum = block_sum(sum, sh.scratch);
  const double mean = sum / (double)maxc;
  double ss = 0.0;
  for (int i = tid; i < N; i += SEL_THREADS)
    if (mask[i]) {
      const double d = residual_norm(sh.Rt, sh.Rt + 9, ya + 3 * i, yb + 3 * i) - mean;
      ss += d * d;
    }
  ss = block_sum(ss, sh.scratch);
  if (tid == 0) {
    pre3_pair_result* out = res + p;
    out->status = 0;
    out->state = sh.st;
    out->best_fit = maxc;
    out->best_sample = win;
    out->best_iter = win + 1;
    out->n_iter = nit;
    out->n_consumed = L;
    out->n_matches = N;
    out->thr = m.thr;
    out->error_sum = sum;
    store_colmajor(out->R, sh.Rt);
    for (int i = 0; i < 3; ++i) out->T[i] = sh.Rt[9 + i];
    store_colmajor(out->R_hyp, f.R);
    for (int i = 0; i < 3; ++i) out->T_hyp[i] = f.t[i];
    sd.error_mean = mean;
    sd.error_std = maxc > 1 ? sqrt(ss / (double)(maxc - 1)) : 0.0;
    if (stat) stat[p] = sd;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
size_t ransac_workspace_bytes(int P, int Nmax, int H) {
  size_t b = 0;
  b += align_up(sizeof(PairMeta) * (size_t)P);
  b += 2 * align_up(sizeof(float4) * (size_t)P * Nmax);
  b += align_up(sizeof(int32_t) * (size_t)P * H);
  b += align_up((size_t)P * H);
  b += align_up((size_t)P * Nmax);  // mask scratch
  b += align_up(sizeof(SelInfo) * (size_t)P) + 3 * align_up(8 * (size_t)P * (H > 0 ? H : 1)) + 512;  // selection
  b += align_up(sizeof(int32_t) * (size_t)P);  // stop flags
  b += align_up(sizeof(FitCacheEntry) * (size_t)P * FIT_CACHE_CAP) + align_up(sizeof(int32_t) * (size_t)P);
  return b + 4096;
}

void ransac_carve(pre3_ctx* ctx, RansacBuffers& b, int H) {
  b.meta = ws_take<PairMeta>(ctx, b.P);
  b.Ya4 = ws_take<float4>(ctx, (size_t)b.P * b.Nmax);
  b.Yb4 = ws_take<float4>(ctx, (size_t)b.P * b.Nmax);
  b.counts = ws_take<int32_t>(ctx, (size_t)b.P * H);
  b.states = ws_take<int8_t>(ctx, (size_t)b.P * H);
  b.stop = ws_take<int32_t>(ctx, b.P);
  b.fcache = ws_take<FitCacheEntry>(ctx, (size_t)b.P * FIT_CACHE_CAP);
  b.fcache_n = ws_take<int32_t>(ctx, b.P);
}

// nIterations(card, N) = mult*ceil(log(epsilon)/log(1-(card/N)^k))  (RANSAC_CALC_VER2.m:139, RANSAC_CALC_VER_test.m:102),
// clamped to [0, MaxIteration] (the loop takes the min, :86); computed on the HOST with the libm the CPU checker uses.
static void adaptive_row(const pre3_ransac_opts& o, int mult, int N, int32_t* row) {
  const double le = std::log(0.01);
  for (int c = 0; c <= N; ++c) {
    int32_t v = o.max_iteration;
    if (N > 0 && c >= 1) {
      const double w = (double)c / (double)N;
      const double d = (double)mult * std::ceil(le / std::log(1.0 - std::pow(w, (double)o.k)));
      if (d == d) {
        if (d < (double)o.max_iteration) v = d <= 0.0 ? 0 : (int32_t)d;
      }
    }
    row[c] = v;
  }
}

constexpr int TAB_TRI_MAX = 2048;             // triangular table up to this Nmax: 2.1 M entries, 8.4 MB
constexpr size_t TAB_ROWS_MAX = (size_t)1 << 26;  // entries of a per-call row set (256 MB) before the call is refused

// Table layouts (only row N of a pair is ever read; PairMeta.pad = offset of that row):
//   Nmax <= TAB_TRI_MAX  triangular over every N <= Nmax (row N at N(N+1)/2), cached in the context;
//   larger               one row per DISTINCT N of the call -- n_corr == nullptr: the single row N = Nmax; otherwise the
//                        P counts are read back and `rowoff[N]` maps a pair's N to its row.  (The triangular form needs
//                        (Nmax+1)(Nmax+2)/2 entries: 200 M at the 20 000-correspondence stress shape.)
int ensure_adaptive_table(pre3_ctx* ctx, const pre3_ransac_opts& o, int Nmax, const int32_t* dn_corr, int P,
                          AdaptiveTable* out) {
  out->tab = nullptr;
  out->rowoff = nullptr;
  out->triangular = 0;
  if (!o.adaptive) return PRE3_OK;
  if (Nmax < 0) return fail(ctx, PRE3_ERR_ARG, "negative correspondence count");
  const int mult = o.method == PRE3_METHOD_HORN ? 1 : 5;  // x5: RANSAC_CALC_VER2.m:139, vodometry_dr_ye.m:177
  try {
    if (Nmax <= TAB_TRI_MAX) {
      out->triangular = 1;
      if (!(ctx->d_tab && ctx->tab_k == o.k && ctx->tab_mult == mult && ctx->tab_nmax >= Nmax &&
            ctx->tab_nmax <= TAB_TRI_MAX && ctx->tab_maxit == o.max_iteration)) {
        const size_t total = ((size_t)Nmax + 1) * ((size_t)Nmax + 2) / 2;
        std::vector<int32_t> tab(total);
        for (int N = 0; N <= Nmax; ++N) adaptive_row(o, mult, N, tab.data() + (size_t)N * (N + 1) / 2);
        if (ctx->d_tab) {
          cudaStreamSynchronize(ctx->stream);
          cudaFree(ctx->d_tab);
          ctx->d_tab = nullptr;
        }
        ctx->tab_nmax = -1;
        if (cudaMalloc((void**)&ctx->d_tab, total * sizeof(int32_t)) != cudaSuccess) {
          ctx->d_tab = nullptr;
          cudaGetLastError();
          return fail(ctx, PRE3_ERR_ALLOC, "cudaMalloc of the adaptive-iteration table failed");
        }
        PRE3_CUDA(cudaMemcpyAsync(ctx->d_tab, tab.data(), total * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        PRE3_CUDA(cudaStreamSynchronize(ctx->stream));  // tab goes out of scope
        ctx->tab_k = o.k;
        ctx->tab_mult = mult;
        ctx->tab_nmax = Nmax;
        ctx->tab_maxit = o.max_iteration;
      }
      out->tab = ctx->d_tab;
      return PRE3_OK;
    }
    // rows for the distinct N of this call
    std::vector<int32_t> off((size_t)Nmax + 1, -1);
    std::vector<int32_t> ns;
    if (!dn_corr) {
      ns.push_back(Nmax);
    } else {
      std::vector<int32_t> nc((size_t)std::max(P, 0));
      if (P > 0) {
        PRE3_CUDA(cudaMemcpyAsync(nc.data(), dn_corr, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
        PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
      }
      std::vector<char> seen((size_t)Nmax + 1, 0);
      for (int p = 0; p < P; ++p) {
        const int N = std::max(0, std::min(nc[p], Nmax));
        if (!seen[N]) {
          seen[N] = 1;
          ns.push_back(N);
        }
      }
    }
    size_t total = 0;
    for (int N : ns) {
      if (total + (size_t)N + 1 > TAB_ROWS_MAX)
        return fail(ctx, PRE3_ERR_ARG, "adaptive stop: too many distinct large correspondence counts in one call "
                                       "(limit 2^26 table entries); split the batch or pass adaptive = 0");
      off[N] = (int32_t)total;
      total += (size_t)N + 1;
    }
    std::vector<int32_t> tab(std::max<size_t>(total, 1));
    for (int N : ns) adaptive_row(o, mult, N, tab.data() + off[N]);
    const size_t need = (tab.size() + off.size()) * sizeof(int32_t);
    if (need > ctx->tab_rows_cap) {
      if (ctx->d_tab_rows) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->d_tab_rows);
        ctx->d_tab_rows = nullptr;
        ctx->tab_rows_cap = 0;
      }
      if (cudaMalloc((void**)&ctx->d_tab_rows, need) != cudaSuccess) {
        ctx->d_tab_rows = nullptr;
        cudaGetLastError();
        return fail(ctx, PRE3_ERR_ALLOC, "cudaMalloc of the adaptive-iteration rows failed");
      }
      ctx->tab_rows_cap = need;
    }
    int32_t* d_rows = ctx->d_tab_rows;
    int32_t* d_off = d_rows + tab.size();
    PRE3_CUDA(cudaMemcpyAsync(d_rows, tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    PRE3_CUDA(cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    PRE3_CUDA(cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
    out->tab = d_rows;
    out->rowoff = d_off;
    return PRE3_OK;
  } catch (const std::bad_alloc&) {
    return fail(ctx, PRE3_ERR_ALLOC, "host allocation of the adaptive-iteration table failed");
  }
}

int launch_prep(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, int thr_given, const double* dthr) {
  Span span__(ctx, T_PREP);
  if (b.P <= 0) return PRE3_OK;
  k_prep<<<b.P, 256, 0, ctx->stream>>>(b.Ya, b.Yb, b.n_corr, b.Nmax, thr_given ? PRE3_METHOD_HORN : o.method,
                                       o.distance_threshold, dthr, b.tab.triangular, b.tab.rowoff, b.meta, b.Ya4,
                                       b.Yb4);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

template <int MODE>
static int launch_eval_mode(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, long long h0, int H,
                            int hbeg, int hend, const int32_t* stop, const double* Rin, const double* Tin) {
  Span span__(ctx, T_EVAL);
  if (hend <= hbeg) return PRE3_OK;
  dim3 grid((hend - hbeg + EV_THREADS - 1) / EV_THREADS, b.P);
#define PRE3_EVAL(KK)                                                                                    \
  k_eval<KK, MODE><<<grid, EV_THREADS, 0, ctx->stream>>>(b.meta, b.Ya, b.Yb, b.Ya4, b.Yb4, b.Nmax,       \
                                                         b.samples, o.seed, b.pair_id0, h0, H, hbeg,     \
                                                         hend, stop, Rin, Tin, b.counts, b.states)
  if (MODE == 2) {
    PRE3_EVAL(1);
  } else if (MODE == 3) {
    PRE3_EVAL(4);
  } else {
    switch (o.k) {
      case 3: PRE3_EVAL(3); break;
      case 4: PRE3_EVAL(4); break;
      case 5: PRE3_EVAL(5); break;
      case 6: PRE3_EVAL(6); break;
      case 7: PRE3_EVAL(7); break;
      case 8: PRE3_EVAL(8); break;
      default: return fail(ctx, PRE3_ERR_ARG, "minimal sample size k must be in 3..8");
    }
  }
#undef PRE3_EVAL
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_eval(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, long long h0, int hbeg, int hend,
                const int32_t* stop) {
  if (b.P <= 0 || o.H <= 0) return PRE3_OK;
  if (o.method == PRE3_METHOD_SVD) return launch_eval_mode<0>(ctx, b, o, h0, o.H, hbeg, hend, stop, nullptr, nullptr);
  if (o.method == PRE3_METHOD_HORN) return launch_eval_mode<1>(ctx, b, o, h0, o.H, hbeg, hend, stop, nullptr, nullptr);
  return fail(ctx, PRE3_ERR_ARG, "unknown RANSAC method");
}

// Adaptive stop on: batches of pairs run k_eval_pairloop (one block per pair, the reference's loop control inside the
// kernel, chunks of EVP_THREADS sample sets); a FEW pairs (one pair cannot fill the machine with one block) are
// evaluated in waves over all SMs, with a scan between waves that drops the pairs whose loop has ended.
constexpr int PAIRLOOP_MIN_P = 64;

// threads per pair block of k_eval_pairloop: 64, or 128 / 256 (k = 5, find_transform_matrix) when the batch leaves block
// slots idle (148 SMs x 8 blocks of 64 threads); PRE3_EVP_NT overrides
static int pairloop_threads(const pre3_ransac_opts& o, int P) {
  if (!(o.k == 5 && o.method == PRE3_METHOD_SVD)) return EVP_THREADS;
  static const int forced = getenv("PRE3_EVP_NT") ? atoi(getenv("PRE3_EVP_NT")) : 0;
  if (forced == 64 || forced == 128 || forced == 256) return forced;
  if (P <= 0) return EVP_THREADS;
  if (P <= 296) return 256;
  if (P <= 592) return 128;
  return EVP_THREADS;
}

static bool use_pairloop(const RansacBuffers& b, const pre3_ransac_opts& o) {
  return o.adaptive && b.P >= PAIRLOOP_MIN_P && b.tab.tab != nullptr && o.k >= 3 && o.k <= 8;
}

// Boundaries of the evaluation schedule: ends[i] = number of sample sets evaluated once step i is done (a pair whose
// loop ends inside step i has had ends[i] sets evaluated).  P: pairs in the call (<= 0: a large batch).
int eval_wave_ends(const pre3_ransac_opts& o, int32_t* ends, int cap, int P) {
  const int H = o.H;
  int n = 0;
  if (H <= 0) return 0;
  if (o.adaptive && (P <= 0 || P >= PAIRLOOP_MIN_P)) {
    const int chunk = pairloop_threads(o, P);
    for (int e = chunk; ; e += chunk) {
      if (n < cap) ends[n] = std::min(e, H);
      ++n;
      if (e >= H) break;
    }
    return n;
  }
  if (!o.adaptive || H <= 384) {
    if (n < cap) ends[n] = H;
    return 1;
  }
  // The reference's loop ends after 5*ceil(log(0.01)/log(1-w^k)) recorded hypotheses once a sample whose support is the
  // fraction w of the matches has been seen: ~190 at the SR4000 shapes of the bench (w ~ 0.65, k = 5).  First wave 256,
  // then doubling.
  int beg = 0, width = 256;
  while (beg < H) {
    int end = std::min(H, beg + width);
    if (H - end < 128) end = H;
    if (n < cap) ends[n] = end;
    ++n;
    beg = end;
    width *= 2;
  }
  return n;
}

template <int MODE, bool SELECT>
static int launch_eval_pairloop_mode(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o,
                                     pre3_pair_result* dres, uint8_t* dmasks, uint8_t* scratch) {
  Span span__(ctx, T_EVAL);
#define PRE3_EVALP(KK)                                                                                             \
  k_eval_pairloop<KK, MODE, SELECT><<<b.P, EVP_THREADS, 0, ctx->stream>>>(                                          \
      b.meta, b.Ya, b.Yb, b.Ya4, b.Yb4, b.Nmax, b.samples, o.seed, b.pair_id0, o.H, o.method, o.max_iteration,     \
      b.tab.tab, b.counts, b.states, b.stop, dres, dmasks, b.Nmax, scratch, SELECT ? nullptr : b.fcache,       \
      SELECT ? nullptr : b.fcache_n)
  // PRE3_EVP_MINB: blocks per SM the k = 5 / find_transform_matrix instance is compiled for (occupancy vs spills)
  // (measured at the sequence shape, eval ms per 4096 pairs: 8 -> 0.338, 10 -> 0.351, 12 -> 0.382, 16 -> 0.393: the
  // spills of the fp64 fit cost more than the extra warps bring)
  static const int minb = getenv("PRE3_EVP_MINB") ? atoi(getenv("PRE3_EVP_MINB")) : 8;
  if constexpr (MODE == 0 && !SELECT) {
    const int nt = pairloop_threads(o, pipe_pairs(ctx, b.P));
    if (nt != EVP_THREADS) {  // few pairs per GPU: wider blocks (same registers per thread: 4 / 2 blocks per SM)
#define PRE3_EVALPW(MB, NTT)                                                                                       \
  k_eval_pairloop<5, 0, false, MB, NTT><<<b.P, NTT, 0, ctx->stream>>>(                                              \
      b.meta, b.Ya, b.Yb, b.Ya4, b.Yb4, b.Nmax, b.samples, o.seed, b.pair_id0, o.H, o.method, o.max_iteration,     \
      b.tab.tab, b.counts, b.states, b.stop, dres, dmasks, b.Nmax, scratch, SELECT ? nullptr : b.fcache,       \
      SELECT ? nullptr : b.fcache_n)
      if (nt == 128) PRE3_EVALPW(4, 128);
      else PRE3_EVALPW(2, 256);
#undef PRE3_EVALPW
      count_launch(ctx);
      PRE3_CUDA(cudaGetLastError());
      return PRE3_OK;
    }
  }
  if (o.k == 5 && MODE == 0 && !SELECT && minb != 8) {
#define PRE3_EVALPB(MB)                                                                                            \
  k_eval_pairloop<5, 0, false, MB><<<b.P, EVP_THREADS, 0, ctx->stream>>>(                                          \
      b.meta, b.Ya, b.Yb, b.Ya4, b.Yb4, b.Nmax, b.samples, o.seed, b.pair_id0, o.H, o.method, o.max_iteration,     \
      b.tab.tab, b.counts, b.states, b.stop, dres, dmasks, b.Nmax, scratch, SELECT ? nullptr : b.fcache,       \
      SELECT ? nullptr : b.fcache_n)
    if (minb == 6) PRE3_EVALPB(6);
    else if (minb == 10) PRE3_EVALPB(10);
    else if (minb == 12) PRE3_EVALPB(12);
    else if (minb == 14) PRE3_EVALPB(14);
    else if (minb == 16) PRE3_EVALPB(16);
    else return fail(ctx, PRE3_ERR_ARG, "PRE3_EVP_MINB: 6, 8, 10, 12, 14 or 16");
#undef PRE3_EVALPB
    count_launch(ctx);
    PRE3_CUDA(cudaGetLastError());
    return PRE3_OK;
  }
  switch (o.k) {
    case 3: PRE3_EVALP(3); break;
    case 4: PRE3_EVALP(4); break;
    case 5: PRE3_EVALP(5); break;
    case 6: PRE3_EVALP(6); break;
    case 7: PRE3_EVALP(7); break;
    case 8: PRE3_EVALP(8); break;
    default: return fail(ctx, PRE3_ERR_ARG, "minimal sample size k must be in 3..8");
  }
#undef PRE3_EVALP
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

// evaluation AND selection of a batch of pairs in one launch (adaptive stop, >= PAIRLOOP_MIN_P pairs)
bool ransac_can_fuse_select(const RansacBuffers& b, const pre3_ransac_opts& o) { return use_pairloop(b, o); }

int launch_eval_select_fused(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, pre3_pair_result* dres,
                             uint8_t* dmasks) {
  if (b.P <= 0) return PRE3_OK;
  uint8_t* scratch = dmasks ? nullptr : ws_take<uint8_t>(ctx, (size_t)b.P * b.Nmax);
  if (o.method == PRE3_METHOD_SVD) return launch_eval_pairloop_mode<0, true>(ctx, b, o, dres, dmasks, scratch);
  if (o.method == PRE3_METHOD_HORN) return launch_eval_pairloop_mode<1, true>(ctx, b, o, dres, dmasks, scratch);
  return fail(ctx, PRE3_ERR_ARG, "unknown RANSAC method");
}

int launch_eval_waves(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o) {
  if (b.P <= 0 || o.H <= 0) return PRE3_OK;
  const int H = o.H;
  if (use_pairloop(b, o)) {
    if (o.method == PRE3_METHOD_SVD) return launch_eval_pairloop_mode<0, false>(ctx, b, o, nullptr, nullptr, nullptr);
    if (o.method == PRE3_METHOD_HORN) return launch_eval_pairloop_mode<1, false>(ctx, b, o, nullptr, nullptr, nullptr);
    return fail(ctx, PRE3_ERR_ARG, "unknown RANSAC method");
  }
  int32_t ends[40];
  const int nw = eval_wave_ends(o, ends, 40, b.P);
  if (nw <= 1) return launch_eval(ctx, b, o, 0, 0, H, nullptr);
  PRE3_CUDA(cudaMemsetAsync(b.counts, 0, sizeof(int32_t) * (size_t)b.P * H, ctx->stream));
  PRE3_CUDA(cudaMemsetAsync(b.states, 0, (size_t)b.P * H, ctx->stream));
  PRE3_CUDA(cudaMemsetAsync(b.stop, 0xFF, sizeof(int32_t) * (size_t)b.P, ctx->stream));
  int beg = 0;
  for (int w = 0; w < nw; ++w) {
    const int end = ends[w];
    PRE3_TRY(launch_eval(ctx, b, o, 0, beg, end, b.stop));
    if (end < H) {
      Span span__(ctx, T_SELECT);
      k_stop<<<(b.P + 7) / 8, 256, 0, ctx->stream>>>(b.meta, b.counts, b.states, b.P, H, end, o.method,
                                                     o.max_iteration, b.tab.tab, b.stop);
      count_launch(ctx);
      PRE3_CUDA(cudaGetLastError());
    }
    beg = end;
  }
  return PRE3_OK;
}

int launch_select(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, pre3_pair_result* dres,
                  uint8_t* dmasks, int32_t* dcounts_out, int8_t* dstates_out) {
  Span span__(ctx, T_SELECT);
  if (b.P <= 0) return PRE3_OK;
  const size_t PH = (size_t)b.P * std::max(o.H, 1);
  uint8_t* scratch = ws_take<uint8_t>(ctx, (size_t)b.P * b.Nmax);
  SelInfo* info = ws_take<SelInfo>(ctx, b.P);
  double* es = ws_take<double>(ctx, PH);
  int2* ties = ws_take<int2>(ctx, PH);
  int32_t* pair_ties = ws_take<int32_t>(ctx, PH);
  int* tie_total = ws_take<int>(ctx, 64);
  if (PH >= ((size_t)1 << 31)) return fail(ctx, PRE3_ERR_ARG, "pairs x sample sets must stay below 2^31");
  PRE3_CUDA(cudaMemsetAsync(tie_total, 0, sizeof(int), ctx->stream));
  const int32_t* tab = o.adaptive ? b.tab.tab : nullptr;
  const bool many = b.P >= 64;  // batches of pairs: one warp per pair; few (large) pairs: 1024 threads each
  // fits kept by k_eval_pairloop (same condition as launch_eval_waves; PRE3_FIT_CACHE=0 refits everything)
  static const int use_cache = getenv("PRE3_FIT_CACHE") ? atoi(getenv("PRE3_FIT_CACHE")) : 1;
  const bool cached = use_cache && use_pairloop(b, o) && b.h0 == 0 && b.fcache != nullptr;
  const FitCacheEntry* fc = cached ? b.fcache : nullptr;
  const int32_t* fcn = cached ? b.fcache_n : nullptr;
  if (many)
    k_sel_scan<32><<<b.P, 32, 0, ctx->stream>>>(b.meta, b.Nmax, o.H, o.k, o.method, o.max_iteration, o.adaptive, tab,
                                                b.counts, b.states, dres, dmasks, b.Nmax, dcounts_out, dstates_out,
                                                info, ties, tie_total, pair_ties);
  else
    k_sel_scan<1024><<<b.P, 1024, 0, ctx->stream>>>(b.meta, b.Nmax, o.H, o.k, o.method, o.max_iteration, o.adaptive,
                                                    tab, b.counts, b.states, dres, dmasks, b.Nmax, dcounts_out,
                                                    dstates_out, info, ties, tie_total, pair_ties);
  const size_t tie_warps = TIE_THREADS / 32;
  const int tie_blocks = (int)std::min<size_t>((PH + tie_warps - 1) / tie_warps, (size_t)ctx->sm_count * 16);
  // Ablation (off by default): ties + winner + refit of a pair in ONE block (PRE3_SEL_FUSED=1: a warp per pair, 2: 64
  // threads with the points staged in shared memory).  Measured on the 4096-pair step: select 0.34 / 0.33 ms against
  // 0.216 ms for the three-launch form below -- serialising a pair's tie sums and its winner's mask + refit in one block
  // lengthens the fp64 critical path more than the saved launch and the re-read of the points (L2 hits) cost.
  static const int fuse_ties = getenv("PRE3_SEL_FUSED") ? atoi(getenv("PRE3_SEL_FUSED")) : 0;
  if (many && fuse_ties == 2 && (size_t)b.Nmax * 48 <= 48 * 1024) {
    k_sel_final<64, true><<<b.P, 64, (size_t)b.Nmax * 48, ctx->stream>>>(
        b.meta, b.Ya, b.Yb, b.P, b.Nmax, b.samples, o.seed, b.pair_id0, b.h0, o.H, o.k, o.method, b.counts, b.states, info,
        pair_ties, es, dres, dmasks, b.Nmax, scratch, nullptr, nullptr);
    count_launch(ctx, 2);
    PRE3_CUDA(cudaGetLastError());
    return PRE3_OK;
  }
  if (many && fuse_ties == 1 && b.Nmax <= 2048) {  // one warp per pair, eight ties at a time, points through L1
    k_sel_final<32, true><<<b.P, 32, 0, ctx->stream>>>(
        b.meta, b.Ya, b.Yb, b.P, b.Nmax, b.samples, o.seed, b.pair_id0, b.h0, o.H, o.k, o.method, b.counts, b.states, info,
        pair_ties, es, dres, dmasks, b.Nmax, scratch, nullptr, nullptr);
    count_launch(ctx, 2);
    PRE3_CUDA(cudaGetLastError());
    return PRE3_OK;
  }
  if (b.Nmax <= 2048) {  // thread per tie; the ordered sum of a long residual list wants a warp per tie
    // 4 lanes per tie measured best (select 0.34 -> 0.27 ms per 4096 pairs; 8 lanes 0.28, 16 lanes 0.31)
    // lanes per tie: 4 when the batch fills the machine (4096 pairs x ~8 ties); with fewer pairs per GPU the residual
    // loop of a tie is the critical path and more lanes shorten it (PRE3_TIE_G overrides)
    static const int forced_g = getenv("PRE3_TIE_G") ? atoi(getenv("PRE3_TIE_G")) : 0;
    const int Pfly = pipe_pairs(ctx, b.P);
    int G = Pfly >= 2048 ? 4 : (Pfly >= 768 ? 8 : 16);
    if (forced_g == 4 || forced_g == 8 || forced_g == 16 || forced_g == 32) G = forced_g;
    const int tb = (int)std::min<size_t>(((size_t)G * PH + TIE_THREADS - 1) / TIE_THREADS, (size_t)ctx->sm_count * 16);
#define PRE3_TIE(GG)                                                                                                  \
  k_sel_tie_quad<GG><<<tb, TIE_THREADS, 0, ctx->stream>>>(b.meta, b.Ya, b.Yb, b.Nmax, b.samples, o.seed, b.pair_id0,  \
                                                          b.h0, o.H, o.k, o.method, ties, tie_total, (int)PH, es, fc, fcn)
    if (G == 4) PRE3_TIE(4);
    else if (G == 8) PRE3_TIE(8);
    else if (G == 16) PRE3_TIE(16);
    else PRE3_TIE(32);
#undef PRE3_TIE
  } else {
    k_sel_tie<<<tie_blocks, TIE_THREADS, 0, ctx->stream>>>(b.meta, b.Ya, b.Yb, b.Nmax, b.samples, o.seed, b.pair_id0,
                                                           b.h0, o.H, o.k, o.method, ties, tie_total, (int)PH, es);
  }
  if (many)
    k_sel_final<32, false><<<b.P, 32, 0, ctx->stream>>>(b.meta, b.Ya, b.Yb, b.P, b.Nmax, b.samples, o.seed, b.pair_id0, b.h0,
                                                 o.H, o.k, o.method, b.counts, b.states, info, pair_ties, es, dres,
                                                 dmasks, b.Nmax, scratch, fc, fcn);
  else
    k_sel_final<512, false><<<b.P, 512, 0, ctx->stream>>>(b.meta, b.Ya, b.Yb, b.P, b.Nmax, b.samples, o.seed, b.pair_id0,
                                                     b.h0, o.H, o.k, o.method, b.counts, b.states, info, pair_ties, es,
                                                     dres, dmasks, b.Nmax, scratch, fc, fcn);
  count_launch(ctx, 3);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

// The code_from_dr_ye variant on device buffers.  The selection scratch of the VER2 path (3 x 8 P H bytes) is
// not needed here; its first 16 P H bytes hold the seeded sample sets.
int launch_dr_ye(pre3_ctx* ctx, RansacBuffers& b, const pre3_ransac_opts& o, const int32_t* dmatch,
                 pre3_pair_result* dres, uint8_t* dmasks, pre3_dr_ye_stat* dstat, int32_t* dcounts_out) {
  if (b.P <= 0) return PRE3_OK;
  if (o.k != 4) return fail(ctx, PRE3_ERR_ARG, "the dr_ye variant draws 4 matches per hypothesis (ransac_dr_ye.m:28)");
  const int H = std::max(o.H, 0);
  const size_t PH = (size_t)b.P * std::max(H, 1);
  if (PH >= ((size_t)1 << 31)) return fail(ctx, PRE3_ERR_ARG, "pairs x sample sets must stay below 2^31");
  uint8_t* scratch = ws_take<uint8_t>(ctx, (size_t)b.P * b.Nmax);
  int32_t* gen = ws_take<int32_t>(ctx, 4 * PH);
  pre3_ransac_opts oa = o;
  oa.adaptive = 1;
  PRE3_TRY(ensure_adaptive_table(ctx, oa, b.Nmax, b.n_corr, b.P, &b.tab));
  {
    Span span__(ctx, T_PREP);
    k_dy_prep<<<b.P, 256, 0, ctx->stream>>>(b.Ya, b.Yb, b.n_corr, b.Nmax, b.tab.triangular, b.tab.rowoff, b.meta, b.Ya4,
                                            b.Yb4);
    count_launch(ctx);
    if (!b.samples && H > 0) {
      k_dy_sample<<<dim3((H + 255) / 256, b.P), 256, 0, ctx->stream>>>(b.meta, dmatch, b.Nmax, o.seed, b.pair_id0, H,
                                                                      gen);
      count_launch(ctx);
      b.samples = gen;
    }
    PRE3_CUDA(cudaGetLastError());
  }
  // every pair runs min(700, nchoosek(pnum,4)) <= max_iteration iterations: the sets beyond are never read
  const int Hrun = std::min(H, o.max_iteration);
  if (Hrun > 0) PRE3_TRY(launch_eval_mode<3>(ctx, b, o, 0, H, 0, Hrun, nullptr, nullptr, nullptr));
  {
    Span span__(ctx, T_SELECT);
    k_dy_select<<<b.P, SEL_THREADS, 0, ctx->stream>>>(b.meta, b.Ya, b.Yb, b.Nmax, b.samples, H, o.max_iteration,
                                                      b.tab.tab, b.counts, dres, dmasks, scratch, dstat,
                                                      dcounts_out);
    count_launch(ctx);
    PRE3_CUDA(cudaGetLastError());
  }
  return PRE3_OK;
}

int launch_fit_only(pre3_ctx* ctx, const double* dYa, const double* dYb, int N, const int32_t* dsamples, int k,
                    int H, int method, double* dR, double* dT, int32_t* dstate) {
  Span span__(ctx, T_OTHER);
  if (H <= 0) return PRE3_OK;
  k_fit_only<<<(H + 127) / 128, 128, 0, ctx->stream>>>(dYa, dYb, N, dsamples, k, H, method, dR, dT, dstate);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_score_given(pre3_ctx* ctx, const double* dR, const double* dT, int H, const double* dYa,
                       const double* dYb, int N, double thr, int32_t* dcount, double* derrsum, uint8_t* dmask) {
  Span span__(ctx, T_OTHER);
  if (H <= 0) return PRE3_OK;
  // cardinalities through the production scorer (fp32 + fp64 recheck) ...
  RansacBuffers b{};
  b.Ya = dYa;
  b.Yb = dYb;
  b.n_corr = nullptr;
  b.P = 1;
  b.Nmax = N;
  b.samples = nullptr;
  b.pair_id0 = 0;
  b.meta = ws_take<PairMeta>(ctx, 1);
  b.Ya4 = ws_take<float4>(ctx, (size_t)N);
  b.Yb4 = ws_take<float4>(ctx, (size_t)N);
  b.counts = dcount;
  b.states = nullptr;
  pre3_ransac_opts o{};
  o.method = PRE3_METHOD_HORN;  // threshold taken as given
  o.distance_threshold = thr;
  o.k = 5;
  k_prep<<<1, 256, 0, ctx->stream>>>(b.Ya, b.Yb, nullptr, N, o.method, thr, nullptr, 0, nullptr, b.meta, b.Ya4, b.Yb4);
  count_launch(ctx);
  PRE3_TRY(launch_eval_mode<2>(ctx, b, o, 0, H, 0, H, nullptr, dR, dT));
  // ... ErrorSum and masks through the exact fp64 kernel
  if (derrsum || dmask) {
    k_score_exact<<<(H * 32 + 255) / 256, 256, 0, ctx->stream>>>(dR, dT, H, dYa, dYb, N, thr, nullptr, derrsum, dmask);
    count_launch(ctx);
  }
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_fit_all(pre3_ctx* ctx, const double* dp1, const double* dp2, int n, int method, int do_scale,
                   double* dout) {
  Span span__(ctx, T_OTHER);
  k_fit_all<<<1, 32, 0, ctx->stream>>>(dp1, dp2, n, method, do_scale, dout);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_block_best(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, long long h0, int Hloc,
                      uint64_t* dkey, double* derrsum) {
  Span span__(ctx, T_OTHER);
  PRE3_CUDA(cudaMemsetAsync(dkey, 0, sizeof(uint64_t), ctx->stream));
  const int blocks = std::min(2 * ctx->sm_count, (Hloc + 255) / 256);
  k_block_best<<<std::max(blocks, 1), 256, 0, ctx->stream>>>(b.counts, b.states, Hloc, h0, o.method,
                                                             (unsigned long long*)dkey);
  count_launch(ctx);
  if (derrsum) {
    k_key_errsum<<<1, 32, 0, ctx->stream>>>(b.meta, b.Ya, b.Yb, b.samples, o.seed, b.pair_id0, h0, Hloc, o.k,
                                            o.method, (const unsigned long long*)dkey, derrsum);
    count_launch(ctx);
  }
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_finish(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, long long winner_id,
                  pre3_pair_result* dres, uint8_t* dmask) {
  Span span__(ctx, T_OTHER);
  k_finish<<<1, SEL_THREADS, 0, ctx->stream>>>(b.meta, b.Ya, b.Yb, b.samples, o.seed, b.pair_id0, winner_id, o.k,
                                               o.method, dres, dmask, nullptr, 0, 0, nullptr);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_finish_key(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, const uint64_t* dkey,
                      long long h0, int Hloc, pre3_pair_result* dres, uint8_t* dmask) {
  Span span__(ctx, T_OTHER);
  k_finish<<<1, SEL_THREADS, 0, ctx->stream>>>(b.meta, b.Ya, b.Yb, nullptr, o.seed, b.pair_id0, 0, o.k, o.method, dres,
                                               dmask, (const unsigned long long*)dkey, h0, Hloc, b.samples);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_split_pack(pre3_ctx* ctx, const pre3_pair_result* dres, long long h0, uint64_t* dkey2) {
  Span span__(ctx, T_OTHER);
  k_split_pack<<<1, 32, 0, ctx->stream>>>(dres, h0, (unsigned long long*)dkey2);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_split_keep(pre3_ctx* ctx, const uint64_t* dgathered, int ws, int rank, long long h0, pre3_pair_result* dres,
                      uint8_t* dmask, int N) {
  Span span__(ctx, T_OTHER);
  k_split_keep<<<dmask ? 8 : 1, 256, 0, ctx->stream>>>((const unsigned long long*)dgathered, ws, rank, h0, dres, dmask, N);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_threshold(pre3_ctx* ctx, const double* dYb, int N, double* dthr) {
  Span span__(ctx, T_OTHER);
  k_threshold<<<1, 256, 0, ctx->stream>>>(dYb, N, dthr);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

}  // namespace pre3
