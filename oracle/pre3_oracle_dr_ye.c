/* pre3_oracle_dr_ye.c -- CPU restatement of the `code_from_dr_ye` visual-odometry variant
 * (SURVEY.md 8f rank 1), the variant the live EKF calls (M/fv.m:47,
 * M/find_consistent_sift_matches.m:8 -> M/code_from_dr_ye/Calculate_V_Omega_RANSAC_dr_ye.m:19-22
 * -> vodometry_dr_ye.m -> ransac_dr_ye.m).
 *
 * TEST INFRASTRUCTURE ONLY (see pre3_oracle.c): only tests/, __graft_entry__.smoke() and bench.py's
 * CPU baseline may build or call this file; the product (3pre_b200/csrc) never does.
 *
 * PARITY UNPINNED: the reference holds no golden vectors for this path and MATLAB/Octave are absent,
 * so this restatement is checked against an independent numpy restatement (oracle/ref_numpy.py) and
 * planted-motion properties only.
 *
 * What differs from RANSAC_CALC_VER2 (pre3_oracle.c: orc_ransac):
 *   - the threshold acts on the SQUARED distance: good = d_diff < 0.001*dist (ransac_dr_ye.m:61-70),
 *     dist = norm of the first point whose z equals the minimum z over the points farther than 0.4 m
 *     (:20-23);
 *   - minimal sample = 4 matches drawn with round((pnum-1)*rand+1) and re-drawn while two of them
 *     coincide or share a feature (:28-48, including the reference's mixed-row comparisons and its
 *     use of duplicate flags computed before earlier slots were re-drawn); used in DRAW order;
 *   - every hypothesis is scored, whatever the state of its fit (rot = H, trans = 0 on failure);
 *   - min(700, nchoosek(pnum,4)) iterations, ALL executed: the loop is `for i=1:min(rst,nIterations)`
 *     (vodometry_dr_ye.m:165) and MATLAB evaluates a for-range once, so the nIterations update
 *     (:175-178) only changes the reported RANSAC_STAT.nIterationRansac (:216);
 *   - selection = first maximum of cnum (:184); fewer than 3 supporters = failure (:187-194);
 *   - refit with find_transform_matrix_dr_ye (threshold 1e-14, :211), then mean / std of the residual
 *     norms over the support set (:212-215).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

int orc_find_transform_thr(const double *pset1, const double *pset2, const int32_t *idx, int pnum,
                           double threshold, double *rot, double *trans);

static uint64_t dy_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

/* Stand-in for MATLAB's global rand stream (not reproducible outside MATLAB): draw number j of
 * hypothesis (pair, hyp).  SPEC shared with 3pre_b200/csrc/ransac.cu (dy_draw):
 *   x = splitmix64(seed ^ pair*0x9E3779B97F4A7C15 ^ hyp*0xD1B54A32D192ED03 ^ (j+1)*0x8CB92BA72F3D8DD7)
 *   u = (x >> 11) * 2^-53;  num = round((pnum-1)*u + 1)  (ransac_dr_ye.m:30), returned 0-based. */
static int dy_draw(uint64_t seed, uint32_t pair, uint32_t hyp, int j, int pnum) {
  const uint64_t x = dy_splitmix64(seed ^ ((uint64_t)pair * 0x9E3779B97F4A7C15ULL) ^
                                   ((uint64_t)hyp * 0xD1B54A32D192ED03ULL) ^
                                   ((uint64_t)(j + 1) * 0x8CB92BA72F3D8DD7ULL));
  const double u = (double)(x >> 11) * 0x1.0p-53;
  return (int)round((double)(pnum - 1) * u + 1.0) - 1;
}

#define DY_MAX_DRAWS 1000 /* the reference re-draws for ever; we give up and keep the last draw */

typedef struct {
  uint64_t seed;
  uint32_t pair, hyp;
  const double *stream; /* recorded uniforms (tests) or NULL (seeded) */
  int n_stream;
  int j;
} dy_rng;

static int dy_next(dy_rng *g, int pnum) {
  if (g->stream) {
    const double u = g->j < g->n_stream ? g->stream[g->j] : 0.0;
    g->j++;
    return (int)round((double)(pnum - 1) * u + 1.0) - 1;
  }
  return dy_draw(g->seed, g->pair, g->hyp, g->j++, pnum);
}

/* ransac_dr_ye.m:28-48.  match: 2 x pnum (column i = [k1;k2]) or NULL (match(:,i) = [i;i]).
 * out[4]: 0-based match indices in draw order. */
static void dy_sample_core(dy_rng *g, const int32_t *match, int pnum, int32_t *out) {
#define M1(i) (match ? match[2 * (size_t)(i)] : (int32_t)(i))
#define M2(i) (match ? match[2 * (size_t)(i) + 1] : (int32_t)(i))
  const int j0 = g->j;
  int n1 = dy_next(g, pnum);
  int n2 = dy_next(g, pnum);
  int n3 = dy_next(g, pnum);
  int n4 = dy_next(g, pnum);
  /* :33-35 -- all three flags are formed from the FIRST draws */
  int dup1 = (M1(n1) == M1(n2)) || (M2(n1) == M2(n2));
  int dup2 = (M1(n1) == M1(n3)) || (M1(n2) == M1(n3)) || (M2(n1) == M2(n3)) || (M2(n2) == M2(n3));
  int dup3 = (M1(n1) == M1(n4)) || (M1(n2) == M2(n4)) || (M1(n3) == M1(n4)) || (M2(n1) == M1(n4)) ||
             (M2(n2) == M2(n4)) || (M2(n3) == M2(n4));
  while (((n2 == n1) || dup1) && g->j - j0 < DY_MAX_DRAWS) { /* :37-40 */
    n2 = dy_next(g, pnum);
    dup1 = (M1(n1) == M1(n2)) || (M2(n1) == M2(n2));
  }
  while (((n3 == n1) || (n3 == n2) || dup2) && g->j - j0 < DY_MAX_DRAWS) { /* :41-44 */
    n3 = dy_next(g, pnum);
    dup2 = (M1(n1) == M1(n3)) || (M1(n2) == M1(n3)) || (M2(n1) == M2(n3)) || (M2(n2) == M2(n3));
  }
  while (((n4 == n1) || (n4 == n2) || (n4 == n3) || dup3) && g->j - j0 < DY_MAX_DRAWS) { /* :45-48 */
    n4 = dy_next(g, pnum);
    dup3 = (M1(n1) == M1(n4)) || (M1(n2) == M2(n4)) || (M1(n3) == M1(n4)) || (M2(n1) == M1(n4)) ||
           (M2(n2) == M2(n4)) || (M2(n3) == M2(n4));
  }
  out[0] = n1;
  out[1] = n2;
  out[2] = n3;
  out[3] = n4;
#undef M1
#undef M2
}

ORC_API void orc_dr_ye_sample(uint64_t seed, uint32_t pair, uint32_t hyp, const int32_t *match, int pnum,
                              int32_t *out) {
  dy_rng g = {seed, pair, hyp, NULL, 0, 0};
  dy_sample_core(&g, match, pnum, out);
}

/* the same sampler fed from a recorded uniform stream (stands for consecutive rand calls); returns the
 * number of uniforms consumed.  n_sets sample sets are drawn one after the other. */
ORC_API int orc_dr_ye_sample_stream(const double *stream, int n_stream, const int32_t *match, int pnum,
                                    int n_sets, int32_t *out) {
  dy_rng g = {0, 0, 0, stream, n_stream, 0};
  for (int s = 0; s < n_sets; ++s) dy_sample_core(&g, match, pnum, out + 4 * (size_t)s);
  return g.j;
}

/* ransac_dr_ye.m:20-23.  Returns dist, or a negative value when no point lies farther than 0.4 m
 * (min([]) is empty and :22 raises an error in the reference). */
ORC_API double orc_dr_ye_dist(const double *pset2, int N) {
  double minZ = INFINITY;
  int any = 0;
  for (int i = 0; i < N; ++i) {
    const double *y = pset2 + 3 * (size_t)i;
    const double nrm = sqrt((y[2] * y[2] + y[1] * y[1]) + y[0] * y[0]); /* :20: z^2 + y^2 + x^2 */
    if (nrm > 0.4) {
      if (!any || y[2] < minZ) minZ = y[2];
      any = 1;
    }
  }
  if (!any) return -1.0;
  for (int i = 0; i < N; ++i) { /* pmZ = find(pset2(3,:)==minZ): over ALL points, first one */
    const double *y = pset2 + 3 * (size_t)i;
    if (y[2] == minZ) return sqrt((y[0] * y[0] + y[1] * y[1]) + y[2] * y[2]); /* :23 */
  }
  return -1.0;
}

/* d_diff(k) of ransac_dr_ye.m:61-68: pset21 = rs_rot*pset2 (+ rs_trans), sum of squares from 0.0 */
static double dy_d2(const double *R /* row-major */, const double *T, const double *ya, const double *yb) {
  double d = 0.0;
  for (int k = 0; k < 3; ++k) {
    const double y0 = ((R[3 * k] * yb[0] + R[3 * k + 1] * yb[1]) + R[3 * k + 2] * yb[2]) + T[k];
    const double e = y0 - ya[k];
    d = d + e * e;
  }
  return d;
}

ORC_API int orc_dr_ye_score(const double *R, const double *T, const double *Ya, const double *Yb, int N,
                            double thr_sq, uint8_t *mask) {
  int c = 0;
  for (int i = 0; i < N; ++i) {
    const int in = dy_d2(R, T, Ya + 3 * (size_t)i, Yb + 3 * (size_t)i) < thr_sq; /* :70 */
    if (mask) mask[i] = (uint8_t)in;
    c += in;
  }
  return c;
}

/* min(max_iteration, nchoosek(pnum, 4))  (vodometry_dr_ye.m:162, max_iteration = 700) */
ORC_API int orc_dr_ye_rst(int pnum, int max_iteration) {
  if (pnum < 4) return 0;
  const double c = ((double)pnum * (pnum - 1) / 2.0) * ((double)(pnum - 2) * (pnum - 3) / 12.0);
  return c < (double)max_iteration ? (int)c : max_iteration;
}

typedef struct {
  int32_t status;      /* 0 ok; 1 pnum < 4 (:152-160); 4 no consensus, op_num < 3 (:187-194);
                          5 no point farther than 0.4 m (ransac_dr_ye.m:21-22 errors) */
  int32_t state;       /* sta of the refit (RANSAC_STAT.SolutionState) */
  int32_t op_num;      /* support of the winner */
  int32_t best_sample; /* 0-based rs_ind */
  int32_t n_loops;     /* iterations executed: min(rst, H) */
  int32_t n_iteration_ransac; /* RANSAC_STAT.nIterationRansac = min(rst, nIterations) (:216) */
  int32_t pnum;
  int32_t pad;
  double thr;          /* 0.001*dist: the bound on the squared distance */
  double error_sum;    /* sum(ErrorRANSAC_Norm) over the support set after the refit */
  double error_mean;   /* :214 */
  double error_std;    /* :215 (normalised by n-1) */
  double R[9];         /* row-major refit rotation */
  double T[3];
  double R_hyp[9];
  double T_hyp[3];
} orc_dr_ye_result;

/* Ya = pset1 (frame 1 / previous), Yb = pset2 (frame 2 / current): 3 x N column-major.
 * match: 2 x N or NULL.  samples: 4 x H 0-based, draw order, or NULL (seeded).
 * mask_out: N bytes or NULL; counts_out: H or NULL (tmp_cnum; -1 beyond the executed loops). */
ORC_API void orc_vodometry_dr_ye(const double *Ya, const double *Yb, const int32_t *match, int N,
                                 int max_iteration, int H, const int32_t *samples, uint64_t seed, uint32_t pair,
                                 orc_dr_ye_result *res, uint8_t *mask_out, int32_t *counts_out) {
  memset(res, 0, sizeof *res);
  res->pnum = N;
  res->best_sample = -1;
  if (mask_out) memset(mask_out, 0, (size_t)(N > 0 ? N : 0));
  if (counts_out)
    for (int h = 0; h < H; ++h) counts_out[h] = -1;
  if (N < 4) {
    res->status = 1;
    return;
  }
  const double dist = orc_dr_ye_dist(Yb, N);
  if (dist < 0.0) {
    res->status = 5;
    return;
  }
  const double thr = 0.001 * dist;
  res->thr = thr;
  const int rst = orc_dr_ye_rst(N, max_iteration);
  const int L = rst < H ? rst : H;
  res->n_loops = L;
  int maxc = 0, best = -1, best_c = -1;
  double nIterations = (double)rst;
  for (int i = 0; i < L; ++i) {
    int32_t s[4];
    if (samples) {
      for (int d = 0; d < 4; ++d) {
        int v = samples[4 * (size_t)i + d];
        s[d] = v < 0 ? 0 : (v > N - 1 ? N - 1 : v);
      }
    } else {
      orc_dr_ye_sample(seed, pair, (uint32_t)i, match, N, s);
    }
    double R[9], T[3];
    orc_find_transform_thr(Ya, Yb, s, 4, 0.00000000001, R, T); /* ransac_dr_ye.m:59: find_transform_matrix */
    const int cnum = orc_dr_ye_score(R, T, Ya, Yb, N, thr, NULL);
    if (counts_out) counts_out[i] = cnum;
    if (cnum > best_c) { /* [rs_max, rs_ind] = max(tmp_cnum): first maximum */
      best_c = cnum;
      best = i;
    }
    if (cnum > maxc) { /* :175-178 */
      maxc = cnum;
      const double w = (double)maxc / (double)N;
      nIterations = 5.0 * ceil(log(0.01) / log(1.0 - pow(w, 4.0)));
    }
  }
  {
    double r = nIterations < (double)rst ? nIterations : (double)rst; /* min(rst, nIterations); NaN -> rst */
    if (!(r == r)) r = (double)rst;
    res->n_iteration_ransac = r <= 0.0 ? 0 : (int)r;
  }
  res->op_num = best_c < 0 ? 0 : best_c;
  res->best_sample = best;
  if (best < 0 || best_c < 3) {
    res->status = 4;
    return;
  }
  int32_t s[4];
  if (samples) {
    for (int d = 0; d < 4; ++d) {
      int v = samples[4 * (size_t)best + d];
      s[d] = v < 0 ? 0 : (v > N - 1 ? N - 1 : v);
    }
  } else {
    orc_dr_ye_sample(seed, pair, (uint32_t)best, match, N, s);
  }
  orc_find_transform_thr(Ya, Yb, s, 4, 0.00000000001, res->R_hyp, res->T_hyp);
  uint8_t *mask = mask_out ? mask_out : (uint8_t *)malloc((size_t)N);
  orc_dr_ye_score(res->R_hyp, res->T_hyp, Ya, Yb, N, thr, mask);
  int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
  int n = 0;
  for (int i = 0; i < N; ++i)
    if (mask[i]) idx[n++] = i;
  res->state = orc_find_transform_thr(Ya, Yb, idx, n, 0.00000000000001, res->R, res->T); /* :211 */
  double sum = 0.0;
  for (int i = 0; i < n; ++i) { /* :212-213 */
    const double *ya = Ya + 3 * (size_t)idx[i], *yb = Yb + 3 * (size_t)idx[i];
    double e[3];
    for (int k = 0; k < 3; ++k)
      e[k] = (((res->R[3 * k] * yb[0] + res->R[3 * k + 1] * yb[1]) + res->R[3 * k + 2] * yb[2]) + res->T[k]) - ya[k];
    sum = sum + sqrt((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]);
  }
  const double mean = sum / (double)n;
  double ss = 0.0;
  for (int i = 0; i < n; ++i) {
    const double *ya = Ya + 3 * (size_t)idx[i], *yb = Yb + 3 * (size_t)idx[i];
    double e[3];
    for (int k = 0; k < 3; ++k)
      e[k] = (((res->R[3 * k] * yb[0] + res->R[3 * k + 1] * yb[1]) + res->R[3 * k + 2] * yb[2]) + res->T[k]) - ya[k];
    const double d = sqrt((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]) - mean;
    ss = ss + d * d;
  }
  res->error_sum = sum;
  res->error_mean = mean;
  res->error_std = n > 1 ? sqrt(ss / (double)(n - 1)) : 0.0;
  free(idx);
  if (!mask_out) free(mask);
}
