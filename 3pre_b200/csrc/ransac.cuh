// ransac.cuh -- internal (C++) interface of the RANSAC kernels, used by capi.cu.
#pragma once
#include "common.cuh"

namespace pre3 {

// Per-pair constants produced by k_prep and consumed by the evaluation / selection kernels.
struct PairMeta {
  double thr;    // distance threshold (RANSAC_CALC_VER2.m:69-72 or options)
  float thr2;    // fl32(thr*thr)
  float y1max;   // max_i (|yb_x|+|yb_y|+|yb_z|)   -> fp32 error band of the scorer
  float xmax;    // max_i max_j |ya_j|
  int32_t N;     // correspondences
  int32_t pad;
};

// adaptive-iteration table of one call (ensure_adaptive_table)
struct AdaptiveTable {
  const int32_t* tab = nullptr;     // rows of nIterations(card, N), card = 0..N
  const int32_t* rowoff = nullptr;  // rowoff[N] = offset of row N (per-call row set), nullptr otherwise
  int triangular = 0;               // row N at N(N+1)/2
};

// (R, t) of the hypotheses that held the running maximum cardinality when k_eval_pairloop evaluated them: the only ones
// that can tie at the final maximum.  The selection reads them instead of fitting the sample again.
constexpr int FIT_CACHE_CAP = 24;
struct FitCacheEntry {
  double Rt[12];  // R row-major, t
  int32_t h;      // sample set
  int32_t c;      // its cardinality
};

struct RansacBuffers {
  const double* Ya;      // P x Nmax x 3
  const double* Yb;
  const int32_t* n_corr; // P
  int P, Nmax;
  PairMeta* meta;        // P
  float4* Ya4;           // P x Nmax
  float4* Yb4;
  int32_t* counts;       // P x H
  int8_t* states;        // P x H
  int32_t* stop;         // P: sample sets consumed when the adaptive loop is known to have ended, else -1
  const int32_t* samples;  // P x H x k or nullptr (seeded)
  uint32_t pair_id0;
  long long h0 = 0;      // global id of local sample set 0 (hypothesis-block sharding); selection kernels only
  AdaptiveTable tab;     // set by ensure_adaptive_table when the adaptive stop is on
  FitCacheEntry* fcache = nullptr;  // P x FIT_CACHE_CAP
  int32_t* fcache_n = nullptr;      // P: valid entries (written by k_eval_pairloop only)
};

size_t ransac_workspace_bytes(int P, int Nmax, int H);
// carve meta / float4 copies / counts / states out of the context arena
void ransac_carve(pre3_ctx* ctx, RansacBuffers& b, int H);

// thr_given != 0: use o.distance_threshold whatever the method (hypothesis-block entry points)
// dthr != nullptr: the threshold is read from device memory (stream-ordered hypothesis-block split)
int launch_prep(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, int thr_given, const double* dthr = nullptr);
// sample sets [hbeg, hend) of every pair (global hypothesis id = h0 + index); pairs with stop[p] >= 0 are skipped
int launch_eval(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, long long h0, int hbeg, int hend,
                const int32_t* stop);
int launch_eval_waves(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o);
int eval_wave_ends(const pre3_ransac_opts& o, int32_t* ends, int cap, int P = 0);
// batches of pairs with the adaptive stop: evaluation and selection in ONE launch (one block per pair)
bool ransac_can_fuse_select(const RansacBuffers& b, const pre3_ransac_opts& o);
int launch_eval_select_fused(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, pre3_pair_result* dres,
                             uint8_t* dmasks);
int launch_select(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, pre3_pair_result* dres,
                  uint8_t* dmasks, int32_t* dcounts_out, int8_t* dstates_out);
// M/code_from_dr_ye variant (vodometry_dr_ye.m:147-220); dmatch: P x Nmax x 2 match ids or nullptr; b.samples:
// P x H x 4 explicit sets or nullptr (seeded sampler of ransac_dr_ye.m:28-48)
int launch_dr_ye(pre3_ctx* ctx, RansacBuffers& b, const pre3_ransac_opts& o, const int32_t* dmatch,
                 pre3_pair_result* dres, uint8_t* dmasks, pre3_dr_ye_stat* dstat, int32_t* dcounts_out);
int ensure_adaptive_table(pre3_ctx* ctx, const pre3_ransac_opts& o, int Nmax, const int32_t* dn_corr, int P,
                          AdaptiveTable* out);

// stage-wise entry points
int launch_fit_only(pre3_ctx* ctx, const double* dYa, const double* dYb, int N, const int32_t* dsamples, int k,
                    int H, int method, double* dR, double* dT, int32_t* dstate);
int launch_score_given(pre3_ctx* ctx, const double* dR, const double* dT, int H, const double* dYa,
                       const double* dYb, int N, double thr, int32_t* dcount, double* derrsum, uint8_t* dmask);
int launch_fit_all(pre3_ctx* ctx, const double* dp1, const double* dp2, int n, int method, int do_scale,
                   double* dout /* R9 colmajor, T3, state, s, err */);
int launch_block_best(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, long long h0, int Hloc,
                      uint64_t* dkey, double* derrsum);
int launch_finish(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, long long winner_id,
                  pre3_pair_result* dres, uint8_t* dmask);
int launch_threshold(pre3_ctx* ctx, const double* dYb, int N, double* dthr);
// stream-ordered hypothesis-block split: winner read from a device key / the gathered (key, ErrorSum) pairs
int launch_finish_key(pre3_ctx* ctx, const RansacBuffers& b, const pre3_ransac_opts& o, const uint64_t* dkey,
                      long long h0, int Hloc, pre3_pair_result* dres, uint8_t* dmask);
int launch_split_pack(pre3_ctx* ctx, const pre3_pair_result* dres, long long h0, uint64_t* dkey2);
int launch_split_keep(pre3_ctx* ctx, const uint64_t* dgathered, int ws, int rank, long long h0, pre3_pair_result* dres,
                      uint8_t* dmask, int N);

}  // namespace pre3
