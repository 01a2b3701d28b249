/* pre3.h -- C ABI of libpre3.so: the B200 (sm_100a) drop-in for the frame-to-frame
 * motion-estimation hot path of ahtamjidi/3PRE.
 *
 * Every entry point is what a MEX gateway (or any FFI) for the reference function it
 * cites would bind: plain pointers and sizes, no MATLAB / torch types.  `M/` below is
 * /root/reference/matlab_code/.  INTEGRATION.md shows the MEX stubs.
 *
 * Conventions
 *   - All matrices use MATLAB's COLUMN-MAJOR layout: a "3 x N" point matrix is N
 *     consecutive (x,y,z) triples, "128 x K" descriptors are K consecutive 128-vectors,
 *     a 3x3 rotation R is stored R(1,1),R(2,1),R(3,1),R(1,2),...
 *   - Indices crossing this ABI are 0-BASED (the gateways add 1 for MATLAB).
 *   - Functions return PRE3_OK (0) or a negative error code; pre3_last_error() gives
 *     the text.  Algorithmic failure is NOT an error: it is reported through
 *     State_RANSAC / status fields exactly like the reference does
 *     (M/mex_files/RANSAC_CALCULATION/find_transform_matrix.m:21-42).
 *   - Host-pointer functions are synchronous (they return after the results are in the
 *     caller's buffers).  *_dev functions take DEVICE pointers, enqueue on the context's
 *     stream and do not synchronise.
 *   - There is no CPU fallback: without a CUDA device every compute call fails with
 *     PRE3_ERR_CUDA.
 */
#ifndef PRE3_H
#define PRE3_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PRE3_API __attribute__((visibility("default")))
#else
#define PRE3_API
#endif

#define PRE3_OK 0
#define PRE3_ERR_ARG (-1)   /* bad argument (the reference would mexErrMsgTxt) */
#define PRE3_ERR_CUDA (-2)  /* CUDA runtime failure / no device */
#define PRE3_ERR_ALLOC (-3)
#define PRE3_ERR_CLASS (-4) /* "Unsupported numeric class" (M/sift/siftmatch.c:213-215) */

/* numeric classes accepted by siftmatch (M/sift/siftmatch.c:56-69,208-216) */
#define PRE3_CLASS_DOUBLE 0
#define PRE3_CLASS_SINGLE 1
#define PRE3_CLASS_INT8 2
#define PRE3_CLASS_UINT8 3

/* RANSAC flavours */
#define PRE3_METHOD_SVD 0  /* RANSAC_CALC_VER2.m: find_transform_matrix fit/refit, threshold override :69-72, x5 adaptive rule :139 */
#define PRE3_METHOD_HORN 1 /* RANSAC_CALC_VER_test.m: absoluteOrientationQuaternion fit :71 / refit :152, options threshold, rule :102 */
#define PRE3_METHOD_DR_YE 2 /* code_from_dr_ye/vodometry_dr_ye.m:147-220 + ransac_dr_ye.m: 4 matches per hypothesis in draw order,
                               squared-distance bound 0.001*dist, every hypothesis scored, min(MaxIteration=700, C(n,4))
                               iterations, first maximum wins, refit with find_transform_matrix_dr_ye (1e-14) */

/* matching engines (pre3_set_match_engine) */
#define PRE3_MATCH_AUTO 0   /* tcgen05 proposal GEMM + exact rescore when ND==128, else exact */
#define PRE3_MATCH_EXACT 1  /* exact brute force only */
#define PRE3_MATCH_TC 2     /* force the tensor-core path (error if the shape does not fit) */

typedef struct pre3_ctx pre3_ctx;

typedef struct {
  int32_t method;            /* PRE3_METHOD_* */
  int32_t k;                 /* minimal sample size: 5 (RANSAC_CALC_VER2.m:85), 4, or 3 */
  int32_t max_iteration;     /* options.MaxIteration (M/SIFT_match_save.m:50: 2000) */
  int32_t adaptive;          /* 1: reference adaptive stop (RANSAC_CALC_VER2.m:137-140); 0: fixed H */
  int32_t H;                 /* sample sets available per pair (the loop stops when they run out) */
  int32_t reserved;
  double distance_threshold; /* options.DistanceThreshold; used by PRE3_METHOD_HORN only */
  double ratio;              /* siftmatch thresh, default 1.5 on squared distances (siftmatch.c:146) */
  uint64_t seed;             /* seeded sample sets (used when no explicit sets are supplied) */
} pre3_ransac_opts;

/* One per frame pair.  Mirrors the outputs of RANSAC_CALC_VER2.m:2 plus the counters
 * the reference keeps in RANSAC_STAT (M/code_from_dr_ye/vodometry_dr_ye.m:13-23). */
typedef struct {
  int32_t status;      /* 0 ok; 1 fewer than k correspondences (get_rand would error); 2 no iteration ran */
  int32_t state;       /* State_RANSAC of the refit: 1, 2, 0, -1 */
  int32_t best_fit;    /* BestFit: cardinality of the winning support set */
  int32_t best_sample; /* 0-based index of the winning sample set */
  int32_t best_iter;   /* BestFitIdx in the reference's M(iter) numbering (1-based) */
  int32_t n_iter;      /* length(M): hypotheses recorded */
  int32_t n_consumed;  /* sample sets consumed (recorded + skipped on state -1) */
  int32_t n_matches;   /* N: correspondences that entered RANSAC */
  double thr;          /* distance threshold used */
  double error_sum;    /* M(BestFitIdx).ErrorSum */
  double R[9];         /* refit rotation, column-major; Ya ~ R*Yb + T */
  double T[3];
  double R_hyp[9];     /* winning minimal-sample hypothesis M(BestFitIdx).R (column-major) */
  double T_hyp[3];
} pre3_pair_result;

/* Covariance of the RANSAC pose estimate (M/cov_est_RANSAC_deriv.m:1-244; call site, commented out in the reference,
 * M/mex_files/RANSAC_CALCULATION/RANSAC_CALC_VER2.m:204-206).  State x = [T1 T2 T3 q1 q2 q3 q4] with q = R2q(R);
 * all matrices column-major. */
typedef struct pre3_cov_result {
  double cov[49];    /* res.sm_cov_censi = dA_dz * blkdiag(R, R) * dA_dz'  (:215) */
  double G2tot[49];  /* summed second derivatives d2E/dx2 (:156) */
  double Gtot[7];    /* summed gradient (:153) */
  double dA_dz[42];  /* 7 x 6: G2tot \ [d2E/dx dz_k] (:200) */
  double Etot;       /* summed squared residuals (:150) */
  double s2;         /* Etot / (k - 3) (:216-217) */
  int32_t n;         /* k: support points used */
  int32_t status;    /* 0 ok, 1 empty support set, 2 singular G2tot (a zero pivot) */
} pre3_cov_result;

/* ---- context ------------------------------------------------------------------- */
/* device < 0 selects the current CUDA device.  Lazy one-time initialisation of the
 * stream / workspace arena, as a MEX file would do on first call and undo in mexAtExit
 * (M/mex_files/CorePar_Ver1/codegen/mex/corrcoef_partitioned/corrcoef_partitioned_mex.c:39-57). */
PRE3_API int pre3_create(pre3_ctx **ctx, int device);
PRE3_API void pre3_destroy(pre3_ctx *ctx);
PRE3_API const char *pre3_last_error(const pre3_ctx *ctx);
PRE3_API const char *pre3_version(void);
/* Use an existing cudaStream_t (e.g. torch's current stream) for all later work. */
PRE3_API int pre3_set_stream(pre3_ctx *ctx, void *cuda_stream);
PRE3_API int pre3_set_match_engine(pre3_ctx *ctx, int engine);
PRE3_API int pre3_sync(pre3_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
PRE3_API int64_t pre3_launch_count(const pre3_ctx *ctx);
/* CUDA-graph replay of pre3_pairs_dev / pre3_sequence_dev: the launch sequence of a call signature (pointers, sizes,
 * options) is captured on its second occurrence and replayed afterwards -- one graph launch instead of ~15-20 kernel
 * launches, for loops that push the same device buffers through the path (streaming ring buffers, a sequence sharded
 * over several GPUs where a few hundred pairs per call make the step launch-bound).  Off by default. */
PRE3_API int pre3_set_graphs(pre3_ctx *ctx, int on);
/* Software pipeline of pre3_pairs_dev / pre3_sequence_dev: the pairs of a call are cut into `chunks` blocks whose stages
 * (descriptor conversion | GEMM + rescore + gather | hypothesis evaluation | selection) run on four internal streams,
 * so that kernels bound by different units of the GPU (HBM, tensor cores, FP32 issue, fp64 latency chains) overlap --
 * the batched stand-in for the reference's frame loop (find_consistent_sift_matches.m:22-32), whose iterations are
 * independent.  Results are those of the unchunked call.  chunks = -1: automatic (the default), 0 or 1: off, 2..64:
 * that many chunks (never fewer than 128 pairs each).  The caller's stream sees one call: the internal streams fork
 * from it and join it again. */
PRE3_API int pre3_set_pipeline(pre3_ctx *ctx, int chunks);
/* bytes pre3_pairs has moved over PCIe so far (host -> device, device -> host): bench.py's e2e figures.
 * Class-double descriptors whose values all survive (double)(float)x == x -- the reference's descriptors
 * are floats stored in a double matrix, M/sift/siftdescriptor.c:500-527 -- cross as float and are widened
 * on the device (same arithmetic); PRE3_HOST_F32=0 in the environment turns that off. */
PRE3_API int pre3_transfer_bytes(const pre3_ctx *ctx, int64_t *h2d, int64_t *d2h);
/* Evaluation schedule of the adaptive-stop paths, for a BATCH of pairs (>= 64 per call: one block per pair walks the
 * sample sets in chunks of 64 and replays the reference's loop control after every chunk; smaller batches run in waves
 * over the whole GPU).  Writes the boundaries (sample sets evaluated after step i) into ends[0..cap) and returns their
 * number: a pair that consumed n sets had ends[min{i : n < ends[i]}] (or H) sets evaluated.
 * bench.py derives the executed hypothesis x match evaluations from this. */
PRE3_API int pre3_eval_schedule(const pre3_ransac_opts *opts, int32_t *ends, int cap);
/* The same for a call with exactly P pairs: fewer than 64 pairs run in waves; batches that leave block slots idle
 * (P <= 592 / P <= 296 with k = 5, find_transform_matrix) use chunks of 128 / 256 sample sets -- a shorter critical
 * path per pair at the price of evaluating sets the reference loop would not reach. */
PRE3_API int pre3_eval_schedule_for(const pre3_ransac_opts *opts, int P, int32_t *ends, int cap);
/* Per-kernel CUDA-event timing on the context's stream (off by default; bench.py's roofline
 * numbers).  pre3_timing_read synchronises, adds the elapsed ms and launch counts of every
 * bracketed launch since the last read into ms[cat] / count[cat] (PRE3_TIMING_NCAT entries
 * each, caller-zeroed) and forgets them.  pre3_timing_name(cat) names a category. */
#define PRE3_TIMING_NCAT 15
PRE3_API int pre3_timing_enable(pre3_ctx *ctx, int on);
PRE3_API int pre3_timing_read(pre3_ctx *ctx, double *ms, int64_t *count);
PRE3_API const char *pre3_timing_name(int cat);
/* FFMA-chain microbenchmark: the measured FP32 CUDA-core peak (TFLOP/s) that the scoring
 * kernel's roofline fraction is quoted against. */
PRE3_API int pre3_measure_fp32_peak(pre3_ctx *ctx, double *tflops);
/* mode 0: FFMA with constant-bank operands (= pre3_measure_fp32_peak); mode 1: all three sources in registers,
 * the form the scorer's inner loop needs (per-thread hypothesis x per-match operand). */
PRE3_API int pre3_measure_fp32_peak_mode(pre3_ctx *ctx, int mode, double *tflops);

/* ---- stage 1: siftmatch -----------------------------------------------------------
 * Replaces compare_mx*_CLASS + mexFunction of M/sift/siftmatch.c:83-250.
 * L1: ND x K1, L2: ND x K2, same class.  pairs: 2 x K1 capacity (k1,k2 0-based, in k1
 * order), score: K1 capacity (best squared distance, siftmatch.c:243-245); *n_out =
 * number of accepted pairs.  thresh is narrowed to float like the reference (:87,:205). */
PRE3_API int pre3_siftmatch(pre3_ctx *ctx, const void *L1, const void *L2, int cls, int K1, int K2,
                   int ND, double thresh, int32_t *pairs, double *score, int32_t *n_out);

/* Callers of siftmatch.
 * pre3_siftmatch_sweep: ONE descriptor set L1 (ND x K1) against P sets L2 (each ND x K2, k2_count valid) -- the loop of
 * M/find_consistent_sift_matches.m:39-65, `matches = siftmatch(descriptor1, descriptor2)` with the first frame's
 * descriptors fixed.  Outputs as pre3_siftmatch_batch (P x K1 x 2 pairs, P x K1 scores, P counts); L1 is converted for
 * the tensor cores once.
 * pre3_matching_sift_based_batch: M/matching_sift_based.m:104-135 for P frames -- des1 = descriptors of the predicted
 * map features (ND x F per frame, f_count valid), des2 = Descriptor_RAW (ND x K2), siftmatch with `thresh` (the reference
 * uses the default 1.5), then the search-region gate :119-135: match i = (k1, k2) becomes individually compatible iff
 * norm(pos2(:,k2) - h(:,k1)) <= ceil(3*sqrt(S11(i))) (40 when S11(i) is NaN = empty S; the reference indexes S with the
 * loop counter i over the matches, reproduced).  h: P x F x 2, S11: P x F, pos2: P x K2 x 2 (column-major 2 x n each).
 * Outputs per predicted feature: ic (P x F, 0/1), z (P x F x 2: pos2 of the match, NaN where ic = 0), match (P x F:
 * 0-based k2 or -1); per frame: n_match (matches before the gate), n_discarded (StatData.DISCARDED_SIFT_MATCH). */
PRE3_API int pre3_siftmatch_sweep(pre3_ctx *ctx, const void *L1, const void *L2, int cls, int P, int K1, int K2, int ND,
                                  const int32_t *k2_count, double thresh, int32_t *pairs, double *score,
                                  int32_t *n_out);
PRE3_API int pre3_siftmatch_sweep_dev(pre3_ctx *ctx, const void *dL1, const void *dL2, int cls, int P, int K1, int K2,
                                      int ND, const int32_t *dk2_count, double thresh, int32_t *dpairs, double *dscore,
                                      int32_t *dn_out);
PRE3_API int pre3_matching_sift_based_batch(pre3_ctx *ctx, const void *des1, const void *des2, int cls, int P, int F,
                                            int K2, int ND, const int32_t *f_count, const int32_t *k2_count,
                                            const double *h, const double *S11, const double *pos2, double thresh,
                                            uint8_t *ic, double *z, int32_t *match, int32_t *n_match,
                                            int32_t *n_discarded);
/* P independent (L1_p, L2_p) problems, each padded to K1 x K2 with the valid counts in
 * k1_count / k2_count (NULL = all K1 / K2 valid).  pairs: P x (2 x K1), score: P x K1,
 * n_out: P.  (find_consistent_sift_matches.m:39-65 and SIFT_match_save.m:33 in bulk.) */
PRE3_API int pre3_siftmatch_batch(pre3_ctx *ctx, const void *L1, const void *L2, int cls, int P,
                         int K1, int K2, int ND, const int32_t *k1_count,
                         const int32_t *k2_count, double thresh, int32_t *pairs,
                         double *score, int32_t *n_out);
PRE3_API int pre3_siftmatch_batch_dev(pre3_ctx *ctx, const void *dL1, const void *dL2, int cls, int P,
                             int K1, int K2, int ND, const int32_t *dk1_count,
                             const int32_t *dk2_count, double thresh, int32_t *dpairs,
                             double *dscore, int32_t *dn_out);

/* ---- stage 2: minimal-sample / least-squares rigid fits ----------------------------
 * [rot,trans,state] = find_transform_matrix(pset1,pset2)
 * (M/mex_files/RANSAC_CALCULATION/find_transform_matrix.m:2-42).  pset: 3 x n. */
PRE3_API int pre3_find_transform_matrix(pre3_ctx *ctx, const double *pset1, const double *pset2, int n,
                               double *rot, double *trans, int32_t *state);
/* [s,R,T,err] = absoluteOrientationQuaternion(A,B,doScale)
 * (M/absoluteOrientationQuaternion.m:28-127).  n < 4 -> PRE3_ERR_ARG like :51-54. */
PRE3_API int pre3_horn(pre3_ctx *ctx, const double *A, const double *B, int n, int do_scale, double *s,
              double *R, double *T, double *err);
/* H fits from sample sets: samples k x H (0-based).  Outputs 9 x H (column-major each),
 * 3 x H, H.  method selects find_transform_matrix (Ya ~ R*Yb+T) or Horn(Yb,Ya). */
PRE3_API int pre3_fit_batch(pre3_ctx *ctx, const double *Ya, const double *Yb, int N,
                   const int32_t *samples, int k, int H, int method, double *R, double *T,
                   int32_t *state);

/* ---- stage 3: hypothesis support ---------------------------------------------------
 * For each of H hypotheses (R 9 x H column-major, T 3 x H): cardinality of
 * { i : ||R*Yb_i + T - Ya_i|| < thr } (RANSAC_CALC_VER2.m:121-125) and optionally the
 * ErrorSum (:135) and the inlier masks (N x H bytes).  fp32 scoring with fp64 recheck
 * of threshold-borderline residuals: results are those of the fp64 reference. */
PRE3_API int pre3_score_batch(pre3_ctx *ctx, const double *R, const double *T, int H, const double *Ya,
                     const double *Yb, int N, double thr, int32_t *count, double *errsum,
                     uint8_t *mask);

/* ---- stages 2-4: the RANSAC loop ---------------------------------------------------
 * [R,T,error,BestFit,State_RANSAC] = RANSAC_CALC_VER2(Ya,Yb,options,...)
 * (M/mex_files/RANSAC_CALCULATION/RANSAC_CALC_VER2.m:2-201; Horn flavour
 * M/RANSAC_CALC_VER_test.m).  Ya,Yb: 3 x N.  samples: k x opts->H explicit sets
 * (0-based) or NULL for the seeded generator.  mask: N bytes (PositionInliers of the
 * winner) or NULL.  counts/states: opts->H entries each or NULL (per sample set:
 * cardinality or -1 if not evaluated; find_transform_matrix state). */
PRE3_API int pre3_ransac(pre3_ctx *ctx, const double *Ya, const double *Yb, int N,
                const pre3_ransac_opts *opts, const int32_t *samples, pre3_pair_result *res,
                uint8_t *mask, int32_t *counts, int8_t *states);

/* RANSAC_CALC_VER2 (M/mex_files/RANSAC_CALCULATION/RANSAC_CALC_VER2.m:2-201) for P independent correspondence sets at
 * once -- the pairs M/find_consistent_sift_matches.m:22-32 loops over --, padded to Nmax columns each; n_corr[P] valid
 * counts.  samples: P x (k x H) or NULL.  masks: P x Nmax or NULL. */
PRE3_API int pre3_ransac_batch(pre3_ctx *ctx, const double *Ya, const double *Yb, const int32_t *n_corr,
                      int P, int Nmax, const pre3_ransac_opts *opts, const int32_t *samples,
                      pre3_pair_result *res, uint8_t *masks);
PRE3_API int pre3_ransac_batch_dev(pre3_ctx *ctx, const double *dYa, const double *dYb,
                          const int32_t *dn_corr, int P, int Nmax, const pre3_ransac_opts *opts,
                          const int32_t *dsamples, pre3_pair_result *dres, uint8_t *dmasks);

/* ---- covariance of the RANSAC pose -------------------------------------
 * cov_est_RANSAC_deriv(M(BestFitIdx).SupportSet.Ya, M(BestFitIdx).SupportSet.Yb, R, T) for P pairs.  Ya, Yb: P x Nmax x 3
 * correspondences (n_corr valid per pair, NULL: Nmax); masks: P x Nmax inlier flags of the winner = the SupportSet
 * (NULL: every correspondence); R: P x 9 column-major, T: P x 3 with Ya ~ R*Yb + T.  The _dev form reads R and T as 12
 * consecutive doubles per pair, rt_stride doubles apart: 12 for a packed array, sizeof(pre3_pair_result)/8 with dRT =
 * &res[0].R to run on the records pre3_pairs_dev / pre3_ransac_batch_dev left on the device.  Nested central
 * differences with eps 1e-6 (M/deriv.m:7-9): agreement with the reference is ~1e-6 relative, not bit for bit. */
PRE3_API int pre3_cov_est_ransac_batch(pre3_ctx *ctx, const double *Ya, const double *Yb, const int32_t *n_corr,
                                       const uint8_t *masks, int P, int Nmax, const double *R, const double *T,
                                       pre3_cov_result *out);
PRE3_API int pre3_cov_est_ransac_batch_dev(pre3_ctx *ctx, const double *dYa, const double *dYb,
                                           const int32_t *dn_corr, const uint8_t *dmasks, int P, int Nmax,
                                           const double *dRT, int rt_stride, pre3_cov_result *dout);

/* ---- EKF partial updates around ransac_hypotheses (first part) -----
 * [x_k_k, p_k_k] = update(x, p, H, R, z, h) (M/update.m:27-56) for Fr frames, with z, h, H stacked from the features
 * whose flag sel is 1, in feature order, and R = r_diag * eye -- what ekf_update_li_inliers.m:15-29 (sel =
 * low_innovation_inlier, x / P = x_k_km1 / p_k_km1) and ekf_update_hi_inliers.m:18-32 (sel = high_innovation_inlier,
 * x / P = x_k_k / p_k_k) do with r_diag = 1.  Feature arrays as in pre3_ransac_hypotheses_batch (type, pos, z, h,
 * Hcam 2 x 13, Hfeat 2 x 6 per feature).  Outputs: x_out Fr x n, P_out Fr x (n x n) (must not alias P), m_out Fr
 * (rows of the stacked system; 0: the frame is copied through unchanged, update.m:50-54) or NULL.
 * Dense fp64 linear algebra: results agree with the reference arithmetic to rounding (tests: 1e-9 of max|P|). */
PRE3_API int pre3_ekf_update_batch(pre3_ctx *ctx, int Fr, int n, int F, const double *x, const double *P,
                          const int32_t *type, const int32_t *pos, const uint8_t *sel, const double *z,
                          const double *h, const double *Hcam, const double *Hfeat, double r_diag, double *x_out,
                          double *P_out, int32_t *m_out);
PRE3_API int pre3_ekf_update_batch_dev(pre3_ctx *ctx, int Fr, int n, int F, const double *dx, const double *dP,
                              const int32_t *dtype, const int32_t *dpos, const uint8_t *dsel, const double *dz,
                              const double *dh, const double *dHcam, const double *dHfeat, double r_diag,
                              double *dx_out, double *dP_out, int32_t *dm_out);
/* [x_k_k, p_k_k, K] = update(x_km1_k, p_km1_k, H, R, z, h) with the function's own signature (M/update.m:27): H m x n
 * and R m x m dense column-major (a sparse MATLAB H is expanded by the gateway), n >= 7.  K_out: n x m or NULL.
 * m == 0: x and P are copied through (:50-54; the reference returns K = 0). */
PRE3_API int pre3_ekf_update_dense(pre3_ctx *ctx, int n, int m, const double *x, const double *P, const double *H,
                          const double *R, const double *z, const double *h, double *x_out, double *P_out,
                          double *K_out);
PRE3_API int pre3_ekf_update_dense_dev(pre3_ctx *ctx, int n, int m, const double *dx, const double *dP,
                              const double *dH, const double *dR, const double *dz, const double *dh,
                              double *dx_out, double *dP_out, double *dK_out);
/* The test of rescue_hi_inliers.m:35-46 on device buffers: for features with ic == 1 and li == 0,
 * hi = (nu' * inv(H p_k_k H') * nu < 5.9915), nu = z - h; other entries of hi are left untouched.  h, Hcam, Hfeat
 * are the measurements re-predicted at x_k_k (:32-33: predict_camera_measurements / calculate_derivatives stay with
 * the caller). */
PRE3_API int pre3_ekf_rescue_hi_inliers_batch_dev(pre3_ctx *ctx, int Fr, int n, int F, const double *dP_kk,
                                         const int32_t *dtype, const int32_t *dpos, const uint8_t *dic,
                                         const uint8_t *dli, const double *dz, const double *dh,
                                         const double *dHcam, const double *dHfeat, uint8_t *dhi);

/* ---- SR4000 frame batches -> per-feature 3-D points ----------------
 * The step before the matching path: what SIFT_extract_save.m:75-88 does per frame through
 * read_xyz_sr4000.m:8-21 (3 x 3 Gaussian smoothing of the x, y, z maps) and
 * inittialize_depth_my_version.m:16,40-85 (x(round(v),round(u)) lookup, rejection of NaN / closer than 0.4 m /
 * confidence <= max/2, [-x,-y,z]); and its code_from_dr_ye flavour (read_sr4000_data_dr_ye.m:8,88-90: sigma 1,
 * 'replicate'; confidence_filtering.m:1-13; the lookup of ransac_dr_ye.m:13-19).
 * sr_data: F frames, each the rows x 176 column-major double matrix `load('d1_%04d.dat')` returns: rows 1:144 z,
 * 145:288 x, 289:432 y, 433:576 amplitude, 577:720 confidence (rows >= 720), row 721 time stamp. */
typedef struct {
  double sigma;           /* fspecial('gaussian',[3 3],sigma): 2 (read_xyz_sr4000.m:8), 1 (read_sr4000_data_dr_ye.m:8) */
  int32_t boundary;       /* 0 zero padding (imfilter(...,'same'), read_xyz_sr4000.m:15-21); 1 'replicate' (dr_ye :88-90) */
  int32_t mode;           /* 0 inittialize_depth_my_version.m rejection rules; 1 dr_ye: confidence_filtering.m only */
  int32_t rows;           /* rows of sr_data per frame: 576, 720 or 721 */
  int32_t use_confidence; /* mode 1: myCONFIG.FLAGS.CONFIDENCE_MAP (vodometry_dr_ye.m:80-82) */
} pre3_frame_opts;

/* [x,y,z,confidence_map] = read_xyz_sr4000(prefix, k) for F frames already loaded: the filtered maps, F x (144 x 176)
 * column-major each; max_conf: F values of max(confidence_map(:)) (NaN when rows < 720) or NULL. */
PRE3_API int pre3_read_xyz_sr4000_batch(pre3_ctx *ctx, const double *sr_data, int F, const pre3_frame_opts *opts,
                               double *x, double *y, double *z, double *max_conf);
PRE3_API int pre3_read_xyz_sr4000_batch_dev(pre3_ctx *ctx, const double *dsr_data, int F,
                                   const pre3_frame_opts *opts, double *dx, double *dy, double *dz,
                                   double *dmax_conf);
/* The per-feature loop of SIFT_extract_save.m:75-88, fused with the smoothing (the stencil is evaluated only at
 * the pixels the features round to).  frames: F x (frame_ld x K), rows 1:2 = the 0-based (x, y) sift returns (:55-56
 * add 1); k_count: valid features per frame or NULL.  Outputs (any may be NULL): xyz_all F x (3 x K) with NaN columns
 * for rejected features (xyz_data of :80-83); keep F x K; n_keep F; idx_remain F x K (0-based idxRemain, -1 beyond
 * n_keep); xyz F x (3 x K) compacted XYZ_DATA (:88); desc_out = Descriptor(:, idxRemain) (:86, class cls, ND rows,
 * zero beyond n_keep); frames_out = SCALE_ORIENT_POS(:, idxRemain) (:87); n_oob: features whose rounded position lies
 * outside the 144 x 176 image (the reference raises an index error; they are rejected here). */
PRE3_API int pre3_features_xyz_batch(pre3_ctx *ctx, const double *sr_data, int F, const pre3_frame_opts *opts,
                            const double *frames, int frame_ld, int K, const int32_t *k_count, double *xyz_all,
                            uint8_t *keep, int32_t *n_keep, int32_t *idx_remain, double *xyz,
                            const void *desc_in, int cls, int ND, void *desc_out, double *frames_out,
                            int32_t *n_oob);
PRE3_API int pre3_features_xyz_batch_dev(pre3_ctx *ctx, const double *dsr_data, int F, const pre3_frame_opts *opts,
                                const double *dframes, int frame_ld, int K, const int32_t *dk_count,
                                double *dxyz_all, uint8_t *dkeep, int32_t *dn_keep, int32_t *didx_remain,
                                double *dxyz, const void *ddesc_in, int cls, int ND, void *ddesc_out,
                                const double *dframes_in, double *dframes_out, int32_t *dn_oob);

/* ---- the code_from_dr_ye variant ---------------------------------
 * The RANSAC part of [rot,phi,theta,psi,trans,error,pnum,op_num,sta,op_pset1,op_pset2,RANSAC_STAT] =
 * vodometry_dr_ye(file1,file2) (M/code_from_dr_ye/vodometry_dr_ye.m:147-220, loop body ransac_dr_ye.m:20-71),
 * the variant the live EKF calls (M/fv.m:47 -> Calculate_V_Omega_RANSAC_dr_ye.m:19-22), for P match sets.
 * Ya = pset1 (frame 1), Yb = pset2 (frame 2): pset1 ~ rot*pset2 + trans; opts->method is taken as
 * PRE3_METHOD_DR_YE, opts->k must be 4, opts->max_iteration = 700 reproduces :162.
 * match: P x (2 x Nmax) feature ids [k1;k2] of every match (what siftmatch returned; only compared for
 * equality by the sampler, ransac_dr_ye.m:33-48) or NULL (match i = [i;i]).
 * samples: P x (4 x H) explicit draws num_rs(1..4) (0-based, draw order) or NULL: the reference's sampler
 * run on a seeded uniform stream (opts->seed, pair id).
 * pre3_pair_result fields in this mode: status 0 ok / 1 pnum < 4 (:152) / 4 no consensus, op_num < 3 (:187) /
 * 5 no point farther than 0.4 m (ransac_dr_ye.m:21-22 raises an error); state = sta of the refit (:211);
 * best_fit = op_num; best_sample = rs_ind - 1; n_consumed = iterations executed; n_iter =
 * RANSAC_STAT.nIterationRansac (:216); thr = 0.001*dist (bound on the SQUARED distance, ransac_dr_ye.m:70);
 * error_sum = sum of the residual norms over the support set after the refit (:212-213).
 * masks: P x Nmax (support set of the winner) or NULL; stat: P or NULL; counts: P x H (tmp_cnum, -1 beyond
 * the executed iterations) or NULL. */
typedef struct {
  double error_mean;          /* RANSAC_STAT.ErrorMean (:214) */
  double error_std;           /* RANSAC_STAT.ErrorStd  (:215, normalised by n - 1) */
  double dist;                /* ransac_dr_ye.m:23 */
  int32_t n_iteration_ransac; /* RANSAC_STAT.nIterationRansac (:216) */
  int32_t n_loops;            /* iterations executed: min(700, nchoosek(pnum,4), H) */
} pre3_dr_ye_stat;

PRE3_API int pre3_vodometry_dr_ye_batch(pre3_ctx *ctx, const double *Ya, const double *Yb, const int32_t *n_corr,
                               const int32_t *match, int P, int Nmax, const pre3_ransac_opts *opts,
                               const int32_t *samples, pre3_pair_result *res, uint8_t *masks,
                               pre3_dr_ye_stat *stat, int32_t *counts);
PRE3_API int pre3_vodometry_dr_ye_batch_dev(pre3_ctx *ctx, const double *dYa, const double *dYb,
                                   const int32_t *dn_corr, const int32_t *dmatch, int P, int Nmax,
                                   const pre3_ransac_opts *opts, const int32_t *dsamples, uint32_t pair_id0,
                                   pre3_pair_result *dres, uint8_t *dmasks, pre3_dr_ye_stat *dstat,
                                   int32_t *dcounts);

/* ---- whole frame pairs: match -> gather -> RANSAC ------------------------------------
 * (opts->method = PRE3_METHOD_DR_YE runs the variant above on the matches of every pair: match ids =
 * the siftmatch output, seeded sampler.)
 * What SIFT_match_save.m:33-53 does per pair (and RANSAC_CALC_SAVE_SR4000.m /
 * Calculate_V_Omega_RANSAC_my_version.m around it), for P pairs in one call:
 *   matches = siftmatch(desc1_p, desc2_p); Ya = xyz1(:,matches(1,:)); Yb = xyz2(:,matches(2,:));
 *   RANSAC_CALC_VER2(Ya, Yb, options).
 * desc: 128 x K per pair (class cls), xyz: 3 x K per pair (double).  matches: P x (2 x K1)
 * (0-based) or NULL; masks: P x K1 or NULL.  Sample sets are seeded (opts->seed, pair id
 * = pair_id0 + p) because N is only known after matching. */
PRE3_API int pre3_pairs(pre3_ctx *ctx, const void *desc1, const void *desc2, int cls, const double *xyz1,
               const double *xyz2, int P, int K1, int K2, int ND, const int32_t *k1_count,
               const int32_t *k2_count, const pre3_ransac_opts *opts, uint32_t pair_id0,
               pre3_pair_result *res, int32_t *matches, uint8_t *masks);
PRE3_API int pre3_pairs_dev(pre3_ctx *ctx, const void *ddesc1, const void *ddesc2, int cls,
                   const double *dxyz1, const double *dxyz2, int P, int K1, int K2, int ND,
                   const int32_t *dk1_count, const int32_t *dk2_count,
                   const pre3_ransac_opts *opts, uint32_t pair_id0, pre3_pair_result *dres,
                   int32_t *dmatches, uint8_t *dmasks);

/* A SEQUENCE of F frames = F - 1 consecutive pairs (frame p, frame p + 1), the way the reference's
 * whole-sequence loops call the pair driver (M/find_consistent_sift_matches.m:22-32:
 * RANSAC_CALC_SAVE_SR4000(i, i + 1) for every i; M/Test_RANSAC_dead_reckoning.m).  Same results as pre3_pairs
 * on (desc[p], desc[p+1]), but every frame's descriptors cross PCIe and are converted for the tensor cores once
 * instead of twice.  desc: F x (128 x K), xyz: F x (3 x K), k_count: F or NULL; res / matches / masks: F - 1
 * entries (pair p uses seeded sample sets of pair id pair_id0 + p). */
PRE3_API int pre3_sequence(pre3_ctx *ctx, const void *desc, int cls, const double *xyz, int F, int K, int ND,
                  const int32_t *k_count, const pre3_ransac_opts *opts, uint32_t pair_id0,
                  pre3_pair_result *res, int32_t *matches, uint8_t *masks);
PRE3_API int pre3_sequence_dev(pre3_ctx *ctx, const void *ddesc, int cls, const double *dxyz, int F, int K, int ND,
                      const int32_t *dk_count, const pre3_ransac_opts *opts, uint32_t pair_id0,
                      pre3_pair_result *dres, int32_t *dmatches, uint8_t *dmasks);

/* ---- hypothesis-block sharding (one large pair over several GPUs) -------
 * Evaluates sample sets [h0, h0+Hloc) of the pair on this device and returns the local
 * best as key = (count << 32) | (0xFFFFFFFF - global_hypothesis_id) ("first-index" mode:
 * max count, lowest id) plus the local (count, id, errsum) triple for the
 * reference-exact (max count, min ErrorSum, first id) all-gather mode.  The caller does
 * the one NCCL max-reduce / all-gather, then pre3_ransac_finish on the winner's id. */
PRE3_API int pre3_ransac_block_dev(pre3_ctx *ctx, const double *dYa, const double *dYb, int N,
                          const pre3_ransac_opts *opts, const int32_t *dsamples, int64_t h0,
                          int Hloc, double thr, uint64_t *dkey, double *derrsum);
/* Reference-exact mode of the split: the block's own winner under the full rule of
 * RANSAC_CALC_VER2.m:165-175 (max cardinality, then min ErrorSum, then first index) with its mask and
 * refit; res->best_sample is LOCAL (add h0).  The caller all-gathers (best_fit, h0 + best_sample,
 * error_sum) -- 16 bytes per rank -- and every rank picks the same global winner; the owning rank
 * already holds R, T and the mask. */
PRE3_API int pre3_ransac_block_select_dev(pre3_ctx *ctx, const double *dYa, const double *dYb, int N,
                                 const pre3_ransac_opts *opts, const int32_t *dsamples, int64_t h0,
                                 int Hloc, double thr, pre3_pair_result *dres, uint8_t *dmask);
PRE3_API int pre3_ransac_finish_dev(pre3_ctx *ctx, const double *dYa, const double *dYb, int N,
                           const pre3_ransac_opts *opts, const int32_t *dsamples_of_winner,
                           int64_t winner_id, double thr, pre3_pair_result *dres, uint8_t *dmask);
/* thr := 0.01*||Yb(:,argmin z)|| (RANSAC_CALC_VER2.m:69-72), computed on the device. */
PRE3_API int pre3_distance_threshold_dev(pre3_ctx *ctx, const double *dYb, int N, double *dthr);
/* The same split, STREAM-ORDERED: nothing is read back between the kernels and the collective (the threshold, the
 * keys and the winner stay in device memory), so one solve is  split_local -> ONE collective -> split_finish -> a SUM
 * all-reduce of the 240-byte record (every rank but the owner holds zeros).
 *   mode 0 ("first": max inlier count, lowest hypothesis id -- what BASELINE.json config 5 specifies):
 *     split_local writes dkey[0] = (count << 32) | (0xFFFFFFFF - global id); the caller all-reduces it with MAX;
 *     split_finish(dexchanged = the reduced key) lets the rank whose block holds the winner compute hypothesis, mask,
 *     ErrorSum and refit (RANSAC_CALC_VER2.m:186); the other ranks write zeros.
 *   mode 1 (the reference's full rule :165-175): split_local runs the block's own selection (record in dres, mask in
 *     dmask) and writes dkey[0..1] = (key, ErrorSum bits); the caller all-gathers the 16 bytes of every rank;
 *     split_finish(dexchanged = world x 2 words) picks max count, min ErrorSum, lowest id: the owner keeps its record
 *     (best_sample made global), the others zero theirs. */
PRE3_API int pre3_ransac_split_local_dev(pre3_ctx *ctx, const double *dYa, const double *dYb, int N,
                                         const pre3_ransac_opts *opts, const int32_t *dsamples, int64_t h0, int Hloc,
                                         int mode, uint64_t *dkey, pre3_pair_result *dres, uint8_t *dmask);
PRE3_API int pre3_ransac_split_finish_dev(pre3_ctx *ctx, const double *dYa, const double *dYb, int N,
                                          const pre3_ransac_opts *opts, const int32_t *dsamples, int64_t h0, int Hloc,
                                          int mode, const uint64_t *dexchanged, int world, int rank,
                                          pre3_pair_result *dres, uint8_t *dmask);

/* ---- config 4: 1-point-RANSAC EKF hypotheses --------------------------------------------
 * ransac_hypotheses (M/ransac_hypotheses.m:27-85) and compute_hypothesis_support_fast
 * (M/compute_hypothesis_support_fast.m:27-116).  fp64 throughout; sin/cos and inv(S) follow the
 * fixed-order algorithms specified in oracle/pre3_oracle_ekf.c. */
typedef struct {
  double f, Cx, Cy, k1, k2; /* cam.f, cam.Cx, cam.Cy, cam.k1, cam.k2 (M/initialize_cam.m:64-76) */
} pre3_cam;

typedef struct {
  int32_t n_hyp_init; /* 1000 (ransac_hypotheses.m:35) */
  int32_t H;          /* match selections available per frame (the loop also ends when they run out) */
  int32_t adaptive;   /* 1: the reference's stop rule (:77-80); 0: evaluate all H, first maximum wins */
  int32_t reserved;
  uint64_t seed;      /* seeded selections (used when none are supplied) */
} pre3_ekf_opts;

typedef struct {
  int32_t status;      /* 0 ok; 1 no individually compatible match (select_random_match.m:50 errors);
                          3 an individually compatible feature has no measurement z */
  int32_t n_evaluated; /* hypotheses the reference's loop evaluates */
  int32_t best_hyp;    /* 0-based index of the most supported hypothesis, -1 if none had support */
  int32_t max_support; /* StatData.RANSAC_HYP_SUPPORT (:85) */
  int32_t num_ic;      /* individually compatible matches */
  int32_t m;           /* matches per hypothesis: 3 if num_ic > 3 else 1 (select_random_match.m:47-51) */
  double n_hyp;        /* StatData.RANSAC_ITER (:84): the final n_hyp */
} pre3_ekf_result;

/* [support, li_id, li_euc] = compute_hypothesis_support_fast(xi, cam, pattern, z_id, z_euc, thr) for B
 * hypothesised states at once (B = 1 is the reference call).  xi: n x B; pattern: n x 4 (the 0/1
 * matrix of generate_state_vector_pattern.m:29-51); z_id: 2 x n_id; z_euc: 2 x n_euc;
 * support: B; li_id: n_id x B bytes; li_euc: n_euc x B bytes (either may be NULL). */
PRE3_API int pre3_ekf_support(pre3_ctx *ctx, const double *xi, int n, int B, const pre3_cam *cam,
                     const double *pattern, const double *z_id, int n_id, const double *z_euc,
                     int n_euc, double threshold, int32_t *support, uint8_t *li_id, uint8_t *li_euc);

/* features_info = ransac_hypotheses(filter, features_info, cam) for Fr frames that share the state
 * size n and the feature count F.  Per frame (all column-major, frames consecutive):
 *   x n (get_x_k_km1), P n x n (get_p_k_km1); std_z = get_std_z(filter) (the threshold, :33);
 *   per feature i of features_info: type (0 'inversedepth', 1 'cartesian'), pos (0-based offset of
 *   its states in x), has_z (~isempty(z)), ic (individually_compatible), z 2, h 2,
 *   Hcam 2 x 13 = H(:,1:13), Hfeat 2 x 6 = H(:,pos+1:pos+6) (cartesian: first 3 columns; H has no
 *   other non-zeros, M/calculate_Hi_inverse_depth_my_version.m:44-49), R 2 x 2.
 *   sel: H x 3 supplied match selections per frame (0-based feature indices in the order
 *   select_random_match.m:58 returns them; only the first m are read) or NULL (seeded, frame id =
 *   frame_id0 + frame).
 * Outputs: li_inlier Fr x F bytes IN/OUT (low_innovation_inlier: written for features with z when a
 * hypothesis had support, untouched otherwise, set_as_most_supported_hypothesis.m:32-53); res Fr;
 * supports Fr x H (support of every hypothesis the loop evaluated, -1 beyond) or NULL. */
PRE3_API int pre3_ransac_hypotheses_batch(pre3_ctx *ctx, int Fr, int n, int F, const double *x, const double *P,
                                 double std_z, const pre3_cam *cam, const int32_t *type,
                                 const int32_t *pos, const uint8_t *has_z, const uint8_t *ic,
                                 const double *z, const double *h, const double *Hcam,
                                 const double *Hfeat, const double *R, const int32_t *sel,
                                 const pre3_ekf_opts *opts, uint32_t frame_id0, uint8_t *li_inlier,
                                 pre3_ekf_result *res, int32_t *supports);
PRE3_API int pre3_ransac_hypotheses_batch_dev(pre3_ctx *ctx, int Fr, int n, int F, const double *dx,
                                     const double *dP, double std_z, const pre3_cam *cam,
                                     const int32_t *dtype, const int32_t *dpos, const uint8_t *dhas_z,
                                     const uint8_t *dic, const double *dz, const double *dh,
                                     const double *dHcam, const double *dHfeat, const double *dR,
                                     const int32_t *dsel, const pre3_ekf_opts *opts,
                                     uint32_t frame_id0, uint8_t *dli_inlier, pre3_ekf_result *dres,
                                     int32_t *dsupports);
/* Wave boundaries of the hypothesis evaluation (like pre3_eval_schedule). */
PRE3_API int pre3_ekf_eval_schedule(const pre3_ekf_opts *opts, int32_t *ends, int cap);
/* tcgen05.ld sweep: measured TMEM -> register read rate (GB/s, whole chip), the bound of the matching GEMM's
 * epilogue (every fp32 accumulator is read once: 4 B of TMEM per 2*128 flops). */
PRE3_API int pre3_measure_tmem_read(pre3_ctx *ctx, double *gbs);
/* DFMA-chain microbenchmark: measured FP64 CUDA-core peak (TFLOP/s), the roofline of the EKF kernels. */
PRE3_API int pre3_measure_fp64_peak(pre3_ctx *ctx, double *tflops);

/* R2q of slamToolbox (M/slamToolbox_11_02_18/FrameTransforms/Rotations/R2q.m:11-55), host
 * helper used by the Calculate_V_Omega_RANSAC* shims: q = [a -b -c -d]'. */
PRE3_API void pre3_R2q(const double *R_colmajor, double *q);

/* Re-prediction at x_k_k, the first two lines of rescue_hi_inliers.m (:32-33): predict_camera_measurements.m:27-68
 * (hi_inverse_depth.m / hi_cartesian.m: +-60 degree field of view, pinhole, radial distortion, image bounds 0 < u < nCols,
 * 0 < v < nRows) and calculate_derivatives.m:27-59 (calculate_Hi_inverse_depth_my_version.m / calculate_Hi_cartesian_
 * my_version.m) for Fr frames of F features.  x: Fr x n; type / pos as above; has_h (Fr x F, NULL = all 1): the feature
 * carries a previous prediction h_in (Fr x F x 2).  Outputs: h_out (a feature that fails a visibility test keeps h_in,
 * predict_camera_measurements.m:37-39), has_h_out, predicted (passed the tests at x), Hcam (2 x 13 per feature; columns
 * 8..13 are zero) and Hfeat (2 x 6) for every feature with has_h_out, zeros otherwise -- the inputs
 * pre3_ekf_rescue_hi_inliers_batch_dev and the hi-inlier update take. */
PRE3_API int pre3_ekf_predict_measurements_batch(pre3_ctx *ctx, int Fr, int n, int F, const double *x,
                                                 const pre3_cam *cam, int nRows, int nCols, const int32_t *type,
                                                 const int32_t *pos, const uint8_t *has_h, const double *h_in,
                                                 double *h_out, uint8_t *has_h_out, uint8_t *predicted, double *Hcam,
                                                 double *Hfeat);
PRE3_API int pre3_ekf_predict_measurements_batch_dev(pre3_ctx *ctx, int Fr, int n, int F, const double *dx,
                                                     const pre3_cam *cam, int nRows, int nCols, const int32_t *dtype,
                                                     const int32_t *dpos, const uint8_t *dhas_h, const double *dh_in,
                                                     double *dh_out, uint8_t *dhas_h_out, uint8_t *dpredicted,
                                                     double *dHcam, double *dHfeat);

#ifdef __cplusplus
}
#endif
#endif /* PRE3_H */
