// capi.cu -- the C ABI of libpre3.so (include/pre3.h): argument checks, workspace
// carving, host<->device staging and kernel sequencing.  No arithmetic lives here.
//
// Every compute entry point fails with PRE3_ERR_CUDA when there is no CUDA device: there is
// no CPU fallback by design (the CPU restatement lives in oracle/ and is test-only).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "common.cuh"
#include "match.cuh"
#include "cov.cuh"
#include "ransac.cuh"

using namespace pre3;

namespace {

size_t class_size(int cls) {
  switch (cls) {
    case PRE3_CLASS_DOUBLE: return 8;
    case PRE3_CLASS_SINGLE: return 4;
    case PRE3_CLASS_INT8:
    case PRE3_CLASS_UINT8: return 1;
    default: return 0;
  }
}

#define PRE3_NEED(cond, msg) \
  do {                       \
    if (!(cond)) return fail(ctx, PRE3_ERR_ARG, msg); \
  } while (0)

int h2d(pre3_ctx* ctx, void* d, const void* h, size_t bytes) {
  if (bytes == 0) return PRE3_OK;
  PRE3_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return PRE3_OK;
}

int d2h(pre3_ctx* ctx, void* h, const void* d, size_t bytes) {
  if (bytes == 0) return PRE3_OK;
  PRE3_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return PRE3_OK;
}

int check_opts(pre3_ctx* ctx, const pre3_ransac_opts* o) {
  PRE3_NEED(o != nullptr, "options missing");
  PRE3_NEED(o->method == PRE3_METHOD_SVD || o->method == PRE3_METHOD_HORN || o->method == PRE3_METHOD_DR_YE,
            "unknown RANSAC method");
  PRE3_NEED(o->k >= 3 && o->k <= 8, "minimal sample size k must be in 3..8");
  PRE3_NEED(o->method != PRE3_METHOD_DR_YE || o->k == 4, "the dr_ye variant draws 4 matches per hypothesis");
  PRE3_NEED(o->H >= 0, "H must be >= 0");
  PRE3_NEED(o->max_iteration >= 0, "MaxIteration must be >= 0");
  return PRE3_OK;
}

// ---- stage 1 on device buffers; workspace must already hold match_ws_bytes() ---------------
size_t match_ws_bytes(int cls, int P, int K1, int K2, int ND) {
  return match_workspace_bytes(cls, P, K1, K2, ND) + align_up(sizeof(MatchRow) * (size_t)P * K1) + 1024;
}

int match_impl(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int P, int K1, int K2, int ND,
               const int32_t* dk1, const int32_t* dk2, double thresh, int need_score, MatchRow** rows_out) {
  MatchRow* rows = ws_take<MatchRow>(ctx, (size_t)P * K1);
  const float th = (float)thresh;  // narrowed like the reference (siftmatch.c:87,:205)
  bool tc = false, ti = false;
  if (ctx->match_engine != PRE3_MATCH_EXACT) {
    // the converters of both tensor-core engines (and the row recheck) read descriptors with 16-byte vector loads
    const bool aligned = (((uintptr_t)dL1 | (uintptr_t)dL2) & 15u) == 0;
    tc = aligned && match_tc_supported(cls, K1, K2, ND);
    ti = aligned && match_i8_supported(cls, K1, K2, ND);  // integer classes: exact kind::i8 GEMM
  }
  if (ctx->match_engine == PRE3_MATCH_TC && !tc && !ti)
    return fail(ctx, PRE3_ERR_ARG, "tensor-core matcher needs ND == 128 and 16-byte aligned descriptor sets");
  if (tc) {
    PRE3_TRY(launch_match_tc(ctx, dL1, dL2, cls, P, K1, K2, ND, dk1, dk2, th, need_score, rows));
  } else if (ti) {
    PRE3_TRY(launch_match_i8(ctx, dL1, dL2, cls, P, K1, K2, ND, dk1, dk2, th, rows));
  } else {
    if (!dL2) {  // sequence mode: pair p = (set p, set p + 1) of dL1
      const size_t esz = cls == PRE3_CLASS_DOUBLE_F32 ? 4 : class_size(cls);
      dL2 = (const char*)dL1 + (size_t)K1 * ND * esz;
      if (dk1) dk2 = dk1 + 1;
    }
    PRE3_TRY(launch_match_exact(ctx, dL1, dL2, cls, P, K1, K2, ND, dk1, dk2, th, rows));
  }
  *rows_out = rows;
  return PRE3_OK;
}

int check_match_args(pre3_ctx* ctx, const void* L1, const void* L2, int cls, int P, int K1, int K2, int ND) {
  PRE3_NEED(ctx != nullptr, "context missing");
  if (class_size(cls) == 0) return fail(ctx, PRE3_ERR_CLASS, "Unsupported numeric class");
  PRE3_NEED(P >= 0 && K1 >= 0 && K2 >= 0 && ND >= 0, "negative size");
  PRE3_NEED((L1 != nullptr || (size_t)P * K1 * ND == 0) && (L2 != nullptr || (size_t)P * K2 * ND == 0),
            "descriptor pointer missing");
  return PRE3_OK;
}

// ---- stages 2-4 on device buffers; workspace must already hold ransac_ws_bytes() -----------
size_t ransac_ws_bytes(int P, int Nmax, int H) { return ransac_workspace_bytes(P, Nmax, H) + 4096; }

int ransac_impl(pre3_ctx* ctx, const double* dYa, const double* dYb, const int32_t* dn_corr, int P, int Nmax,
                const pre3_ransac_opts& o, const int32_t* dsamples, uint32_t pair_id0, pre3_pair_result* dres,
                uint8_t* dmasks, int32_t* dcounts, int8_t* dstates, const int32_t* dmatch = nullptr,
                pre3_dr_ye_stat* dstat = nullptr) {
  RansacBuffers b{};
  b.Ya = dYa;
  b.Yb = dYb;
  b.n_corr = dn_corr;
  b.P = P;
  b.Nmax = Nmax;
  b.samples = dsamples;
  b.pair_id0 = pair_id0;
  ransac_carve(ctx, b, o.H);
  if (o.method == PRE3_METHOD_DR_YE) {
    if (dstates) return fail(ctx, PRE3_ERR_ARG, "per-hypothesis states are not reported by the dr_ye variant");
    PRE3_TRY(pipe_enter(ctx, PS_EVAL));
    return launch_dr_ye(ctx, b, o, dmatch, dres, dmasks, dstat, dcounts);
  }
  PRE3_TRY(ensure_adaptive_table(ctx, o, Nmax, dn_corr, P, &b.tab));
  PRE3_TRY(pipe_enter(ctx, PS_EVAL));
  PRE3_TRY(launch_prep(ctx, b, o, 0));
  // Selection inside the evaluation kernel (one launch for stages 2-4) is built and parity-green, but measured slower
  // at the sequence shape (eval + select 0.57 ms separate, 0.73 ms fused per 4096 pairs: 64 threads per pair leave the
  // fp64 tie sums and the refit too little parallelism); PRE3_FUSED_SELECT=1 switches it on.
  static const int fuse = getenv("PRE3_FUSED_SELECT") ? atoi(getenv("PRE3_FUSED_SELECT")) : 0;
  if (!dcounts && !dstates && fuse && ransac_can_fuse_select(b, o)) return launch_eval_select_fused(ctx, b, o, dres, dmasks);
  PRE3_TRY(launch_eval_waves(ctx, b, o));
  PRE3_TRY(pipe_enter(ctx, PS_SELECT));
  PRE3_TRY(launch_select(ctx, b, o, dres, dmasks, dcounts, dstates));
  return PRE3_OK;
}

}  // namespace

// ---- CUDA-graph replay (pre3_set_graphs) ------------------------------------------------------------------------
// A whole-pair call is ~15-20 short launches; at a few hundred pairs per call (a sequence sharded over 8 GPUs) the
// host's launch rate, not the GPU, bounds the step.  With graphs on, the FIRST call with a given signature runs eagerly
// (it also sizes the arena and the cached tables), the SECOND is captured from the stream into a graph, and every later
// call with the same signature is one cudaGraphLaunch.  Per-launch timing switches replay off.
namespace {
template <typename T>
void key_put(std::vector<unsigned char>& k, const T& v) {
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&v);
  k.insert(k.end(), b, b + sizeof(T));
}

template <typename Body>
int graph_or_eager(pre3_ctx* ctx, const std::vector<unsigned char>& key, Body body) {
  // the legacy default stream cannot be captured: run eagerly there
  if (!ctx->graphs || ctx->timing || ctx->stream == nullptr || ctx->stream == cudaStreamLegacy ||
      ctx->stream == cudaStreamPerThread)
    return body();
  if (ctx->graph_exec && key == ctx->graph_key) {
    PRE3_CUDA(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
    ctx->launches += ctx->graph_launches;
    return PRE3_OK;
  }
  if (key != ctx->graph_seen) {  // first sight: eager
    ctx->graph_seen = key;
    return body();
  }
  // second call with this signature: capture
  if (ctx->graph_exec) {
    cudaGraphExecDestroy(ctx->graph_exec);
    ctx->graph_exec = nullptr;
  }
  const int64_t l0 = ctx->launches;
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();  // this stream cannot be captured (e.g. it is itself being captured by the caller)
    ctx->graph_seen.clear();
    return body();
  }
  const int rc = body();
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
  if (rc != PRE3_OK || e != cudaSuccess || !g) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    ctx->graph_seen.clear();
    if (rc != PRE3_OK) return rc;
    return body();  // not capturable in this state: run eagerly
  }
  const cudaError_t ei = cudaGraphInstantiate(&ctx->graph_exec, g, 0);
  cudaGraphDestroy(g);
  if (ei != cudaSuccess) {
    ctx->graph_exec = nullptr;
    cudaGetLastError();
    return body();
  }
  ctx->graph_key = key;
  ctx->graph_launches = ctx->launches - l0;
  PRE3_CUDA(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
  return PRE3_OK;
}
}  // namespace

namespace {
// ---- software pipeline over chunks of pairs (pre3_set_pipeline) ---------------------------------------------------
// The stages of the path are bound by different units: convert by HBM, the descriptor GEMM by the tensor cores and the
// shared-memory operand stream, the hypothesis evaluation by FP32 issue slots, the selection by dependent fp64 chains.
// Run back to back each leaves the other units idle.  The pairs of a call are cut into chunks, every stage gets its own
// stream, and chunk c + 1 enters a stage as soon as chunk c has left it: kernels of different stages then share the SMs
// (a persistent GEMM CTA with 352 threads leaves room for convert / evaluation / selection blocks beside it).  Each
// chunk carves its own part of the arena, so the chunks have no buffers in common; results are those of the unchunked
// call (every pair is independent: tests/test_gpu_parity.py::test_pipeline_*).  In sequence mode the frame on a chunk
// boundary is converted by both chunks.
static int ensure_pipe_streams(pre3_ctx* ctx) {
  if (ctx->pipe_fork) return PRE3_OK;
  int lo = 0, hi = 0;  // numerically lower = higher priority
  PRE3_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  // later stages drain first: a chunk that has reached the selection should not wait behind the conversion of the next
  const int use_prio = getenv("PRE3_PIPE_PRIO") ? atoi(getenv("PRE3_PIPE_PRIO")) : 1;
  for (int i = 0; i < 4; ++i) {
    const int pr = use_prio ? std::max(hi, lo - i) : lo;
    PRE3_CUDA(cudaStreamCreateWithPriority(&ctx->pipe_stream[i], cudaStreamNonBlocking, pr));
    PRE3_CUDA(cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming));
  }
  PRE3_CUDA(cudaEventCreateWithFlags(&ctx->pipe_fork, cudaEventDisableTiming));
  if (const char* m = getenv("PRE3_PIPE_MAP")) {  // e.g. "0122": evaluation and selection on one stream
    for (int i = 0; i < 4 && m[i] >= '0' && m[i] <= '3'; ++i) ctx->pipe_map[i] = m[i] - '0';
  }
  return PRE3_OK;
}

// chunks a call of P pairs is cut into (1: no pipeline)
static int pipe_chunk_count(const pre3_ctx* ctx, int P, int K1, int K2) {
  if (ctx->timing || ctx->pipe_chunks == 0 || ctx->pipe_chunks == 1) return 1;
  if (K1 > 2048 || K2 > 2048) return 1;  // large pairs: the per-call adaptive rows are built with a host round trip
  static const int env = getenv("PRE3_PIPE_CHUNKS") ? atoi(getenv("PRE3_PIPE_CHUNKS")) : -1;
  int n = env >= 0 ? env : ctx->pipe_chunks;
  if (n < 0) n = 1;  // automatic = off: measured slower at 4096 and at 512 pairs (profiles/r02_pipeline_sweep.log)
  const int min_pairs = 128;  // a chunk must keep the batch forms of the kernels (>= 64 pairs) and fill the GEMM grid
  n = std::min(n, P / min_pairs);
  return std::max(n, 1);
}

// Body(p0, n): issue the launches of pairs [p0, p0 + n) on ctx->stream (stage switches through pipe_enter).
template <typename Body>
static int run_pipelined(pre3_ctx* ctx, int P, int nchunks, Body body) {
  if (nchunks <= 1) return body(0, P);
  PRE3_TRY(ensure_pipe_streams(ctx));
  cudaStream_t main_stream = ctx->stream;
  PRE3_CUDA(cudaEventRecord(ctx->pipe_fork, main_stream));
  for (int i = 0; i < 4; ++i) PRE3_CUDA(cudaStreamWaitEvent(ctx->pipe_stream[i], ctx->pipe_fork, 0));
  ctx->pipe_total_P = P;
  int rc = PRE3_OK;
  for (int c = 0; c < nchunks && rc == PRE3_OK; ++c) {
    const int p0 = (int)((long long)P * c / nchunks), p1 = (int)((long long)P * (c + 1) / nchunks);
    ctx->stream = ctx->pipe_stream[ctx->pipe_map[0]];  // a new chunk starts at the first stage, behind nothing but
    ctx->pipe_stage = 0;                               // the previous chunk's launches on that stream
    rc = body(p0, p1 - p0);
  }
  ctx->pipe_stage = -1;
  ctx->pipe_total_P = 0;
  ctx->stream = main_stream;
  // join (also on failure: a capture must not be left with unjoined streams)
  for (int i = 0; i < 4; ++i) {
    cudaEventRecord(ctx->pipe_ev[i], ctx->pipe_stream[i]);
    cudaStreamWaitEvent(main_stream, ctx->pipe_ev[i], 0);
  }
  return rc;
}

}  // namespace

// ================================================================================================
// context
// ================================================================================================
extern "C" {

const char* pre3_version(void) { return "pre3-b200 0.1 (sm_100a)"; }

int pre3_create(pre3_ctx** out, int device) {
  if (!out) return PRE3_ERR_ARG;
  *out = nullptr;
  pre3_ctx* ctx = new pre3_ctx();
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    // keep the context alive so that pre3_last_error() can report why
    ctx->err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
               " (libpre3 has no CPU fallback)";
    ctx->device = -1;
    *out = ctx;
    return PRE3_ERR_CUDA;
  }
  if (device < 0) {
    e = cudaGetDevice(&device);
    if (e != cudaSuccess) device = 0;
  }
  if (device >= ndev) {
    ctx->err = "device index out of range";
    ctx->device = -1;
    *out = ctx;
    return PRE3_ERR_ARG;
  }
  ctx->device = device;
  *out = ctx;
  PRE3_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PRE3_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    ctx->err = std::string("libpre3 is built for sm_100a only; device is sm_") + std::to_string(prop.major) +
               std::to_string(prop.minor);
    ctx->device = -1;
    return PRE3_ERR_CUDA;
  }
  PRE3_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = true;
  return PRE3_OK;
}

void pre3_destroy(pre3_ctx* ctx) {
  if (!ctx) return;
  if (ctx->device >= 0) {
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->ws) cudaFree(ctx->ws);
    for (char* r : ctx->retired) cudaFree(r);
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    if (ctx->d_tab) cudaFree(ctx->d_tab);
    if (ctx->d_tab_rows) cudaFree(ctx->d_tab_rows);
    if (ctx->d_ekf_tab) cudaFree(ctx->d_ekf_tab);
    for (int i = 0; i < 2; ++i) {
      if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
      if (ctx->ev_stage_done[i]) cudaEventDestroy(ctx->ev_stage_done[i]);
    }
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    for (void* p : ctx->aux) cudaFree(p);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 4; ++i) {
      if (ctx->pipe_stream[i]) cudaStreamDestroy(ctx->pipe_stream[i]);
      if (ctx->pipe_ev[i]) cudaEventDestroy(ctx->pipe_ev[i]);
    }
    if (ctx->pipe_fork) cudaEventDestroy(ctx->pipe_fork);
    for (int i = 0; i < 2; ++i) {
      if (ctx->ev_slot_full[i]) cudaEventDestroy(ctx->ev_slot_full[i]);
      if (ctx->ev_slot_free[i]) cudaEventDestroy(ctx->ev_slot_free[i]);
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  }
  if (ctx->pool) host_pool_destroy(ctx->pool);
  delete ctx;
}

const char* pre3_last_error(const pre3_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int pre3_set_stream(pre3_ctx* ctx, void* cuda_stream) {
  if (!ctx || ctx->device < 0) return PRE3_ERR_CUDA;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return PRE3_OK;
}

int pre3_set_match_engine(pre3_ctx* ctx, int engine) {
  if (!ctx) return PRE3_ERR_ARG;
  if (engine < PRE3_MATCH_AUTO || engine > PRE3_MATCH_TC) return fail(ctx, PRE3_ERR_ARG, "unknown match engine");
  ctx->match_engine = engine;
  return PRE3_OK;
}

int pre3_sync(pre3_ctx* ctx) {
  if (!ctx || ctx->device < 0) return ctx ? fail(ctx, PRE3_ERR_CUDA, "no CUDA device") : PRE3_ERR_ARG;
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

int64_t pre3_launch_count(const pre3_ctx* ctx) { return ctx ? ctx->launches : 0; }

int pre3_transfer_bytes(const pre3_ctx* ctx, int64_t* h2d, int64_t* d2h) {
  if (!ctx) return PRE3_ERR_ARG;
  if (h2d) *h2d = ctx->h2d_bytes;
  if (d2h) *d2h = ctx->d2h_bytes;
  return PRE3_OK;
}

int pre3_eval_schedule(const pre3_ransac_opts* opts, int32_t* ends, int cap) {
  if (!opts || (cap > 0 && !ends)) return PRE3_ERR_ARG;
  return eval_wave_ends(*opts, ends, cap);
}

int pre3_eval_schedule_for(const pre3_ransac_opts* opts, int P, int32_t* ends, int cap) {
  if (!opts || P < 1 || (cap > 0 && !ends)) return PRE3_ERR_ARG;
  return eval_wave_ends(*opts, ends, cap, P);
}

int pre3_timing_enable(pre3_ctx* ctx, int on) {
  if (!ctx || ctx->device < 0) return PRE3_ERR_CUDA;
  ctx->timing = on != 0;
  return PRE3_OK;
}

int pre3_timing_read(pre3_ctx* ctx, double* ms, int64_t* count) {
  if (!ctx || ctx->device < 0) return PRE3_ERR_CUDA;
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  for (const TimedSpan& s : ctx->spans) {
    float t = 0.f;
    PRE3_CUDA(cudaEventElapsedTime(&t, s.a, s.b));
    if (ms) ms[s.cat] += (double)t;
    if (count) count[s.cat] += 1;
  }
  ctx->spans.clear();
  ctx->ev_used = 0;
  return PRE3_OK;
}

const char* pre3_timing_name(int cat) {
  static const char* names[T_NCAT] = {"convert", "match_tc", "match_exact", "rescore", "compact",
                                      "prep",    "eval",     "select",      "other",   "ekf_gain",
                                      "ekf_score", "ekf_select", "frames", "ekf_update", "match_fused"};
  return (cat >= 0 && cat < T_NCAT) ? names[cat] : "?";
}

#define PRE3_LIVE()                                                                          \
  do {                                                                                       \
    if (!ctx) return PRE3_ERR_ARG;                                                           \
    if (ctx->device < 0) return fail(ctx, PRE3_ERR_CUDA, "no CUDA device (libpre3 has no CPU fallback)"); \
    PRE3_CUDA(cudaSetDevice(ctx->device));                                                   \
  } while (0)

// ================================================================================================
// stage 1
// ================================================================================================
int pre3_siftmatch_batch_dev(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int P, int K1, int K2,
                             int ND, const int32_t* dk1_count, const int32_t* dk2_count, double thresh,
                             int32_t* dpairs, double* dscore, int32_t* dn_out) {
  PRE3_LIVE();
  PRE3_TRY(check_match_args(ctx, dL1, dL2, cls, P, K1, K2, ND));
  PRE3_NEED(dn_out != nullptr, "n_out missing");
  if (P == 0) return PRE3_OK;
  PRE3_TRY(ws_reserve(ctx, match_ws_bytes(cls, P, K1, K2, ND)));
  MatchRow* rows = nullptr;
  if (K1 > 0)
    PRE3_TRY(match_impl(ctx, dL1, dL2, cls, P, K1, K2, ND, dk1_count, dk2_count, thresh, dscore != nullptr, &rows));
  PRE3_TRY(launch_match_compact(ctx, rows, P, K1, dk1_count, dpairs, dscore, dn_out, nullptr, nullptr, K2, nullptr,
                                nullptr));
  return PRE3_OK;
}

int pre3_siftmatch_batch(pre3_ctx* ctx, const void* L1, const void* L2, int cls, int P, int K1, int K2, int ND,
                         const int32_t* k1_count, const int32_t* k2_count, double thresh, int32_t* pairs,
                         double* score, int32_t* n_out) {
  PRE3_LIVE();
  PRE3_TRY(check_match_args(ctx, L1, L2, cls, P, K1, K2, ND));
  PRE3_NEED(n_out != nullptr, "n_out missing");
  if (P == 0) return PRE3_OK;
  const size_t es = class_size(cls);
  const size_t b1 = (size_t)P * K1 * ND * es, b2 = (size_t)P * K2 * ND * es;
  const size_t np = (size_t)P * K1;
  size_t need = match_ws_bytes(cls, P, K1, K2, ND) + align_up(b1) + align_up(b2) + 2 * align_up(4 * (size_t)P) +
                align_up(8 * np) + align_up(8 * np) + align_up(4 * (size_t)P) + 4096;
  PRE3_TRY(ws_reserve(ctx, need));
  char* d1 = ws_take<char>(ctx, b1);
  char* d2 = ws_take<char>(ctx, b2);
  int32_t* dk1 = k1_count ? ws_take<int32_t>(ctx, P) : nullptr;
  int32_t* dk2 = k2_count ? ws_take<int32_t>(ctx, P) : nullptr;
  int32_t* dpairs = ws_take<int32_t>(ctx, 2 * np);
  double* dscore = ws_take<double>(ctx, np);
  int32_t* dn = ws_take<int32_t>(ctx, P);
  PRE3_TRY(h2d(ctx, d1, L1, b1));
  PRE3_TRY(h2d(ctx, d2, L2, b2));
  if (dk1) PRE3_TRY(h2d(ctx, dk1, k1_count, 4 * (size_t)P));
  if (dk2) PRE3_TRY(h2d(ctx, dk2, k2_count, 4 * (size_t)P));
  MatchRow* rows = nullptr;
  if (K1 > 0) PRE3_TRY(match_impl(ctx, d1, d2, cls, P, K1, K2, ND, dk1, dk2, thresh, score != nullptr, &rows));
  PRE3_TRY(launch_match_compact(ctx, rows, P, K1, dk1, dpairs, dscore, dn, nullptr, nullptr, K2, nullptr, nullptr));
  if (pairs) PRE3_TRY(d2h(ctx, pairs, dpairs, 8 * np));
  if (score) PRE3_TRY(d2h(ctx, score, dscore, 8 * np));
  PRE3_TRY(d2h(ctx, n_out, dn, 4 * (size_t)P));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

// ONE descriptor set L1 against P sets L2: the sweep of find_consistent_sift_matches.m:39-65 (the first frame's good
// descriptors against every later frame).  L1 is converted for the tensor cores once and stays L2-resident; the rest is
// the batched matcher.
int pre3_siftmatch_sweep_dev(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int P, int K1, int K2, int ND,
                             const int32_t* dk2_count, double thresh, int32_t* dpairs, double* dscore, int32_t* dn_out) {
  PRE3_LIVE();
  PRE3_TRY(check_match_args(ctx, dL1, dL2, cls, P, K1, K2, ND));
  PRE3_NEED(dn_out != nullptr, "n_out missing");
  if (P == 0) return PRE3_OK;
  PRE3_TRY(ws_reserve(ctx, match_ws_bytes(cls, P, K1, K2, ND)));
  MatchRow* rows = nullptr;
  ctx->l1_shared = 1;
  int rc = PRE3_OK;
  if (K1 > 0) rc = match_impl(ctx, dL1, dL2, cls, P, K1, K2, ND, nullptr, dk2_count, thresh, dscore != nullptr, &rows);
  ctx->l1_shared = 0;
  PRE3_TRY(rc);
  return launch_match_compact(ctx, rows, P, K1, nullptr, dpairs, dscore, dn_out, nullptr, nullptr, K2, nullptr, nullptr);
}

int pre3_siftmatch_sweep(pre3_ctx* ctx, const void* L1, const void* L2, int cls, int P, int K1, int K2, int ND,
                         const int32_t* k2_count, double thresh, int32_t* pairs, double* score, int32_t* n_out) {
  PRE3_LIVE();
  PRE3_TRY(check_match_args(ctx, L1, L2, cls, P, K1, K2, ND));
  PRE3_NEED(n_out != nullptr, "n_out missing");
  if (P == 0) return PRE3_OK;
  const size_t es = class_size(cls);
  const size_t b1 = (size_t)K1 * ND * es, b2 = (size_t)P * K2 * ND * es;
  const size_t np = (size_t)P * K1;
  PRE3_TRY(ws_reserve(ctx, match_ws_bytes(cls, P, K1, K2, ND) + align_up(b1) + align_up(b2) + align_up(4 * (size_t)P) +
                               2 * align_up(8 * np) + align_up(4 * (size_t)P) + 4096));
  char* d1 = ws_take<char>(ctx, b1);
  char* d2 = ws_take<char>(ctx, b2);
  int32_t* dk2 = k2_count ? ws_take<int32_t>(ctx, P) : nullptr;
  int32_t* dpairs = ws_take<int32_t>(ctx, 2 * np);
  double* dscore = ws_take<double>(ctx, np);
  int32_t* dn = ws_take<int32_t>(ctx, P);
  PRE3_TRY(h2d(ctx, d1, L1, b1));
  PRE3_TRY(h2d(ctx, d2, L2, b2));
  if (dk2) PRE3_TRY(h2d(ctx, dk2, k2_count, 4 * (size_t)P));
  MatchRow* rows = nullptr;
  ctx->l1_shared = 1;
  int rc = PRE3_OK;
  if (K1 > 0) rc = match_impl(ctx, d1, d2, cls, P, K1, K2, ND, nullptr, dk2, thresh, score != nullptr, &rows);
  ctx->l1_shared = 0;
  PRE3_TRY(rc);
  PRE3_TRY(launch_match_compact(ctx, rows, P, K1, nullptr, dpairs, dscore, dn, nullptr, nullptr, K2, nullptr, nullptr));
  if (pairs) PRE3_TRY(d2h(ctx, pairs, dpairs, 8 * np));
  if (score) PRE3_TRY(d2h(ctx, score, dscore, 8 * np));
  PRE3_TRY(d2h(ctx, n_out, dn, 4 * (size_t)P));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

// matching_sift_based.m:104-135 for P frames: des1 = the descriptors of the PREDICTED map features (F per frame,
// f_count valid), des2 = Descriptor_RAW of the frame (K2, k2_count valid), h = predicted image positions (2 x F),
// S11 = S(1,1) of every predicted feature (NaN: S empty -> radius 40), pos2 = SCALE_ORIENT_POS_RAW(1:2,:) (2 x K2).
// siftmatch (default threshold 1.5) followed by the search-region gate; outputs per predicted feature.
int pre3_matching_sift_based_batch(pre3_ctx* ctx, const void* des1, const void* des2, int cls, int P, int F, int K2,
                                   int ND, const int32_t* f_count, const int32_t* k2_count, const double* h,
                                   const double* S11, const double* pos2, double thresh, uint8_t* ic, double* z,
                                   int32_t* match, int32_t* n_match, int32_t* n_discarded) {
  PRE3_LIVE();
  PRE3_TRY(check_match_args(ctx, des1, des2, cls, P, F, K2, ND));
  PRE3_NEED(P == 0 || F == 0 || (h && S11 && ic && z && match), "null pointer");
  PRE3_NEED(P == 0 || K2 == 0 || pos2, "null pointer");
  if (P == 0) return PRE3_OK;
  const size_t es = class_size(cls);
  const size_t b1 = (size_t)P * F * ND * es, b2 = (size_t)P * K2 * ND * es;
  const size_t np = (size_t)P * F;
  PRE3_TRY(ws_reserve(ctx, match_ws_bytes(cls, P, F, K2, ND) + align_up(b1) + align_up(b2) + 4 * align_up(4 * (size_t)P) +
                               align_up(8 * np) + 2 * align_up(16 * np) + align_up(8 * np) + align_up(16 * (size_t)P * K2) +
                               align_up(np) + align_up(4 * np) + 8192));
  char* d1 = ws_take<char>(ctx, b1);
  char* d2 = ws_take<char>(ctx, b2);
  int32_t* dk1 = f_count ? ws_take<int32_t>(ctx, P) : nullptr;
  int32_t* dk2 = k2_count ? ws_take<int32_t>(ctx, P) : nullptr;
  int32_t* dpairs = ws_take<int32_t>(ctx, 2 * np);
  int32_t* dn = ws_take<int32_t>(ctx, P);
  int32_t* dnd = ws_take<int32_t>(ctx, P);
  double* dh = ws_take<double>(ctx, 2 * np);
  double* dz = ws_take<double>(ctx, 2 * np);
  double* dS = ws_take<double>(ctx, np);
  double* dpos = ws_take<double>(ctx, 2 * (size_t)P * K2);
  uint8_t* dic = ws_take<uint8_t>(ctx, np);
  int32_t* dm = ws_take<int32_t>(ctx, np);
  PRE3_TRY(h2d(ctx, d1, des1, b1));
  PRE3_TRY(h2d(ctx, d2, des2, b2));
  if (dk1) PRE3_TRY(h2d(ctx, dk1, f_count, 4 * (size_t)P));
  if (dk2) PRE3_TRY(h2d(ctx, dk2, k2_count, 4 * (size_t)P));
  PRE3_TRY(h2d(ctx, dh, h, 16 * np));
  PRE3_TRY(h2d(ctx, dS, S11, 8 * np));
  PRE3_TRY(h2d(ctx, dpos, pos2, 16 * (size_t)P * K2));
  MatchRow* rows = nullptr;
  if (F > 0) PRE3_TRY(match_impl(ctx, d1, d2, cls, P, F, K2, ND, dk1, dk2, thresh, 0, &rows));
  PRE3_TRY(launch_match_compact(ctx, rows, P, F, dk1, dpairs, nullptr, dn, nullptr, nullptr, K2, nullptr, nullptr));
  PRE3_TRY(launch_match_gate(ctx, dpairs, dn, P, F, K2, dh, dS, dpos, dic, dz, dm, dnd));
  PRE3_TRY(d2h(ctx, ic, dic, np));
  PRE3_TRY(d2h(ctx, z, dz, 16 * np));
  PRE3_TRY(d2h(ctx, match, dm, 4 * np));
  if (n_match) PRE3_TRY(d2h(ctx, n_match, dn, 4 * (size_t)P));
  if (n_discarded) PRE3_TRY(d2h(ctx, n_discarded, dnd, 4 * (size_t)P));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

int pre3_siftmatch(pre3_ctx* ctx, const void* L1, const void* L2, int cls, int K1, int K2, int ND, double thresh,
                   int32_t* pairs, double* score, int32_t* n_out) {
  return pre3_siftmatch_batch(ctx, L1, L2, cls, 1, K1, K2, ND, nullptr, nullptr, thresh, pairs, score, n_out);
}

// ================================================================================================
// stage 2
// ================================================================================================
static int fit_all_host(pre3_ctx* ctx, const double* p1, const double* p2, int n, int method, int do_scale,
                        double* out15) {
  const size_t pb = 3 * (size_t)n * sizeof(double);
  PRE3_TRY(ws_reserve(ctx, 2 * align_up(pb) + 4096));
  double* d1 = ws_take<double>(ctx, 3 * (size_t)n);
  double* d2 = ws_take<double>(ctx, 3 * (size_t)n);
  double* dout = ws_take<double>(ctx, 16);
  PRE3_TRY(h2d(ctx, d1, p1, pb));
  PRE3_TRY(h2d(ctx, d2, p2, pb));
  PRE3_TRY(launch_fit_all(ctx, d1, d2, n, method, do_scale, dout));
  PRE3_TRY(d2h(ctx, out15, dout, 15 * sizeof(double)));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

int pre3_find_transform_matrix(pre3_ctx* ctx, const double* pset1, const double* pset2, int n, double* rot,
                               double* trans, int32_t* state) {
  PRE3_LIVE();
  PRE3_NEED(pset1 && pset2 && rot && trans && state, "null pointer");
  PRE3_NEED(n >= 0, "negative point count");
  double out[15];
  PRE3_TRY(fit_all_host(ctx, pset1, pset2, n, PRE3_METHOD_SVD, 0, out));
  memcpy(rot, out, 9 * sizeof(double));
  memcpy(trans, out + 9, 3 * sizeof(double));
  *state = (int32_t)out[12];
  return PRE3_OK;
}

int pre3_horn(pre3_ctx* ctx, const double* A, const double* B, int n, int do_scale, double* s, double* R, double* T,
              double* err) {
  PRE3_LIVE();
  PRE3_NEED(A && B && R && T, "null pointer");
  PRE3_NEED(n >= 4, "Need at least 4 point pairs");  // absoluteOrientationQuaternion.m:51-54
  double out[15];
  PRE3_TRY(fit_all_host(ctx, A, B, n, PRE3_METHOD_HORN, do_scale ? 1 : 0, out));
  memcpy(R, out, 9 * sizeof(double));
  memcpy(T, out + 9, 3 * sizeof(double));
  if (s) *s = out[13];
  if (err) *err = out[14];
  return PRE3_OK;
}

// Explicit sample sets index the correspondences: an index outside 0..N-1 is MATLAB's "Index exceeds matrix dimensions"
// (Ya(:, idx)), reported here instead of being clamped.  Pairs with fewer than k correspondences never sample.
static int check_samples(pre3_ctx* ctx, const int32_t* samples, int P, int H, int k, const int32_t* n_corr, int Nmax) {
  if (!samples) return PRE3_OK;
  for (int p = 0; p < P; ++p) {
    const int N = n_corr ? std::min(n_corr[p], Nmax) : Nmax;
    if (N < k) continue;
    const int32_t* sp = samples + (size_t)p * H * k;
    for (size_t i = 0; i < (size_t)H * k; ++i)
      if (sp[i] < 0 || sp[i] >= N)
        return fail(ctx, PRE3_ERR_ARG, "sample index out of range (Index exceeds matrix dimensions): pair " +
                                           std::to_string(p) + ", set " + std::to_string(i / k) + ", value " +
                                           std::to_string(sp[i]) + " with N = " + std::to_string(N));
  }
  return PRE3_OK;
}

int pre3_fit_batch(pre3_ctx* ctx, const double* Ya, const double* Yb, int N, const int32_t* samples, int k, int H,
                   int method, double* R, double* T, int32_t* state) {
  PRE3_LIVE();
  PRE3_NEED(Ya && Yb && samples && R && T && state, "null pointer");
  PRE3_NEED(N >= 1 && H >= 0, "bad sizes");
  PRE3_NEED(k >= 3 && k <= 8, "minimal sample size k must be in 3..8");
  PRE3_NEED(method == PRE3_METHOD_SVD || method == PRE3_METHOD_HORN, "unknown method");
  if (H == 0) return PRE3_OK;
  PRE3_TRY(check_samples(ctx, samples, 1, H, k, nullptr, N));
  const size_t pb = 3 * (size_t)N * 8, sb = 4 * (size_t)k * H;
  PRE3_TRY(ws_reserve(ctx, 2 * align_up(pb) + align_up(sb) + align_up(72 * (size_t)H) + align_up(24 * (size_t)H) +
                               align_up(4 * (size_t)H) + 4096));
  double* dYa = ws_take<double>(ctx, 3 * (size_t)N);
  double* dYb = ws_take<double>(ctx, 3 * (size_t)N);
  int32_t* ds = ws_take<int32_t>(ctx, (size_t)k * H);
  double* dR = ws_take<double>(ctx, 9 * (size_t)H);
  double* dT = ws_take<double>(ctx, 3 * (size_t)H);
  int32_t* dst = ws_take<int32_t>(ctx, H);
  PRE3_TRY(h2d(ctx, dYa, Ya, pb));
  PRE3_TRY(h2d(ctx, dYb, Yb, pb));
  PRE3_TRY(h2d(ctx, ds, samples, sb));
  PRE3_TRY(launch_fit_only(ctx, dYa, dYb, N, ds, k, H, method, dR, dT, dst));
  PRE3_TRY(d2h(ctx, R, dR, 72 * (size_t)H));
  PRE3_TRY(d2h(ctx, T, dT, 24 * (size_t)H));
  PRE3_TRY(d2h(ctx, state, dst, 4 * (size_t)H));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

// ================================================================================================
// stage 3
// ================================================================================================
int pre3_score_batch(pre3_ctx* ctx, const double* R, const double* T, int H, const double* Ya, const double* Yb,
                     int N, double thr, int32_t* count, double* errsum, uint8_t* mask) {
  PRE3_LIVE();
  PRE3_NEED(R && T && Ya && Yb, "null pointer");
  PRE3_NEED(N >= 0 && H >= 0, "bad sizes");
  if (H == 0) return PRE3_OK;
  const size_t pb = 3 * (size_t)N * 8;
  const size_t mb = mask ? (size_t)N * H : 0;
  PRE3_TRY(ws_reserve(ctx, 2 * align_up(pb) + align_up(72 * (size_t)H) + align_up(24 * (size_t)H) +
                               align_up(4 * (size_t)H) + align_up(8 * (size_t)H) + align_up(mb) +
                               2 * align_up(16 * (size_t)N) + 8192));
  double* dYa = ws_take<double>(ctx, 3 * (size_t)N);
  double* dYb = ws_take<double>(ctx, 3 * (size_t)N);
  double* dR = ws_take<double>(ctx, 9 * (size_t)H);
  double* dT = ws_take<double>(ctx, 3 * (size_t)H);
  int32_t* dc = ws_take<int32_t>(ctx, H);
  double* de = ws_take<double>(ctx, H);
  uint8_t* dm = mask ? ws_take<uint8_t>(ctx, mb) : nullptr;
  PRE3_TRY(h2d(ctx, dYa, Ya, pb));
  PRE3_TRY(h2d(ctx, dYb, Yb, pb));
  PRE3_TRY(h2d(ctx, dR, R, 72 * (size_t)H));
  PRE3_TRY(h2d(ctx, dT, T, 24 * (size_t)H));
  PRE3_TRY(launch_score_given(ctx, dR, dT, H, dYa, dYb, N, thr, dc, errsum ? de : nullptr, dm));
  if (count) PRE3_TRY(d2h(ctx, count, dc, 4 * (size_t)H));
  if (errsum) PRE3_TRY(d2h(ctx, errsum, de, 8 * (size_t)H));
  if (mask) PRE3_TRY(d2h(ctx, mask, dm, mb));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

// ================================================================================================
// stages 2-4
// ================================================================================================
int pre3_ransac_batch_dev(pre3_ctx* ctx, const double* dYa, const double* dYb, const int32_t* dn_corr, int P,
                          int Nmax, const pre3_ransac_opts* opts, const int32_t* dsamples, pre3_pair_result* dres,
                          uint8_t* dmasks) {
  PRE3_LIVE();
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(P >= 0 && Nmax >= 0, "bad sizes");
  PRE3_NEED(dres != nullptr, "result pointer missing");
  if (P == 0) return PRE3_OK;
  PRE3_TRY(ws_reserve(ctx, ransac_ws_bytes(P, Nmax, opts->H)));
  return ransac_impl(ctx, dYa, dYb, dn_corr, P, Nmax, *opts, dsamples, 0, dres, dmasks, nullptr, nullptr);
}

static int ransac_batch_host(pre3_ctx* ctx, const double* Ya, const double* Yb, const int32_t* n_corr, int P,
                             int Nmax, const pre3_ransac_opts* opts, const int32_t* samples, pre3_pair_result* res,
                             uint8_t* masks, int32_t* counts, int8_t* states) {
  PRE3_LIVE();
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(P >= 0 && Nmax >= 0, "bad sizes");
  PRE3_NEED(res != nullptr, "result pointer missing");
  PRE3_NEED((Ya && Yb) || (size_t)P * Nmax == 0, "null correspondences");
  if (P == 0) return PRE3_OK;
  const int H = opts->H;
  PRE3_TRY(check_samples(ctx, samples, P, H, opts->k, n_corr, Nmax));
  const size_t pb = 3 * (size_t)P * Nmax * 8;
  const size_t sb = samples ? 4 * (size_t)P * opts->k * H : 0;
  const size_t rb = sizeof(pre3_pair_result) * (size_t)P;
  const size_t mb = masks ? (size_t)P * Nmax : 0;
  const size_t cb = counts ? 4 * (size_t)P * H : 0, stb = states ? (size_t)P * H : 0;
  PRE3_TRY(ws_reserve(ctx, ransac_ws_bytes(P, Nmax, H) + 2 * align_up(pb) + align_up(sb) + align_up(rb) +
                               align_up(mb) + align_up(cb) + align_up(stb) + align_up(4 * (size_t)P) + 8192));
  double* dYa = ws_take<double>(ctx, 3 * (size_t)P * Nmax);
  double* dYb = ws_take<double>(ctx, 3 * (size_t)P * Nmax);
  int32_t* ds = samples ? ws_take<int32_t>(ctx, (size_t)P * opts->k * H) : nullptr;
  int32_t* dn = n_corr ? ws_take<int32_t>(ctx, P) : nullptr;
  pre3_pair_result* dres = ws_take<pre3_pair_result>(ctx, P);
  uint8_t* dm = masks ? ws_take<uint8_t>(ctx, mb) : nullptr;
  int32_t* dc = counts ? ws_take<int32_t>(ctx, (size_t)P * H) : nullptr;
  int8_t* dst = states ? ws_take<int8_t>(ctx, (size_t)P * H) : nullptr;
  PRE3_TRY(h2d(ctx, dYa, Ya, pb));
  PRE3_TRY(h2d(ctx, dYb, Yb, pb));
  if (ds) PRE3_TRY(h2d(ctx, ds, samples, sb));
  if (dn) PRE3_TRY(h2d(ctx, dn, n_corr, 4 * (size_t)P));
  PRE3_TRY(ransac_impl(ctx, dYa, dYb, dn, P, Nmax, *opts, ds, 0, dres, dm, dc, dst));
  PRE3_TRY(d2h(ctx, res, dres, rb));
  if (masks) PRE3_TRY(d2h(ctx, masks, dm, mb));
  if (counts) PRE3_TRY(d2h(ctx, counts, dc, cb));
  if (states) PRE3_TRY(d2h(ctx, states, dst, stb));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

int pre3_ransac_batch(pre3_ctx* ctx, const double* Ya, const double* Yb, const int32_t* n_corr, int P, int Nmax,
                      const pre3_ransac_opts* opts, const int32_t* samples, pre3_pair_result* res, uint8_t* masks) {
  return ransac_batch_host(ctx, Ya, Yb, n_corr, P, Nmax, opts, samples, res, masks, nullptr, nullptr);
}

int pre3_ransac(pre3_ctx* ctx, const double* Ya, const double* Yb, int N, const pre3_ransac_opts* opts,
                const int32_t* samples, pre3_pair_result* res, uint8_t* mask, int32_t* counts, int8_t* states) {
  return ransac_batch_host(ctx, Ya, Yb, nullptr, 1, N, opts, samples, res, mask, counts, states);
}

// ================================================================================================
// code_from_dr_ye variant
// ================================================================================================
int pre3_vodometry_dr_ye_batch_dev(pre3_ctx* ctx, const double* dYa, const double* dYb, const int32_t* dn_corr,
                                   const int32_t* dmatch, int P, int Nmax, const pre3_ransac_opts* opts,
                                   const int32_t* dsamples, uint32_t pair_id0, pre3_pair_result* dres,
                                   uint8_t* dmasks, pre3_dr_ye_stat* dstat, int32_t* dcounts) {
  PRE3_LIVE();
  PRE3_NEED(opts != nullptr, "options missing");
  pre3_ransac_opts o = *opts;
  o.method = PRE3_METHOD_DR_YE;
  PRE3_TRY(check_opts(ctx, &o));
  PRE3_NEED(P >= 0 && Nmax >= 0, "bad sizes");
  PRE3_NEED(dres != nullptr, "result pointer missing");
  if (P == 0) return PRE3_OK;
  PRE3_TRY(ws_reserve(ctx, ransac_ws_bytes(P, Nmax, o.H)));
  return ransac_impl(ctx, dYa, dYb, dn_corr, P, Nmax, o, dsamples, pair_id0, dres, dmasks, dcounts, nullptr, dmatch,
                     dstat);
}

int pre3_vodometry_dr_ye_batch(pre3_ctx* ctx, const double* Ya, const double* Yb, const int32_t* n_corr,
                               const int32_t* match, int P, int Nmax, const pre3_ransac_opts* opts,
                               const int32_t* samples, pre3_pair_result* res, uint8_t* masks, pre3_dr_ye_stat* stat,
                               int32_t* counts) {
  PRE3_LIVE();
  PRE3_NEED(opts != nullptr, "options missing");
  pre3_ransac_opts o = *opts;
  o.method = PRE3_METHOD_DR_YE;
  PRE3_TRY(check_opts(ctx, &o));
  PRE3_NEED(P >= 0 && Nmax >= 0, "bad sizes");
  PRE3_NEED(res != nullptr, "result pointer missing");
  PRE3_NEED((Ya && Yb) || (size_t)P * Nmax == 0, "null correspondences");
  if (P == 0) return PRE3_OK;
  const int H = o.H;
  PRE3_TRY(check_samples(ctx, samples, P, H, 4, n_corr, Nmax));
  const size_t pb = 3 * (size_t)P * Nmax * 8;
  const size_t sb = samples ? 16 * (size_t)P * H : 0;
  const size_t mtb = match ? 8 * (size_t)P * Nmax : 0;
  const size_t rb = sizeof(pre3_pair_result) * (size_t)P;
  const size_t mb = masks ? (size_t)P * Nmax : 0;
  const size_t cb = counts ? 4 * (size_t)P * H : 0;
  const size_t stb = stat ? sizeof(pre3_dr_ye_stat) * (size_t)P : 0;
  PRE3_TRY(ws_reserve(ctx, ransac_ws_bytes(P, Nmax, H) + 2 * align_up(pb) + align_up(sb) + align_up(mtb) + align_up(rb) +
                               align_up(mb) + align_up(cb) + align_up(stb) + align_up(4 * (size_t)P) + 8192));
  double* dYa = ws_take<double>(ctx, 3 * (size_t)P * Nmax);
  double* dYb = ws_take<double>(ctx, 3 * (size_t)P * Nmax);
  int32_t* ds = samples ? ws_take<int32_t>(ctx, 4 * (size_t)P * H) : nullptr;
  int32_t* dmt = match ? ws_take<int32_t>(ctx, 2 * (size_t)P * Nmax) : nullptr;
  int32_t* dn = n_corr ? ws_take<int32_t>(ctx, P) : nullptr;
  pre3_pair_result* dres = ws_take<pre3_pair_result>(ctx, P);
  uint8_t* dm = masks ? ws_take<uint8_t>(ctx, mb) : nullptr;
  int32_t* dc = counts ? ws_take<int32_t>(ctx, (size_t)P * H) : nullptr;
  pre3_dr_ye_stat* dst = stat ? ws_take<pre3_dr_ye_stat>(ctx, P) : nullptr;
  PRE3_TRY(h2d(ctx, dYa, Ya, pb));
  PRE3_TRY(h2d(ctx, dYb, Yb, pb));
  if (ds) PRE3_TRY(h2d(ctx, ds, samples, sb));
  if (dmt) PRE3_TRY(h2d(ctx, dmt, match, mtb));
  if (dn) PRE3_TRY(h2d(ctx, dn, n_corr, 4 * (size_t)P));
  PRE3_TRY(ransac_impl(ctx, dYa, dYb, dn, P, Nmax, o, ds, 0, dres, dm, dc, nullptr, dmt, dst));
  PRE3_TRY(d2h(ctx, res, dres, rb));
  if (masks) PRE3_TRY(d2h(ctx, masks, dm, mb));
  if (counts) PRE3_TRY(d2h(ctx, counts, dc, cb));
  if (stat) PRE3_TRY(d2h(ctx, stat, dst, stb));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

// ================================================================================================
// whole frame pairs
// ================================================================================================
static size_t pairs_ws_bytes(int cls, int P, int K1, int K2, int ND, int H) {
  return match_ws_bytes(cls, P, K1, K2, ND) + ransac_ws_bytes(P, K1, H) + 2 * align_up(24 * (size_t)P * K1) +
         align_up(4 * (size_t)P) + align_up(8 * (size_t)P * K1) + 8192;
}

static int pairs_impl(pre3_ctx* ctx, const void* ddesc1, const void* ddesc2, int cls, const double* dxyz1,
                      const double* dxyz2, int P, int K1, int K2, int ND, const int32_t* dk1, const int32_t* dk2,
                      const pre3_ransac_opts& o, uint32_t pair_id0, pre3_pair_result* dres, int32_t* dmatches,
                      uint8_t* dmasks) {
  MatchRow* rows = nullptr;
  if (K1 > 0) PRE3_TRY(match_impl(ctx, ddesc1, ddesc2, cls, P, K1, K2, ND, dk1, dk2, o.ratio, 0, &rows));
  if (!ddesc2) {  // sequence mode: frame p + 1 is the "current" frame of pair p
    dxyz2 = dxyz1 + 3 * (size_t)K1;
    if (dk1) dk2 = dk1 + 1;
  }
  double* dYa = ws_take<double>(ctx, 3 * (size_t)P * K1);
  double* dYb = ws_take<double>(ctx, 3 * (size_t)P * K1);
  int32_t* dn = ws_take<int32_t>(ctx, P);
  int32_t* dpairs = dmatches ? dmatches : ws_take<int32_t>(ctx, 2 * (size_t)P * K1);
  PRE3_TRY(launch_match_compact(ctx, rows, P, K1, dk1, dpairs, nullptr, dn, dxyz1, dxyz2, K2, dYa, dYb));
  return ransac_impl(ctx, dYa, dYb, dn, P, K1, o, nullptr, pair_id0, dres, dmasks, nullptr, nullptr, dpairs);
}

int pre3_set_graphs(pre3_ctx* ctx, int on) {
  if (!ctx || ctx->device < 0) return PRE3_ERR_CUDA;
  ctx->graphs = on != 0;
  if (!ctx->graphs && ctx->graph_exec) {
    cudaStreamSynchronize(ctx->stream);
    cudaGraphExecDestroy(ctx->graph_exec);
    ctx->graph_exec = nullptr;
    ctx->graph_key.clear();
    ctx->graph_seen.clear();
  }
  return PRE3_OK;
}

int pre3_set_pipeline(pre3_ctx* ctx, int chunks) {
  if (!ctx || ctx->device < 0) return PRE3_ERR_CUDA;
  if (chunks < -1 || chunks > 64) return fail(ctx, PRE3_ERR_ARG, "pipeline chunks: -1 (automatic), 0 / 1 (off) or 2..64");
  if (chunks != ctx->pipe_chunks && ctx->graph_exec) {  // a captured graph has the old chunking baked in
    cudaStreamSynchronize(ctx->stream);
    cudaGraphExecDestroy(ctx->graph_exec);
    ctx->graph_exec = nullptr;
    ctx->graph_key.clear();
    ctx->graph_seen.clear();
  }
  ctx->pipe_chunks = chunks;
  return PRE3_OK;
}

int pre3_pairs_dev(pre3_ctx* ctx, const void* ddesc1, const void* ddesc2, int cls, const double* dxyz1,
                   const double* dxyz2, int P, int K1, int K2, int ND, const int32_t* dk1_count,
                   const int32_t* dk2_count, const pre3_ransac_opts* opts, uint32_t pair_id0, pre3_pair_result* dres,
                   int32_t* dmatches, uint8_t* dmasks) {
  PRE3_LIVE();
  PRE3_TRY(check_match_args(ctx, ddesc1, ddesc2, cls, P, K1, K2, ND));
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(dres && dxyz1 && dxyz2, "null pointer");
  if (P == 0) return PRE3_OK;
  std::vector<unsigned char> key;
  if (ctx->graphs) {
    key_put(key, 1);
    key_put(key, ddesc1); key_put(key, ddesc2); key_put(key, cls); key_put(key, dxyz1); key_put(key, dxyz2);
    key_put(key, P); key_put(key, K1); key_put(key, K2); key_put(key, ND); key_put(key, dk1_count); key_put(key, dk2_count);
    key_put(key, *opts); key_put(key, pair_id0); key_put(key, dres); key_put(key, dmatches); key_put(key, dmasks);
    key_put(key, ctx->stream); key_put(key, ctx->match_engine);
  }
  const int nch = pipe_chunk_count(ctx, P, K1, K2);
  if (ctx->graphs) key_put(key, nch);
  return graph_or_eager(ctx, key, [&]() -> int {
    const int cmax = (P + nch - 1) / nch;
    PRE3_TRY(ws_reserve(ctx, (size_t)nch * pairs_ws_bytes(cls, cmax, K1, K2, ND, opts->H)));
    const size_t es = class_size(cls);
    return run_pipelined(ctx, P, nch, [&](int p0, int n) -> int {
      return pairs_impl(ctx, (const char*)ddesc1 + (size_t)p0 * K1 * ND * es, (const char*)ddesc2 + (size_t)p0 * K2 * ND * es,
                        cls, dxyz1 + 3 * (size_t)p0 * K1, dxyz2 + 3 * (size_t)p0 * K2, n, K1, K2, ND,
                        dk1_count ? dk1_count + p0 : nullptr, dk2_count ? dk2_count + p0 : nullptr, *opts,
                        pair_id0 + (uint32_t)p0, dres + p0, dmatches ? dmatches + 2 * (size_t)p0 * K1 : nullptr,
                        dmasks ? dmasks + (size_t)p0 * K1 : nullptr);
    });
  });
}

// A SEQUENCE of F frames = F - 1 consecutive pairs (frame p, frame p + 1): what the reference's whole-sequence
// loops run (find_consistent_sift_matches.m:22-32: RANSAC_CALC_SAVE_SR4000(i, i+1) for every i).  Every frame's
// descriptors are converted once instead of twice.
int pre3_sequence_dev(pre3_ctx* ctx, const void* ddesc, int cls, const double* dxyz, int F, int K, int ND,
                      const int32_t* dk_count, const pre3_ransac_opts* opts, uint32_t pair_id0, pre3_pair_result* dres,
                      int32_t* dmatches, uint8_t* dmasks) {
  PRE3_LIVE();
  PRE3_TRY(check_match_args(ctx, ddesc, ddesc, cls, F, K, K, ND));
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(F <= 1 || (dres && dxyz), "null pointer");
  if (F <= 1) return PRE3_OK;
  std::vector<unsigned char> key;
  if (ctx->graphs) {
    key_put(key, 2);
    key_put(key, ddesc); key_put(key, cls); key_put(key, dxyz); key_put(key, F); key_put(key, K); key_put(key, ND);
    key_put(key, dk_count); key_put(key, *opts); key_put(key, pair_id0); key_put(key, dres); key_put(key, dmatches);
    key_put(key, dmasks); key_put(key, ctx->stream); key_put(key, ctx->match_engine);
  }
  const int nch = pipe_chunk_count(ctx, F - 1, K, K);
  if (ctx->graphs) key_put(key, nch);
  return graph_or_eager(ctx, key, [&]() -> int {
    const int cmax = (F - 1 + nch - 1) / nch;
    PRE3_TRY(ws_reserve(ctx, (size_t)nch * pairs_ws_bytes(cls, cmax + 1, K, K, ND, opts->H)));
    const size_t es = class_size(cls);
    return run_pipelined(ctx, F - 1, nch, [&](int p0, int n) -> int {  // pairs [p0, p0 + n) = frames p0 .. p0 + n
      return pairs_impl(ctx, (const char*)ddesc + (size_t)p0 * K * ND * es, nullptr, cls, dxyz + 3 * (size_t)p0 * K, nullptr,
                        n, K, K, ND, dk_count ? dk_count + p0 : nullptr, nullptr, *opts, pair_id0 + (uint32_t)p0,
                        dres + p0, dmatches ? dmatches + 2 * (size_t)p0 * K : nullptr,
                        dmasks ? dmasks + (size_t)p0 * K : nullptr);
    });
  });
}

// Host buffers: the P pairs are cut into chunks that are staged on a second stream while the
// previous chunk computes (double-buffered device staging), so that for whole-sequence runs the
// PCIe copy of the descriptors overlaps the kernels.  Class-double descriptors (1 MB per 512x512
// pair) are narrowed to float into pinned staging by the host pool when every value survives the
// round trip (hostconv.cpp): half the PCIe bytes, same arithmetic on the device.
static int ensure_host_stage(pre3_ctx* ctx, size_t bytes_per_slot) {
  if (ctx->host_f32 < 0) {
    // Narrowing trades host memory traffic (read 8 B + write 4 B + DMA 4 B per value instead of DMA 8 B) for PCIe
    // bytes: it pays while ONE link is the bottleneck (measured on a 16-core host: 50 k -> 63 k pairs/s).  With
    // several ranks on one host (LOCAL_WORLD_SIZE, set by torchrun) every GPU has its own link and the shared host
    // memory becomes the limit (measured at 2 ranks: 89 k pairs/s narrowed vs 2 x 50 k raw), so it is switched off.
    // PRE3_HOST_F32=0/1 overrides.
    const char* e = getenv("PRE3_HOST_F32");
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    const unsigned ranks = lws && atoi(lws) > 0 ? (unsigned)atoi(lws) : 1u;
    const unsigned hw = std::thread::hardware_concurrency();
    ctx->host_f32 = e ? (atoi(e) != 0) : (ranks == 1 && hw >= 8);
    if (ctx->host_f32) ctx->pool = host_pool_create((int)std::min<unsigned>(hw > 1 ? hw - 1 : 1, 15));
  }
  if (!ctx->host_f32) return PRE3_OK;
  if (ctx->h_stage_cap < bytes_per_slot) {
    for (int i = 0; i < 2; ++i) {
      if (ctx->h_stage[i]) {
        cudaStreamSynchronize(ctx->copy_stream ? ctx->copy_stream : ctx->stream);
        cudaFreeHost(ctx->h_stage[i]);
        ctx->h_stage[i] = nullptr;
      }
      PRE3_CUDA(cudaHostAlloc((void**)&ctx->h_stage[i], bytes_per_slot, cudaHostAllocDefault));
      if (!ctx->ev_stage_done[i]) PRE3_CUDA(cudaEventCreateWithFlags(&ctx->ev_stage_done[i], cudaEventDisableTiming));
    }
    ctx->h_stage_cap = bytes_per_slot;
  }
  return PRE3_OK;
}

// seq: desc1 / xyz1 / k1_count hold P + 1 frames, pair p = (frame p, frame p + 1); desc2 / xyz2 / k2_count unused.
// A chunk of n pairs then stages n + 1 frames (the boundary frame is sent twice, once per chunk).
static int pairs_host(pre3_ctx* ctx, const void* desc1, const void* desc2, int cls, const double* xyz1,
                      const double* xyz2, int P, int K1, int K2, int ND, const int32_t* k1_count,
                      const int32_t* k2_count, const pre3_ransac_opts* opts, uint32_t pair_id0, pre3_pair_result* res,
                      int32_t* matches, uint8_t* masks, bool seq) {
  PRE3_LIVE();
  PRE3_TRY(check_match_args(ctx, desc1, seq ? desc1 : desc2, cls, P, K1, K2, ND));
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(res && xyz1 && (seq || xyz2), "null pointer");
  if (P == 0) return PRE3_OK;
  const int X = seq ? 1 : 0;  // extra frame per chunk on the "1" side
  const size_t es = class_size(cls);
  const size_t pb1 = (size_t)K1 * ND * es, pb2 = (size_t)K2 * ND * es;  // per pair
  const size_t per_pair_in = pb1 + pb2 + 24 * ((size_t)K1 + K2) + 8;
  // chunk size: ~64 MB of input per chunk, at least 1 pair, at most P
  int C = (int)std::max<size_t>(1, std::min<size_t>((size_t)P, ((size_t)64 << 20) / std::max<size_t>(per_pair_in, 1)));
  const int nchunks = (P + C - 1) / C;
  // staging (two slots) lives outside the per-call arena: the arena is reset by every chunk
  struct Slot {
    char *d1, *d2;
    double *x1, *x2;
    int32_t *k1, *k2;
    pre3_pair_result* r;
    int32_t* m;
    uint8_t* msk;
  } slot[2];
  const size_t slot_bytes = align_up(pb1 * (C + 1)) + align_up(pb2 * C) + align_up(24 * (size_t)K1 * (C + 1)) +
                            align_up(24 * (size_t)K2 * C) + 2 * align_up(4 * (size_t)(C + 1)) +
                            align_up(sizeof(pre3_pair_result) * (size_t)C) + align_up(8 * (size_t)K1 * C) +
                            align_up((size_t)K1 * C) + 4096;
  PRE3_TRY(aux_reserve(ctx, 0, 2 * slot_bytes));
  for (int s = 0; s < 2; ++s) {
    char* base = ctx->aux[0] + s * slot_bytes;
    size_t off = 0;
    auto take = [&](size_t bytes) {
      char* p = base + off;
      off += align_up(bytes);
      return p;
    };
    slot[s].d1 = take(pb1 * (C + 1));
    slot[s].d2 = take(pb2 * C);
    slot[s].x1 = (double*)take(24 * (size_t)K1 * (C + 1));
    slot[s].x2 = (double*)take(24 * (size_t)K2 * C);
    slot[s].k1 = (int32_t*)take(4 * (size_t)(C + 1));
    slot[s].k2 = (int32_t*)take(4 * (size_t)(C + 1));
    slot[s].r = (pre3_pair_result*)take(sizeof(pre3_pair_result) * (size_t)C);
    slot[s].m = (int32_t*)take(8 * (size_t)K1 * C);
    slot[s].msk = (uint8_t*)take((size_t)K1 * C);
  }
  PRE3_TRY(ensure_copy_stream(ctx));
  PRE3_TRY(ws_reserve(ctx, pairs_ws_bytes(cls, C + 1, K1, K2, ND, opts->H)));
  const size_t n1 = (size_t)K1 * ND, n2 = (size_t)K2 * ND;  // descriptor values per pair
  bool narrow = cls == PRE3_CLASS_DOUBLE && ND > 0;
  if (narrow) {
    PRE3_TRY(ensure_host_stage(ctx, 4 * (n1 + n2) * (size_t)(C + 1)));
    narrow = ctx->host_f32 == 1;
  }
  // results come back through pinned staging: a device->host copy into pageable user memory would block
  // the host until the chunk's kernels have finished and serialise the whole pipeline
  const size_t rb_res = sizeof(pre3_pair_result) * (size_t)P, rb_m = matches ? 8 * (size_t)K1 * P : 0,
               rb_k = masks ? (size_t)K1 * P : 0;
  if (ctx->h_pin_cap < rb_res + rb_m + rb_k) {
    if (ctx->h_pin) {
      cudaStreamSynchronize(ctx->stream);
      cudaFreeHost(ctx->h_pin);
      ctx->h_pin = nullptr;
      ctx->h_pin_cap = 0;
    }
    PRE3_CUDA(cudaHostAlloc(&ctx->h_pin, rb_res + rb_m + rb_k, cudaHostAllocDefault));
    ctx->h_pin_cap = rb_res + rb_m + rb_k;
  }
  char* hp_res = (char*)ctx->h_pin;
  char* hp_m = hp_res + rb_res;
  char* hp_k = hp_m + rb_m;
  int chunk_cls[2] = {cls, cls};
  const int raw_every = getenv("PRE3_HOST_F32_RAW_EVERY") ? atoi(getenv("PRE3_HOST_F32_RAW_EVERY")) : 5;
  cudaStream_t cs = ctx->copy_stream;
  double t_conv = 0.0, t_wait = 0.0, t_enq = 0.0, t_stage = 0.0;  // PRE3_DEBUG diagnostics
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  auto stage_in = [&](int c) -> int {
    const int s = c & 1;
    const int p0 = c * C, n = std::min(C, P - p0);
    bool as_f32 = false;
    // every raw_every-th chunk crosses as doubles: the host pool (~75 GB/s of doubles on a 16-core host)
    // and the PCIe link (~54 GB/s) then finish together instead of the link waiting for the pool
    if (narrow && !(raw_every > 0 && c % raw_every == raw_every - 1)) {
      // the pinned slot is free once the H2D of chunk c-2 has finished
      const double ta = now();
      if (c >= 2) PRE3_CUDA(cudaEventSynchronize(ctx->ev_stage_done[s]));
      const double tb = now();
      float* f1 = (float*)ctx->h_stage[s];
      float* f2 = f1 + n1 * (size_t)(C + 1);
      as_f32 = host_narrow(ctx->pool, (const double*)desc1 + (size_t)p0 * n1, f1, n1 * (n + X)) &&
               (seq || host_narrow(ctx->pool, (const double*)desc2 + (size_t)p0 * n2, f2, n2 * n));
      t_wait += tb - ta;
      t_conv += now() - tb;
    }
    chunk_cls[s] = as_f32 ? PRE3_CLASS_DOUBLE_F32 : cls;
    ctx->h2d_bytes += (int64_t)((as_f32 ? 4 * n1 : pb1) * (n + X) + (seq ? 0 : (as_f32 ? 4 * n2 : pb2) * n) +
                                24 * (size_t)K1 * (n + X) + (seq ? 0 : 24 * (size_t)K2 * n) +
                                (k1_count ? 4 * (size_t)(n + X) : 0) + (!seq && k2_count ? 4 * (size_t)n : 0));
    // the device slot may still be read by the compute of chunk c-2 / written back by its D2H
    PRE3_CUDA(cudaStreamWaitEvent(cs, ctx->ev_slot_free[s], 0));
    if (as_f32) {
      const float* f1 = (const float*)ctx->h_stage[s];
      PRE3_CUDA(cudaMemcpyAsync(slot[s].d1, f1, 4 * n1 * (n + X), cudaMemcpyHostToDevice, cs));
      if (!seq)
        PRE3_CUDA(cudaMemcpyAsync(slot[s].d2, f1 + n1 * (size_t)(C + 1), 4 * n2 * n, cudaMemcpyHostToDevice, cs));
      PRE3_CUDA(cudaEventRecord(ctx->ev_stage_done[s], cs));
    } else {
      PRE3_CUDA(cudaMemcpyAsync(slot[s].d1, (const char*)desc1 + (size_t)p0 * pb1, pb1 * (n + X), cudaMemcpyHostToDevice, cs));
      if (!seq)
        PRE3_CUDA(cudaMemcpyAsync(slot[s].d2, (const char*)desc2 + (size_t)p0 * pb2, pb2 * n, cudaMemcpyHostToDevice, cs));
      if (narrow) PRE3_CUDA(cudaEventRecord(ctx->ev_stage_done[s], cs));  // keeps the slot's event current
    }
    PRE3_CUDA(cudaMemcpyAsync(slot[s].x1, xyz1 + 3 * (size_t)p0 * K1, 24 * (size_t)K1 * (n + X), cudaMemcpyHostToDevice, cs));
    if (!seq)
      PRE3_CUDA(cudaMemcpyAsync(slot[s].x2, xyz2 + 3 * (size_t)p0 * K2, 24 * (size_t)K2 * n, cudaMemcpyHostToDevice, cs));
    if (k1_count) PRE3_CUDA(cudaMemcpyAsync(slot[s].k1, k1_count + p0, 4 * (size_t)(n + X), cudaMemcpyHostToDevice, cs));
    if (!seq && k2_count) PRE3_CUDA(cudaMemcpyAsync(slot[s].k2, k2_count + p0, 4 * (size_t)n, cudaMemcpyHostToDevice, cs));
    PRE3_CUDA(cudaEventRecord(ctx->ev_slot_full[s], cs));
    return PRE3_OK;
  };
  // both slots start free
  PRE3_CUDA(cudaEventRecord(ctx->ev_slot_free[0], ctx->stream));
  PRE3_CUDA(cudaEventRecord(ctx->ev_slot_free[1], ctx->stream));
  PRE3_TRY(stage_in(0));
  for (int c = 0; c < nchunks; ++c) {
    const int s = c & 1;
    const int p0 = c * C, n = std::min(C, P - p0);
    // enqueue the compute of chunk c first: the host then narrows chunk c+1 while the copy engine moves
    // chunk c (already enqueued) and the SMs work on it
    const double tq0 = now();
    PRE3_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_slot_full[s], 0));
    ctx->ws_off = 0;  // arena reused by every chunk (stream order keeps it safe)
    PRE3_TRY(pairs_impl(ctx, slot[s].d1, seq ? nullptr : slot[s].d2, chunk_cls[s], slot[s].x1, seq ? nullptr : slot[s].x2,
                        n, K1, K2, ND, k1_count ? slot[s].k1 : nullptr, !seq && k2_count ? slot[s].k2 : nullptr, *opts,
                        pair_id0 + (uint32_t)p0, slot[s].r, matches ? slot[s].m : nullptr,
                        masks ? slot[s].msk : nullptr));
    PRE3_TRY(d2h(ctx, hp_res + sizeof(pre3_pair_result) * (size_t)p0, slot[s].r, sizeof(pre3_pair_result) * (size_t)n));
    if (matches) PRE3_TRY(d2h(ctx, hp_m + 8 * (size_t)p0 * K1, slot[s].m, 8 * (size_t)K1 * n));
    if (masks) PRE3_TRY(d2h(ctx, hp_k + (size_t)p0 * K1, slot[s].msk, (size_t)K1 * n));
    ctx->d2h_bytes += (int64_t)(sizeof(pre3_pair_result) * (size_t)n + (matches ? 8 * (size_t)K1 * n : 0) +
                                (masks ? (size_t)K1 * n : 0));
    PRE3_CUDA(cudaEventRecord(ctx->ev_slot_free[s], ctx->stream));
    const double tq1 = now();
    t_enq += tq1 - tq0;
    if (c + 1 < nchunks) PRE3_TRY(stage_in(c + 1));
    t_stage += now() - tq1;
  }
  const double t_sync0 = now();
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  PRE3_CUDA(cudaStreamSynchronize(cs));
  memcpy(res, hp_res, rb_res);
  if (matches) memcpy(matches, hp_m, rb_m);
  if (masks) memcpy(masks, hp_k, rb_k);
  if (getenv("PRE3_DEBUG"))
    fprintf(stderr, "[pre3] pairs: P=%d chunks=%d narrow=%d host narrowing %.1f ms, staging waits %.1f ms, enqueue compute %.1f ms, stage_in %.1f ms, final sync %.1f ms\n",
            P, nchunks, (int)narrow, t_conv * 1e3, t_wait * 1e3, t_enq * 1e3, t_stage * 1e3, (now() - t_sync0) * 1e3);
  return PRE3_OK;
}

int pre3_pairs(pre3_ctx* ctx, const void* desc1, const void* desc2, int cls, const double* xyz1, const double* xyz2,
               int P, int K1, int K2, int ND, const int32_t* k1_count, const int32_t* k2_count,
               const pre3_ransac_opts* opts, uint32_t pair_id0, pre3_pair_result* res, int32_t* matches,
               uint8_t* masks) {
  if (ctx && !(desc2 || (size_t)P * K2 * ND == 0)) return fail(ctx, PRE3_ERR_ARG, "descriptor pointer missing");
  return pairs_host(ctx, desc1, desc2, cls, xyz1, xyz2, P, K1, K2, ND, k1_count, k2_count, opts, pair_id0, res, matches,
                    masks, false);
}

int pre3_sequence(pre3_ctx* ctx, const void* desc, int cls, const double* xyz, int F, int K, int ND,
                  const int32_t* k_count, const pre3_ransac_opts* opts, uint32_t pair_id0, pre3_pair_result* res,
                  int32_t* matches, uint8_t* masks) {
  if (F <= 1) return ctx ? PRE3_OK : PRE3_ERR_ARG;
  return pairs_host(ctx, desc, nullptr, cls, xyz, nullptr, F - 1, K, K, ND, k_count, nullptr, opts, pair_id0, res,
                    matches, masks, true);
}

// ================================================================================================
// covariance of the RANSAC pose (cov_est_RANSAC_deriv.m)
// ================================================================================================
int pre3_cov_est_ransac_batch_dev(pre3_ctx* ctx, const double* dYa, const double* dYb, const int32_t* dn_corr,
                                  const uint8_t* dmasks, int P, int Nmax, const double* dRT, int rt_stride,
                                  pre3_cov_result* dout) {
  PRE3_LIVE();
  PRE3_NEED(P >= 0 && Nmax >= 0 && rt_stride >= 12, "bad sizes");
  if (P == 0) return PRE3_OK;
  PRE3_NEED(dYa && dYb && dRT && dout, "null pointer");
  PRE3_TRY(ws_reserve(ctx, cov_workspace_bytes(P, Nmax)));
  return launch_cov_est(ctx, dYa, dYb, dn_corr, dmasks, P, Nmax, dRT, rt_stride, dout);
}

int pre3_cov_est_ransac_batch(pre3_ctx* ctx, const double* Ya, const double* Yb, const int32_t* n_corr,
                              const uint8_t* masks, int P, int Nmax, const double* R, const double* T,
                              pre3_cov_result* out) {
  PRE3_LIVE();
  PRE3_NEED(P >= 0 && Nmax >= 0, "bad sizes");
  if (P == 0) return PRE3_OK;
  PRE3_NEED(Ya && Yb && R && T && out, "null pointer");
  const size_t pts = (size_t)P * Nmax * 3 * sizeof(double);
  PRE3_TRY(ws_reserve(ctx, cov_workspace_bytes(P, Nmax) + 2 * align_up(pts) + align_up((size_t)P * Nmax) +
                               align_up(4 * (size_t)P) + align_up(96 * (size_t)P) +
                               align_up(sizeof(pre3_cov_result) * (size_t)P) + 4096));
  double* dYa = ws_take<double>(ctx, (size_t)P * Nmax * 3);
  double* dYb = ws_take<double>(ctx, (size_t)P * Nmax * 3);
  uint8_t* dm = masks ? ws_take<uint8_t>(ctx, (size_t)P * Nmax) : nullptr;
  int32_t* dn = n_corr ? ws_take<int32_t>(ctx, P) : nullptr;
  double* dRT = ws_take<double>(ctx, (size_t)P * 12);
  pre3_cov_result* dout = ws_take<pre3_cov_result>(ctx, P);
  std::vector<double> rt((size_t)P * 12);
  for (int p = 0; p < P; ++p) {
    for (int i = 0; i < 9; ++i) rt[(size_t)p * 12 + i] = R[(size_t)p * 9 + i];
    for (int i = 0; i < 3; ++i) rt[(size_t)p * 12 + 9 + i] = T[(size_t)p * 3 + i];
  }
  PRE3_TRY(h2d(ctx, dYa, Ya, pts));
  PRE3_TRY(h2d(ctx, dYb, Yb, pts));
  if (dm) PRE3_TRY(h2d(ctx, dm, masks, (size_t)P * Nmax));
  if (dn) PRE3_TRY(h2d(ctx, dn, n_corr, 4 * (size_t)P));
  PRE3_TRY(h2d(ctx, dRT, rt.data(), 96 * (size_t)P));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));  // rt is a local
  PRE3_TRY(launch_cov_est(ctx, dYa, dYb, dn, dm, P, Nmax, dRT, 12, dout));
  PRE3_TRY(d2h(ctx, out, dout, sizeof(pre3_cov_result) * (size_t)P));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

// ================================================================================================
// hypothesis-block sharding
// ================================================================================================
int pre3_distance_threshold_dev(pre3_ctx* ctx, const double* dYb, int N, double* dthr) {
  PRE3_LIVE();
  PRE3_NEED(dYb && dthr, "null pointer");
  return launch_threshold(ctx, dYb, N, dthr);
}

int pre3_ransac_block_dev(pre3_ctx* ctx, const double* dYa, const double* dYb, int N, const pre3_ransac_opts* opts,
                          const int32_t* dsamples, int64_t h0, int Hloc, double thr, uint64_t* dkey,
                          double* derrsum) {
  PRE3_LIVE();
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(dYa && dYb && dkey, "null pointer");
  PRE3_NEED(N >= 1 && Hloc >= 0 && h0 >= 0, "bad sizes");
  pre3_ransac_opts o = *opts;
  o.H = Hloc;
  o.adaptive = 0;  // fixed-H by construction
  o.distance_threshold = thr;
  PRE3_TRY(ws_reserve(ctx, ransac_ws_bytes(1, N, Hloc)));
  RansacBuffers b{};
  b.Ya = dYa;
  b.Yb = dYb;
  b.n_corr = nullptr;
  b.P = 1;
  b.Nmax = N;
  b.samples = dsamples;
  b.pair_id0 = 0;
  ransac_carve(ctx, b, Hloc);
  PRE3_TRY(launch_prep(ctx, b, o, 1));
  PRE3_TRY(launch_eval(ctx, b, o, h0, 0, Hloc, nullptr));
  PRE3_TRY(launch_block_best(ctx, b, o, h0, Hloc, dkey, derrsum));
  return PRE3_OK;
}

// Reference-exact local selection over one hypothesis block: evaluate [h0, h0+Hloc), then the full
// (max cardinality, min ErrorSum, first index) rule + mask + refit, as if the block were the whole run.
int pre3_ransac_block_select_dev(pre3_ctx* ctx, const double* dYa, const double* dYb, int N,
                                 const pre3_ransac_opts* opts, const int32_t* dsamples, int64_t h0, int Hloc,
                                 double thr, pre3_pair_result* dres, uint8_t* dmask) {
  PRE3_LIVE();
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(dYa && dYb && dres, "null pointer");
  PRE3_NEED(N >= 1 && Hloc >= 0 && h0 >= 0, "bad sizes");
  pre3_ransac_opts o = *opts;
  o.H = Hloc;
  o.adaptive = 0;
  o.max_iteration = Hloc + 1;  // the loop runs over every sample set of the block
  o.distance_threshold = thr;
  PRE3_TRY(ws_reserve(ctx, ransac_ws_bytes(1, N, Hloc)));
  RansacBuffers b{};
  b.Ya = dYa;
  b.Yb = dYb;
  b.n_corr = nullptr;
  b.P = 1;
  b.Nmax = N;
  b.samples = dsamples;
  b.pair_id0 = 0;
  b.h0 = h0;
  ransac_carve(ctx, b, Hloc);
  PRE3_TRY(launch_prep(ctx, b, o, 1));
  PRE3_TRY(launch_eval(ctx, b, o, h0, 0, Hloc, nullptr));
  PRE3_TRY(launch_select(ctx, b, o, dres, dmask, nullptr, nullptr));
  return PRE3_OK;
}

int pre3_ransac_finish_dev(pre3_ctx* ctx, const double* dYa, const double* dYb, int N, const pre3_ransac_opts* opts,
                           const int32_t* dsamples_of_winner, int64_t winner_id, double thr, pre3_pair_result* dres,
                           uint8_t* dmask) {
  PRE3_LIVE();
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(dYa && dYb && dres, "null pointer");
  PRE3_NEED(N >= 1 && winner_id >= 0, "bad sizes");
  pre3_ransac_opts o = *opts;
  o.distance_threshold = thr;
  PRE3_TRY(ws_reserve(ctx, ransac_ws_bytes(1, N, 1)));
  RansacBuffers b{};
  b.Ya = dYa;
  b.Yb = dYb;
  b.n_corr = nullptr;
  b.P = 1;
  b.Nmax = N;
  b.samples = dsamples_of_winner;
  b.pair_id0 = 0;
  ransac_carve(ctx, b, 1);
  uint8_t* scratch = ws_take<uint8_t>(ctx, (size_t)N);
  PRE3_TRY(launch_prep(ctx, b, o, 1));
  PRE3_TRY(launch_finish(ctx, b, o, winner_id, dres, dmask ? dmask : scratch));
  return PRE3_OK;
}

// ---- the hypothesis-block split, stream-ordered (no host read between the kernels and the collective) ----------
// mode 0 ("first": max count, lowest id): dkey[0] = this block's key; the caller all-reduces it (MAX) on the same
// stream and calls pre3_ransac_split_finish_dev.  mode 1 (reference rule): the block's own full selection (record +
// mask) and dkey[0..1] = (key, ErrorSum bits); the caller all-gathers the 16 bytes and calls split_finish.
// The threshold (RANSAC_CALC_VER2.m:69-72 for the SVD method) is computed on the device.
static int split_setup(pre3_ctx* ctx, const double* dYa, const double* dYb, int N, pre3_ransac_opts& o,
                       const int32_t* dsamples, int Hloc, RansacBuffers& b) {
  PRE3_TRY(ws_reserve(ctx, ransac_ws_bytes(1, N, Hloc) + 1024));
  b.Ya = dYa;
  b.Yb = dYb;
  b.n_corr = nullptr;
  b.P = 1;
  b.Nmax = N;
  b.samples = dsamples;
  b.pair_id0 = 0;
  ransac_carve(ctx, b, std::max(Hloc, 1));
  double* dthr = nullptr;
  if (o.method == PRE3_METHOD_SVD) {
    dthr = ws_take<double>(ctx, 1);
    PRE3_TRY(launch_threshold(ctx, dYb, N, dthr));
  }
  o.adaptive = 0;  // fixed-H by construction
  PRE3_TRY(launch_prep(ctx, b, o, 1, dthr));
  return PRE3_OK;
}

int pre3_ransac_split_local_dev(pre3_ctx* ctx, const double* dYa, const double* dYb, int N, const pre3_ransac_opts* opts,
                                const int32_t* dsamples, int64_t h0, int Hloc, int mode, uint64_t* dkey,
                                pre3_pair_result* dres, uint8_t* dmask) {
  PRE3_LIVE();
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(dYa && dYb && dkey, "null pointer");
  PRE3_NEED(N >= 1 && Hloc >= 0 && h0 >= 0, "bad sizes");
  PRE3_NEED(mode == 0 || (mode == 1 && dres), "mode 1 needs the record buffer");
  pre3_ransac_opts o = *opts;
  o.H = Hloc;
  if (mode == 1) o.max_iteration = Hloc + 1;  // the loop runs over every sample set of the block
  RansacBuffers b{};
  PRE3_TRY(split_setup(ctx, dYa, dYb, N, o, dsamples, Hloc, b));
  b.h0 = h0;
  PRE3_TRY(launch_eval(ctx, b, o, h0, 0, Hloc, nullptr));
  if (mode == 0) return launch_block_best(ctx, b, o, h0, Hloc, dkey, nullptr);
  PRE3_TRY(launch_select(ctx, b, o, dres, dmask, nullptr, nullptr));
  return launch_split_pack(ctx, dres, h0, dkey);
}

int pre3_ransac_split_finish_dev(pre3_ctx* ctx, const double* dYa, const double* dYb, int N, const pre3_ransac_opts* opts,
                                 const int32_t* dsamples, int64_t h0, int Hloc, int mode, const uint64_t* dexchanged,
                                 int world, int rank, pre3_pair_result* dres, uint8_t* dmask) {
  PRE3_LIVE();
  PRE3_TRY(check_opts(ctx, opts));
  PRE3_NEED(dYa && dYb && dexchanged && dres, "null pointer");
  PRE3_NEED(N >= 1 && Hloc >= 0 && h0 >= 0 && world >= 1 && rank >= 0 && rank < world, "bad sizes");
  if (mode == 1) return launch_split_keep(ctx, dexchanged, world, rank, h0, dres, dmask, N);
  pre3_ransac_opts o = *opts;
  o.H = Hloc;
  RansacBuffers b{};
  PRE3_TRY(split_setup(ctx, dYa, dYb, N, o, dsamples, 1, b));
  uint8_t* scratch = dmask ? nullptr : ws_take<uint8_t>(ctx, (size_t)N);
  return launch_finish_key(ctx, b, o, dexchanged, h0, Hloc, dres, dmask ? dmask : scratch);
}

// R2q of slamToolbox (M/slamToolbox_11_02_18/FrameTransforms/Rotations/R2q.m:11-55): host helper.
void pre3_R2q(const double* Rc, double* q) {
  // column-major: R(i,j) = Rc[3*(j-1) + (i-1)]
  auto R = [&](int i, int j) { return Rc[3 * (j - 1) + (i - 1)]; };
  const double T = ((R(1, 1) + R(2, 2)) + R(3, 3)) + 1.0;
  double a, b, c, d;
  if (T > 0.00000001) {
    const double S = 2.0 * std::sqrt(T);
    a = 0.25 * S;
    b = (R(2, 3) - R(3, 2)) / S;
    c = (R(3, 1) - R(1, 3)) / S;
    d = (R(1, 2) - R(2, 1)) / S;
  } else if (R(1, 1) > R(2, 2) && R(1, 1) > R(3, 3)) {
    const double S = 2.0 * std::sqrt(1.0 + R(1, 1) - R(2, 2) - R(3, 3));
    a = (R(2, 3) - R(3, 2)) / S;
    b = 0.25 * S;
    c = (R(1, 2) + R(2, 1)) / S;
    d = (R(3, 1) + R(1, 3)) / S;
  } else if (R(2, 2) > R(3, 3)) {
    const double S = 2.0 * std::sqrt(1.0 + R(2, 2) - R(1, 1) - R(3, 3));
    a = (R(3, 1) - R(1, 3)) / S;
    b = (R(1, 2) + R(2, 1)) / S;
    c = 0.25 * S;
    d = (R(2, 3) + R(3, 2)) / S;
  } else {
    const double S = 2.0 * std::sqrt(1.0 + R(3, 3) - R(1, 1) - R(2, 2));
    a = (R(1, 2) - R(2, 1)) / S;
    b = (R(3, 1) + R(1, 3)) / S;
    c = (R(2, 3) + R(3, 2)) / S;
    d = 0.25 * S;
  }
  q[0] = a;
  q[1] = -b;
  q[2] = -c;
  q[3] = -d;
}

}  // extern "C"
