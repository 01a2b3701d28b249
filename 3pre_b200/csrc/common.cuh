// common.cuh -- context, workspace arena and launch helpers shared by the libpre3 TUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/pre3.h"

// Internal storage class of the host-pointer batch entry points: values of class double that were
// narrowed to float for the PCIe transfer (every value survives the round trip, hostconv.cpp); the
// arithmetic stays that of PRE3_CLASS_DOUBLE (double accumulation).
#define PRE3_CLASS_DOUBLE_F32 100

namespace pre3 {
class HostPool;
HostPool* host_pool_create(int threads);
void host_pool_destroy(HostPool* p);
int host_pool_size(const HostPool* p);
bool host_narrow(HostPool* pool, const double* s, float* d, size_t n);
// per-kernel timing categories (pre3_timing_*): bench.py's roofline numbers come from here
enum TimeCat { T_CONVERT = 0, T_MATCH_TC, T_MATCH_EXACT, T_RESCORE, T_COMPACT, T_PREP, T_EVAL, T_SELECT, T_OTHER, T_EKF_GAIN, T_EKF_SCORE, T_EKF_SELECT, T_FRAMES, T_EKF_UPDATE, T_MATCH_FUSED, T_NCAT };
struct TimedSpan {
  int cat;
  cudaEvent_t a, b;
};
}  // namespace pre3

struct pre3_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int match_engine = PRE3_MATCH_AUTO;
  int l1_shared = 0;  // set by the sweep entry points: ONE L1 descriptor set serves every problem of the batch
  int sm_count = 148;
  int64_t launches = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;  // moved by the host-pointer whole-pair entry point (bench.py's e2e)
  std::string err;
  // growable device workspace (bump allocator, reset per API call)
  char* ws = nullptr;
  size_t ws_cap = 0;
  size_t ws_off = 0;
  std::vector<char*> retired;  // blocks replaced while a call was being assembled
  // cached adaptive-iteration tables, keyed by (k, mult, Nmax, max_iteration)
  int tab_k = -1, tab_mult = -1, tab_nmax = -1, tab_maxit = -1;
  int32_t* d_tab = nullptr;       // triangular: row N starts at N*(N+1)/2, entries c = 0..N (Nmax <= 2048)
  int32_t* d_tab_rows = nullptr;  // per-call rows + row offsets for larger Nmax
  size_t tab_rows_cap = 0;
  // cached n_hyp table of the EKF path: (F+1) x (F+1) doubles, row num_ic, column support
  double* d_ekf_tab = nullptr;
  int ekf_tab_F = -1;
  // pinned staging for small results
  void* h_pin = nullptr;
  size_t h_pin_cap = 0;
  // long-lived device buffers outside the per-call arena (chunk staging of pre3_pairs,
  // pre-tiled fp16 operand images of the tensor-core matcher)
  std::vector<char*> aux;
  std::vector<size_t> aux_cap;
  // second stream + events for copy/compute overlap in the host-pointer batch entry points
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_slot_full[2] = {nullptr, nullptr};
  cudaEvent_t ev_slot_free[2] = {nullptr, nullptr};
  // host-side narrowing of double descriptors for the transfer (hostconv.cpp): pool, pinned float
  // staging (two slots), "H2D of this slot has finished" events
  pre3::HostPool* pool = nullptr;
  int host_f32 = -1;  // -1 undecided, 0 off, 1 on
  char* h_stage[2] = {nullptr, nullptr};
  size_t h_stage_cap = 0;
  cudaEvent_t ev_stage_done[2] = {nullptr, nullptr};
  // optional CUDA-graph replay of the device-resident whole-pair entry points (pre3_set_graphs): the launch sequence
  // of one (arguments, sizes, options) signature is captured once and replayed
  bool graphs = false;
  cudaGraphExec_t graph_exec = nullptr;
  std::vector<unsigned char> graph_key, graph_seen;  // signature of the captured graph / of the last eager call
  int64_t graph_launches = 0;                        // kernel launches one replay stands for
  // Software pipeline of the whole-pair entry points (pre3_set_pipeline): the pairs of a call are cut into chunks
  // and the stages of a chunk (convert | GEMM + rescore + compact | prep + eval | select) run on their own streams, so
  // that the HBM-bound, tensor-bound, FP32-issue-bound and latency-bound kernels of DIFFERENT chunks share the SMs.
  int pipe_chunks = -1;                  // -1: automatic, 0 / 1: off, n: chunks per call
  cudaStream_t pipe_stream[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t pipe_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t pipe_fork = nullptr;
  int pipe_stage = -1;                   // stage the launches are currently issued for (-1: no pipeline running)
  int pipe_map[4] = {0, 1, 2, 3};        // stage -> stream
  int pipe_total_P = 0;                  // pairs of the whole call while a pipeline runs (occupancy heuristics)
  // optional per-launch CUDA-event timing (off by default)
  bool timing = false;
  std::vector<pre3::TimedSpan> spans;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
};

namespace pre3 {

inline int fail(pre3_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define PRE3_CUDA(call)                                                                       \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return pre3::fail(ctx, PRE3_ERR_CUDA,                                                   \
                        std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                            ":" + std::to_string(__LINE__) + ")");                           \
  } while (0)

#define PRE3_TRY(call)         \
  do {                         \
    int rc__ = (call);         \
    if (rc__ != PRE3_OK) return rc__; \
  } while (0)

// Bump allocation out of the context arena.  All sub-allocations of one API call must be
// requested through ws_reserve() first (so that the block never moves mid-call).
inline int ws_reserve(pre3_ctx* ctx, size_t bytes) {
  ctx->ws_off = 0;
  if (!ctx->retired.empty()) {  // spill blocks of the previous call (ws_take)
    cudaStreamSynchronize(ctx->stream);
    for (char* r : ctx->retired) cudaFree(r);
    ctx->retired.clear();
  }
  if (bytes <= ctx->ws_cap) return PRE3_OK;
  size_t cap = ctx->ws_cap ? ctx->ws_cap : (size_t)1 << 20;
  while (cap < bytes) cap *= 2;
  if (ctx->ws) {
    // earlier work on the stream may still use the old block: free after a sync
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->ws);
    ctx->ws = nullptr;
    ctx->ws_cap = 0;
  }
  cudaError_t e = cudaMalloc((void**)&ctx->ws, cap);
  if (e != cudaSuccess) return fail(ctx, PRE3_ERR_ALLOC, std::string("cudaMalloc workspace: ") + cudaGetErrorString(e));
  ctx->ws_cap = cap;
  return PRE3_OK;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Long-lived device buffer number `slot` of at least `bytes` (grown on demand, never shrunk).
inline int aux_reserve(pre3_ctx* ctx, size_t slot, size_t bytes) {
  if (ctx->aux.size() <= slot) {
    ctx->aux.resize(slot + 1, nullptr);
    ctx->aux_cap.resize(slot + 1, 0);
  }
  if (ctx->aux_cap[slot] >= bytes) return PRE3_OK;
  if (ctx->aux[slot]) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    cudaFree(ctx->aux[slot]);
    ctx->aux[slot] = nullptr;
    ctx->aux_cap[slot] = 0;
  }
  cudaError_t e = cudaMalloc((void**)&ctx->aux[slot], bytes);
  if (e != cudaSuccess) return fail(ctx, PRE3_ERR_ALLOC, std::string("cudaMalloc aux buffer: ") + cudaGetErrorString(e));
  ctx->aux_cap[slot] = bytes;
  return PRE3_OK;
}

inline int ensure_copy_stream(pre3_ctx* ctx) {
  if (ctx->copy_stream) return PRE3_OK;
  PRE3_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    PRE3_CUDA(cudaEventCreateWithFlags(&ctx->ev_slot_full[i], cudaEventDisableTiming));
    PRE3_CUDA(cudaEventCreateWithFlags(&ctx->ev_slot_free[i], cudaEventDisableTiming));
  }
  return PRE3_OK;
}

// Bump allocation.  The arena is sized by the *_workspace_bytes estimates; a request that does not fit (an estimate
// that drifted from the sequence of takes) is served from a block of its own instead of running past the arena --
// freed at the next ws_reserve -- and reported under PRE3_DEBUG.
template <typename T>
inline T* ws_take(pre3_ctx* ctx, size_t count) {
  const size_t off = align_up(ctx->ws_off);
  const size_t bytes = count * sizeof(T);
  if (off + bytes > ctx->ws_cap || !ctx->ws) {
    char* blk = nullptr;
    if (cudaMalloc((void**)&blk, bytes ? align_up(bytes, 1024) : 1024) != cudaSuccess) {
      cudaGetLastError();
      ctx->err = "workspace overflow and cudaMalloc of the spill block failed";
      return nullptr;
    }
    ctx->retired.push_back(blk);
    if (getenv("PRE3_DEBUG")) fprintf(stderr, "[pre3] workspace estimate short by %zu bytes (served separately)\n", off + bytes - ctx->ws_cap);
    return reinterpret_cast<T*>(blk);
  }
  ctx->ws_off = off + bytes;
  return reinterpret_cast<T*>(ctx->ws + off);
}

inline void count_launch(pre3_ctx* ctx, int n = 1) { ctx->launches += n; }

// Pipeline stages of a whole-pair call (pre3_set_pipeline).  No-ops unless a pipeline is running.
enum PipeStage { PS_CONVERT = 0, PS_MATCH = 1, PS_EVAL = 2, PS_SELECT = 3 };
// The launches that follow belong to `stage`: they are issued on that stage's stream, after everything issued so far
// for the current chunk (event from the stream being left).
inline int pipe_enter(pre3_ctx* ctx, int stage) {
  if (ctx->pipe_stage < 0 || stage == ctx->pipe_stage) return PRE3_OK;
  cudaStream_t next = ctx->pipe_stream[ctx->pipe_map[stage]];
  if (next != ctx->stream) {
    cudaEvent_t ev = ctx->pipe_ev[ctx->pipe_stage];
    PRE3_CUDA(cudaEventRecord(ev, ctx->stream));
    PRE3_CUDA(cudaStreamWaitEvent(next, ev, 0));
    ctx->stream = next;
  }
  ctx->pipe_stage = stage;
  return PRE3_OK;
}
// number of pairs the occupancy heuristics should assume are in flight
inline int pipe_pairs(const pre3_ctx* ctx, int P) { return ctx->pipe_stage >= 0 && ctx->pipe_total_P > P ? ctx->pipe_total_P : P; }

inline cudaEvent_t timing_event(pre3_ctx* ctx) {
  if (ctx->ev_used == ctx->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->ev_pool.push_back(e);
  }
  return ctx->ev_pool[ctx->ev_used++];
}

// RAII: brackets the kernel launch(es) issued in its scope with two events on the context's
// stream when timing is enabled; free otherwise.
struct Span {
  pre3_ctx* ctx;
  TimedSpan s;
  Span(pre3_ctx* c, int cat) : ctx(c) {
    if (!ctx->timing) return;
    s.cat = cat;
    s.a = timing_event(ctx);
    s.b = timing_event(ctx);
    cudaEventRecord(s.a, ctx->stream);
  }
  ~Span() {
    if (!ctx->timing) return;
    cudaEventRecord(s.b, ctx->stream);
    ctx->spans.push_back(s);
  }
};

}  // namespace pre3
