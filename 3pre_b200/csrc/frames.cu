// frames.cu -- the step BEFORE the matching path: SR4000 frame batches -> per-feature
// 3-D points.  Specification: oracle/pre3_oracle_frames.c.
//
// Reference:
//   M/read_xyz_sr4000.m:8-21                 z, x, y = imfilter(., fspecial('gaussian',[3 3],2), 'same') (zero padding)
//   M/code_from_dr_ye/read_sr4000_data_dr_ye.m:8,88-90   the same with sigma 1 and 'replicate'
//   M/inittialize_depth_my_version.m:16,40-85  x(round(v),round(u)) lookup; reject NaN, closer than 0.4 m,
//                                              confidence <= max/2; [-x,-y,z]
//   M/SIFT_extract_save.m:55-56,75-88        frames + 1; XYZ_DATA / Descriptor / SCALE_ORIENT_POS of the survivors
//   M/code_from_dr_ye/confidence_filtering.m:1-13, ransac_dr_ye.m:13-19  the dr_ye flavour (confidence < max/2
//                                              removes the feature, no other rejection)
//
// sr_data layout (MATLAB `load` of d1_%04d.dat, column-major): rows x 176 doubles per frame with rows = 576
// (z | x | y | amplitude), 720 (+ confidence) or 721 (+ time stamp row).
//
//   k_conf_max      per frame: max(confidence_map(:)) (NaN ignored, like MATLAB max)           -- streams 203 KB / frame
//   k_smooth_maps   per frame: the three filtered 144 x 176 maps (read_xyz_sr4000's outputs)   -- streams 608 KB in, 608 KB out
//   k_feature_xyz   per frame: FUSED lookup -- the 3 x 3 stencil is evaluated only at the pixels the features round
//                   to (27 taps x K features instead of 76 k pixels x 27 taps), rejection tests, survivors compacted
//                   in feature order (block scan) -- nothing the consumer does not read is ever written
//   k_gather_cols   Descriptor(:, idxRemain) / SCALE_ORIENT_POS(:, idxRemain): column gather, HBM bound
// Tap order of the stencil (fixed, shared with the oracle): acc = 0; for dc = -1..1, for dr = -1..1:
// acc = acc + h(dr,dc) * A(r+dr, c+dc); out-of-range taps read 0.0 (zero padding) or the clamped pixel (replicate).
// Compiled with -fmad=false.
#include <cmath>

#include "common.cuh"

namespace pre3 {
namespace {

constexpr int FR_ROWS = 144, FR_COLS = 176, FR_PIX = FR_ROWS * FR_COLS;

struct Taps {
  double h[9];  // h[(dr+1) + 3*(dc+1)]
};

// fspecial('gaussian', [3 3], sigma): h = exp(-(x^2+y^2)/(2 sigma^2)); h(h < eps*max(h(:))) = 0; h = h / sum(h(:))
Taps gaussian3(double sigma) {
  Taps t;
  double mx = 0.0;
  for (int dc = -1; dc <= 1; ++dc)
    for (int dr = -1; dr <= 1; ++dr) {
      const double arg = -((double)(dc * dc) + (double)(dr * dr)) / (2.0 * sigma * sigma);
      const double v = std::exp(arg);
      t.h[(dr + 1) + 3 * (dc + 1)] = v;
      mx = v > mx ? v : mx;
    }
  double sum = 0.0;
  for (int i = 0; i < 9; ++i) {
    if (t.h[i] < 2.220446049250313e-16 * mx) t.h[i] = 0.0;
    sum = sum + t.h[i];
  }
  if (sum != 0.0)
    for (int i = 0; i < 9; ++i) t.h[i] = t.h[i] / sum;
  return t;
}

template <int BOUNDARY>
__device__ __forceinline__ double stencil(const double* __restrict__ map, int ld, int r, int c, const Taps& t) {
  double acc = 0.0;
#pragma unroll
  for (int dc = -1; dc <= 1; ++dc)
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr) {
      int rr = r + dr, cc = c + dc;
      double v;
      if (BOUNDARY == 1) {
        rr = min(max(rr, 0), FR_ROWS - 1);
        cc = min(max(cc, 0), FR_COLS - 1);
        v = map[(size_t)cc * ld + rr];
      } else {
        v = (rr >= 0 && rr < FR_ROWS && cc >= 0 && cc < FR_COLS) ? map[(size_t)cc * ld + rr] : 0.0;
      }
      acc = acc + t.h[(dr + 1) + 3 * (dc + 1)] * v;
    }
  return acc;
}

// 288 threads: thread = (row, column parity); every column is one contiguous 1152-byte read
__global__ void __launch_bounds__(2 * FR_ROWS) k_conf_max(const double* __restrict__ sr, int rows, int F,
                                                          double* __restrict__ max_conf) {
  const int f = blockIdx.x;
  const double* cm = sr + (size_t)f * rows * FR_COLS + 4 * FR_ROWS;
  const int r = threadIdx.x % FR_ROWS, c0 = threadIdx.x / FR_ROWS;
  double m = -INFINITY;
  bool any = false;
#pragma unroll 8
  for (int c = c0; c < FR_COLS; c += 2) {
    const double v = __ldg(cm + (size_t)c * rows + r);
    if (v == v) {  // max ignores NaN
      m = any ? fmax(m, v) : v;
      any = true;
    }
  }
  __shared__ double s_m[2 * FR_ROWS];
  __shared__ int s_a[2 * FR_ROWS];
  s_m[threadIdx.x] = m;
  s_a[threadIdx.x] = any;
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int i = threadIdx.x + 32; i < 2 * FR_ROWS; i += 32)
      if (s_a[i]) {
        m = any ? fmax(m, s_m[i]) : s_m[i];
        any = true;
      }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double om = __shfl_xor_sync(0xffffffffu, m, off);
      const int oa = __shfl_xor_sync(0xffffffffu, (int)any, off);
      if (oa) {
        m = any ? fmax(m, om) : om;
        any = true;
      }
    }
    if (threadIdx.x == 0) max_conf[f] = any ? m : NAN;  // all-NaN map: max is NaN
  }
}

// grid (column tiles, 3 maps, F); block 144 threads = one thread per image row marching along SM_TC columns with
// a 3 x 3 register window: three loads per output (the row above / below come out of L1: the neighbouring threads
// read the same 128-byte lines), no shared memory, no barrier.  Reads and writes are 1152-byte contiguous columns.
constexpr int SM_TC = 44;  // 176 = 4 x 44: one halo column per 44
template <int BOUNDARY>
__global__ void __launch_bounds__(FR_ROWS) k_smooth_maps(const double* __restrict__ sr, int rows, Taps t,
                                                         double* __restrict__ ox, double* __restrict__ oy,
                                                         double* __restrict__ oz) {
  const int f = blockIdx.z, which = blockIdx.y, c0 = blockIdx.x * SM_TC;
  // sr_data rows: z 0..143, x 144..287, y 288..431 (read_xyz_sr4000.m:10-12)
  const double* map = sr + (size_t)f * rows * FR_COLS + (which == 0 ? FR_ROWS : which == 1 ? 2 * FR_ROWS : 0);
  double* out = (which == 0 ? ox : which == 1 ? oy : oz) + (size_t)f * FR_PIX;
  const int r = threadIdx.x;
  int rr[3];
  bool rin[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int q = r + d - 1;
    rin[d] = q >= 0 && q < FR_ROWS;
    rr[d] = min(max(q, 0), FR_ROWS - 1);
  }
  auto load_col = [&](int c, double* v) {
    const bool cin = c >= 0 && c < FR_COLS;
    const double* col = map + (size_t)min(max(c, 0), FR_COLS - 1) * rows;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double x = __ldg(col + rr[d]);
      v[d] = (BOUNDARY == 1 || (cin && rin[d])) ? x : 0.0;
    }
  };
  double w[3][3];  // w[column slot][row offset]: columns c-1, c, c+1
  load_col(c0 - 1, w[0]);
  load_col(c0, w[1]);
#pragma unroll 4
  for (int j = 0; j < SM_TC; ++j) {
    const int c = c0 + j;
    load_col(c + 1, w[2]);
    double acc = 0.0;
#pragma unroll
    for (int dc = 0; dc < 3; ++dc)
#pragma unroll
      for (int dr = 0; dr < 3; ++dr) acc = acc + t.h[dr + 3 * dc] * w[dc][dr];
    out[(size_t)c * FR_ROWS + r] = acc;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      w[0][d] = w[1][d];
      w[1][d] = w[2][d];
    }
  }
}

__device__ __forceinline__ long long mround(double v) {  // MATLAB round: half away from zero
  return (long long)round(v);
}

// one block per frame, thread per feature (K <= blockDim * passes); survivors compacted in feature order
template <int BOUNDARY>
__global__ void __launch_bounds__(256)
k_feature_xyz(const double* __restrict__ sr, int rows, Taps t, int mode, int use_conf,
              const double* __restrict__ max_conf, const double* __restrict__ frames, int frame_ld, int K,
              const int32_t* __restrict__ k_count, double* __restrict__ xyz_all, uint8_t* __restrict__ keep_out,
              int32_t* __restrict__ n_keep, int32_t* __restrict__ idx_remain, double* __restrict__ xyz,
              int32_t* __restrict__ n_oob) {
  __shared__ int s_w[8];
  __shared__ int s_base;
  const int f = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* base = sr + (size_t)f * rows * FR_COLS;
  const double* zm = base;
  const double* xm = base + FR_ROWS;
  const double* ym = base + 2 * FR_ROWS;
  const bool has_conf = rows >= 720;  // read_xyz_sr4000.m:25
  const double* cm = base + 4 * FR_ROWS;
  const double mc = has_conf ? max_conf[f] : 0.0;
  int Kf = k_count ? k_count[f] : K;
  Kf = max(0, min(Kf, K));
  if (tid == 0) s_base = 0;
  int oob = 0;
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += 256) {
    const int k = k0 + tid;
    bool keep = false;
    double px = NAN, py = NAN, pz = NAN;
    if (k < Kf) {
      const double* fr = frames + ((size_t)f * K + k) * frame_ld;
      // frames(1:2,:) + 1 (SIFT_extract_save.m:55-56, vodometry_dr_ye.m:73-74); uv = [uvd(2), uvd(1)]
      // (inittialize_depth_my_version.m:16): row = round(y + 1), column = round(x + 1), 1-based
      const long long c1 = mround(fr[0] + 1.0), r1 = mround(fr[1] + 1.0);
      if (r1 >= 1 && r1 <= FR_ROWS && c1 >= 1 && c1 <= FR_COLS) {
        const int r = (int)r1 - 1, c = (int)c1 - 1;
        const double xf = stencil<BOUNDARY>(xm, rows, r, c, t);
        const double yf = stencil<BOUNDARY>(ym, rows, r, c, t);
        const double zf = stencil<BOUNDARY>(zm, rows, r, c, t);
        const double conf = has_conf ? cm[(size_t)c * rows + r] : 0.0;
        if (mode == 0) {
          // inittialize_depth_my_version.m:40-79
          if (!(xf != xf)) {
            const double df = sqrt((xf * xf + yf * yf) + zf * zf);
            keep = !(df < 0.4 || (has_conf && conf <= (2.0 / 4.0) * mc));
          }
        } else {
          // confidence_filtering.m:5-11 (only when myCONFIG.FLAGS.CONFIDENCE_MAP), then the plain lookup
          keep = !(use_conf && has_conf && conf < 0.5 * mc);
        }
        if (keep) {
          px = -xf;
          py = -yf;
          pz = zf;
        }
      } else {
        ++oob;  // the reference raises an index error here
      }
    }
    if (k < K) {
      if (xyz_all) {
        double* o = xyz_all + ((size_t)f * K + k) * 3;
        o[0] = px;
        o[1] = py;
        o[2] = pz;
      }
      if (keep_out) keep_out[(size_t)f * K + k] = keep ? 1 : 0;
    }
    // ordered compaction: idxRemain = [idxRemain, idxFrame] (SIFT_extract_save.m:82)
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    int off = s_base + __popc(bal & ((1u << lane) - 1u));
    int tot = 0;
    for (int w = 0; w < 8; ++w) {
      if (w < warp) off += s_w[w];
      tot += s_w[w];
    }
    if (keep) {
      if (idx_remain) idx_remain[(size_t)f * K + off] = k;
      if (xyz) {
        double* o = xyz + ((size_t)f * K + off) * 3;
        o[0] = px;
        o[1] = py;
        o[2] = pz;
      }
    }
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
  // padding of the compacted outputs
  const int nk = s_base;
  for (int k = nk + tid; k < K; k += 256) {
    if (idx_remain) idx_remain[(size_t)f * K + k] = -1;
    if (xyz) {
      double* o = xyz + ((size_t)f * K + k) * 3;
      o[0] = o[1] = o[2] = NAN;
    }
  }
  if (tid == 0 && n_keep) n_keep[f] = nk;
  if (oob && n_oob) atomicAdd(n_oob, oob);
}

// out(:, j) = in(:, idx(j)) per frame, `bytes` per column; zero beyond n_keep.  One warp per column.
__global__ void __launch_bounds__(256) k_gather_cols(const unsigned char* __restrict__ in,
                                                     unsigned char* __restrict__ out, int K, int bytes,
                                                     const int32_t* __restrict__ idx, int F) {
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= F * K) return;
  const int f = wid / K;
  const int src = idx[wid];
  unsigned char* o = out + (size_t)wid * bytes;
  if ((bytes & 15) == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0) {
    const uint4* s = src >= 0 ? reinterpret_cast<const uint4*>(in + ((size_t)f * K + src) * bytes) : nullptr;
    uint4* d = reinterpret_cast<uint4*>(o);
    for (int i = lane; i < bytes / 16; i += 32) d[i] = s ? s[i] : make_uint4(0, 0, 0, 0);
  } else {
    const unsigned char* s = src >= 0 ? in + ((size_t)f * K + src) * bytes : nullptr;
    for (int i = lane; i < bytes; i += 32) o[i] = s ? s[i] : 0;
  }
}

int check_frame_opts(pre3_ctx* ctx, const pre3_frame_opts* o) {
  if (!o) return fail(ctx, PRE3_ERR_ARG, "frame options missing");
  if (!(o->sigma > 0.0)) return fail(ctx, PRE3_ERR_ARG, "sigma must be positive");
  if (o->boundary != 0 && o->boundary != 1) return fail(ctx, PRE3_ERR_ARG, "boundary must be 0 (zero) or 1 (replicate)");
  if (o->mode != 0 && o->mode != 1) return fail(ctx, PRE3_ERR_ARG, "mode must be 0 (SIFT_extract_save) or 1 (dr_ye)");
  if (o->rows != 576 && o->rows != 720 && o->rows != 721)
    return fail(ctx, PRE3_ERR_ARG, "sr_data must have 576, 720 or 721 rows");
  return PRE3_OK;
}

size_t cls_bytes(int cls) {
  switch (cls) {
    case PRE3_CLASS_DOUBLE: return 8;
    case PRE3_CLASS_SINGLE: return 4;
    case PRE3_CLASS_INT8:
    case PRE3_CLASS_UINT8: return 1;
    default: return 0;
  }
}

#define FR_LIVE()                                                                                          \
  do {                                                                                                     \
    if (!ctx) return PRE3_ERR_ARG;                                                                         \
    if (ctx->device < 0) return fail(ctx, PRE3_ERR_CUDA, "no CUDA device (libpre3 has no CPU fallback)"); \
    PRE3_CUDA(cudaSetDevice(ctx->device));                                                                 \
  } while (0)

int smooth_impl(pre3_ctx* ctx, const double* dsr, int F, const pre3_frame_opts& o, double* dx, double* dy, double* dz,
                double* dmaxc) {
  Span span__(ctx, T_FRAMES);
  const Taps t = gaussian3(o.sigma);
  if (dmaxc && o.rows >= 720) {
    k_conf_max<<<F, 2 * FR_ROWS, 0, ctx->stream>>>(dsr, o.rows, F, dmaxc);
    count_launch(ctx);
  }
  if (dx && dy && dz) {
    dim3 grid(FR_COLS / SM_TC, 3, F);
    if (o.boundary == 1)
      k_smooth_maps<1><<<grid, FR_ROWS, 0, ctx->stream>>>(dsr, o.rows, t, dx, dy, dz);
    else
      k_smooth_maps<0><<<grid, FR_ROWS, 0, ctx->stream>>>(dsr, o.rows, t, dx, dy, dz);
    count_launch(ctx);
  }
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

// workspace: max_conf (F doubles) + oob counter
int features_impl(pre3_ctx* ctx, const double* dsr, int F, const pre3_frame_opts& o, const double* dframes,
                  int frame_ld, int K, const int32_t* dk_count, double* dxyz_all, uint8_t* dkeep, int32_t* dn_keep,
                  int32_t* didx, double* dxyz, const void* ddesc_in, int cls, int ND, void* ddesc_out,
                  const double* dfr_in, double* dfr_out, int32_t* dn_oob) {
  Span span__(ctx, T_FRAMES);
  const Taps t = gaussian3(o.sigma);
  double* dmaxc = ws_take<double>(ctx, F);
  int32_t* idx = didx ? didx : ws_take<int32_t>(ctx, (size_t)F * K);
  if (dn_oob) PRE3_CUDA(cudaMemsetAsync(dn_oob, 0, sizeof(int32_t), ctx->stream));
  if (o.rows >= 720) {
    k_conf_max<<<F, 2 * FR_ROWS, 0, ctx->stream>>>(dsr, o.rows, F, dmaxc);
    count_launch(ctx);
  }
  if (o.boundary == 1)
    k_feature_xyz<1><<<F, 256, 0, ctx->stream>>>(dsr, o.rows, t, o.mode, o.use_confidence, dmaxc, dframes, frame_ld, K,
                                                 dk_count, dxyz_all, dkeep, dn_keep, idx, dxyz, dn_oob);
  else
    k_feature_xyz<0><<<F, 256, 0, ctx->stream>>>(dsr, o.rows, t, o.mode, o.use_confidence, dmaxc, dframes, frame_ld, K,
                                                 dk_count, dxyz_all, dkeep, dn_keep, idx, dxyz, dn_oob);
  count_launch(ctx);
  const int blocks = (int)(((size_t)F * K * 32 + 255) / 256);
  if (ddesc_in && ddesc_out) {
    k_gather_cols<<<blocks, 256, 0, ctx->stream>>>((const unsigned char*)ddesc_in, (unsigned char*)ddesc_out, K,
                                                   (int)(cls_bytes(cls) * ND), idx, F);
    count_launch(ctx);
  }
  if (dfr_in && dfr_out) {
    k_gather_cols<<<blocks, 256, 0, ctx->stream>>>((const unsigned char*)dfr_in, (unsigned char*)dfr_out, K,
                                                   frame_ld * 8, idx, F);
    count_launch(ctx);
  }
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

}  // namespace
}  // namespace pre3

using namespace pre3;

extern "C" {

int pre3_read_xyz_sr4000_batch_dev(pre3_ctx* ctx, const double* dsr_data, int F, const pre3_frame_opts* opts,
                                   double* dx, double* dy, double* dz, double* dmax_conf) {
  FR_LIVE();
  PRE3_TRY(check_frame_opts(ctx, opts));
  if (F < 0) return fail(ctx, PRE3_ERR_ARG, "negative frame count");
  if (F == 0) return PRE3_OK;
  if (!dsr_data) return fail(ctx, PRE3_ERR_ARG, "sr_data missing");
  return smooth_impl(ctx, dsr_data, F, *opts, dx, dy, dz, dmax_conf);
}

int pre3_read_xyz_sr4000_batch(pre3_ctx* ctx, const double* sr_data, int F, const pre3_frame_opts* opts, double* x,
                               double* y, double* z, double* max_conf) {
  FR_LIVE();
  PRE3_TRY(check_frame_opts(ctx, opts));
  if (F < 0) return fail(ctx, PRE3_ERR_ARG, "negative frame count");
  if (F == 0) return PRE3_OK;
  if (!sr_data || !x || !y || !z) return fail(ctx, PRE3_ERR_ARG, "sr_data / output maps missing");
  const size_t sb = (size_t)F * opts->rows * FR_COLS * 8, mb = (size_t)F * FR_PIX * 8;
  PRE3_TRY(ws_reserve(ctx, align_up(sb) + 3 * align_up(mb) + align_up(8 * (size_t)F) + 4096));
  double* dsr = ws_take<double>(ctx, sb / 8);
  double* dx = ws_take<double>(ctx, mb / 8);
  double* dy = ws_take<double>(ctx, mb / 8);
  double* dz = ws_take<double>(ctx, mb / 8);
  double* dmc = ws_take<double>(ctx, F);
  PRE3_CUDA(cudaMemcpyAsync(dsr, sr_data, sb, cudaMemcpyHostToDevice, ctx->stream));
  PRE3_TRY(smooth_impl(ctx, dsr, F, *opts, dx, dy, dz, max_conf ? dmc : nullptr));
  PRE3_CUDA(cudaMemcpyAsync(x, dx, mb, cudaMemcpyDeviceToHost, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(y, dy, mb, cudaMemcpyDeviceToHost, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(z, dz, mb, cudaMemcpyDeviceToHost, ctx->stream));
  if (max_conf) {
    if (opts->rows >= 720) {
      PRE3_CUDA(cudaMemcpyAsync(max_conf, dmc, 8 * (size_t)F, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      for (int f = 0; f < F; ++f) max_conf[f] = NAN;
    }
  }
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

int pre3_features_xyz_batch_dev(pre3_ctx* ctx, const double* dsr_data, int F, const pre3_frame_opts* opts,
                                const double* dframes, int frame_ld, int K, const int32_t* dk_count, double* dxyz_all,
                                uint8_t* dkeep, int32_t* dn_keep, int32_t* didx_remain, double* dxyz,
                                const void* ddesc_in, int cls, int ND, void* ddesc_out, const double* dframes_in,
                                double* dframes_out, int32_t* dn_oob) {
  FR_LIVE();
  PRE3_TRY(check_frame_opts(ctx, opts));
  if (F < 0 || K < 0 || frame_ld < 2) return fail(ctx, PRE3_ERR_ARG, "bad sizes (frames need at least 2 rows)");
  if (F == 0 || K == 0) return PRE3_OK;
  if (!dsr_data || !dframes) return fail(ctx, PRE3_ERR_ARG, "sr_data / frames missing");
  if (ddesc_in && cls_bytes(cls) == 0) return fail(ctx, PRE3_ERR_CLASS, "Unsupported numeric class");
  PRE3_TRY(ws_reserve(ctx, align_up(8 * (size_t)F) + align_up(4 * (size_t)F * K) + 4096));
  return features_impl(ctx, dsr_data, F, *opts, dframes, frame_ld, K, dk_count, dxyz_all, dkeep, dn_keep, didx_remain,
                       dxyz, ddesc_in, cls, ND, ddesc_out, dframes_in, dframes_out, dn_oob);
}

int pre3_features_xyz_batch(pre3_ctx* ctx, const double* sr_data, int F, const pre3_frame_opts* opts,
                            const double* frames, int frame_ld, int K, const int32_t* k_count, double* xyz_all,
                            uint8_t* keep, int32_t* n_keep, int32_t* idx_remain, double* xyz, const void* desc_in,
                            int cls, int ND, void* desc_out, double* frames_out, int32_t* n_oob) {
  FR_LIVE();
  PRE3_TRY(check_frame_opts(ctx, opts));
  if (F < 0 || K < 0 || frame_ld < 2) return fail(ctx, PRE3_ERR_ARG, "bad sizes (frames need at least 2 rows)");
  if (F == 0) return PRE3_OK;
  if (K == 0) {
    if (n_keep)
      for (int f = 0; f < F; ++f) n_keep[f] = 0;
    if (n_oob) *n_oob = 0;
    return PRE3_OK;
  }
  if (!sr_data || !frames) return fail(ctx, PRE3_ERR_ARG, "sr_data / frames missing");
  if (desc_in && cls_bytes(cls) == 0) return fail(ctx, PRE3_ERR_CLASS, "Unsupported numeric class");
  const size_t FK = (size_t)F * K;
  const size_t sb = (size_t)F * opts->rows * FR_COLS * 8, fb = FK * frame_ld * 8, xb = FK * 24;
  const size_t db = desc_in ? FK * cls_bytes(cls) * ND : 0;
  PRE3_TRY(ws_reserve(ctx, align_up(sb) + 2 * align_up(fb) + 2 * align_up(xb) + 2 * align_up(db) + align_up(FK) +
                               2 * align_up(4 * FK) + 2 * align_up(4 * (size_t)F) + align_up(8 * (size_t)F) + 8192));
  double* dsr = ws_take<double>(ctx, sb / 8);
  double* dfr = ws_take<double>(ctx, fb / 8);
  double* dfro = frames_out ? ws_take<double>(ctx, fb / 8) : nullptr;
  double* dxa = ws_take<double>(ctx, FK * 3);
  double* dxc = ws_take<double>(ctx, FK * 3);
  unsigned char* ddi = desc_in ? ws_take<unsigned char>(ctx, db) : nullptr;
  unsigned char* ddo = (desc_in && desc_out) ? ws_take<unsigned char>(ctx, db) : nullptr;
  uint8_t* dkeep = ws_take<uint8_t>(ctx, FK);
  int32_t* didx = ws_take<int32_t>(ctx, FK);
  int32_t* dnk = ws_take<int32_t>(ctx, F);
  int32_t* dkc = k_count ? ws_take<int32_t>(ctx, F) : nullptr;
  int32_t* doob = ws_take<int32_t>(ctx, 1);
  PRE3_CUDA(cudaMemcpyAsync(dsr, sr_data, sb, cudaMemcpyHostToDevice, ctx->stream));
  PRE3_CUDA(cudaMemcpyAsync(dfr, frames, fb, cudaMemcpyHostToDevice, ctx->stream));
  if (ddi) PRE3_CUDA(cudaMemcpyAsync(ddi, desc_in, db, cudaMemcpyHostToDevice, ctx->stream));
  if (dkc) PRE3_CUDA(cudaMemcpyAsync(dkc, k_count, 4 * (size_t)F, cudaMemcpyHostToDevice, ctx->stream));
  PRE3_TRY(features_impl(ctx, dsr, F, *opts, dfr, frame_ld, K, dkc, dxa, dkeep, dnk, didx, dxc, ddi, cls, ND, ddo,
                         dfro ? dfr : nullptr, dfro, doob));
  if (xyz_all) PRE3_CUDA(cudaMemcpyAsync(xyz_all, dxa, xb, cudaMemcpyDeviceToHost, ctx->stream));
  if (xyz) PRE3_CUDA(cudaMemcpyAsync(xyz, dxc, xb, cudaMemcpyDeviceToHost, ctx->stream));
  if (keep) PRE3_CUDA(cudaMemcpyAsync(keep, dkeep, FK, cudaMemcpyDeviceToHost, ctx->stream));
  if (idx_remain) PRE3_CUDA(cudaMemcpyAsync(idx_remain, didx, 4 * FK, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_keep) PRE3_CUDA(cudaMemcpyAsync(n_keep, dnk, 4 * (size_t)F, cudaMemcpyDeviceToHost, ctx->stream));
  if (ddo) PRE3_CUDA(cudaMemcpyAsync(desc_out, ddo, db, cudaMemcpyDeviceToHost, ctx->stream));
  if (dfro) PRE3_CUDA(cudaMemcpyAsync(frames_out, dfro, fb, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_oob) PRE3_CUDA(cudaMemcpyAsync(n_oob, doob, 4, cudaMemcpyDeviceToHost, ctx->stream));
  PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
  return PRE3_OK;
}

}  // extern "C"
