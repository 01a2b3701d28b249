// peak.cu -- FP32 FFMA-chain microbenchmark: the denominator of the scoring kernel's roofline
// (MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only; SURVEY.md 8d asks for the FP32
// CUDA-core peak to be measured on the box).
#include "common.cuh"

namespace pre3 {

__global__ void __launch_bounds__(256) k_fp32_peak(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
        x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fmaf(x0, a, b);
      x1 = fmaf(x1, a, b);
      x2 = fmaf(x2, a, b);
      x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b);
      x5 = fmaf(x5, a, b);
      x6 = fmaf(x6, a, b);
      x7 = fmaf(x7, a, b);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456f) out[0] = s;  // never true; keeps the chains alive
}

// Same chains with all three FFMA sources in registers (the form a scorer with per-thread hypotheses and
// per-match operands needs): the multiplier / addend come from memory, so ptxas cannot use the constant-bank
// or immediate forms.
__global__ void __launch_bounds__(256) k_fp32_peak_3reg(float* out, int iters, const float* __restrict__ ab) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
        x7 = x0 + 7.f;
  const float a0 = ab[threadIdx.x & 7], b0 = ab[8 + (threadIdx.x & 7)], a1 = ab[16 + (threadIdx.x & 3)],
              b1 = ab[24 + (threadIdx.x & 3)];
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x0 = fmaf(x0, a0, b0);
      x1 = fmaf(x1, a1, b1);
      x2 = fmaf(x2, a0, b1);
      x3 = fmaf(x3, a1, b0);
      x4 = fmaf(x4, b0, a0);
      x5 = fmaf(x5, b1, a1);
      x6 = fmaf(x6, b0, a1);
      x7 = fmaf(x7, b1, a0);
      x0 = fmaf(x0, a1, b1);
      x1 = fmaf(x1, a0, b0);
      x2 = fmaf(x2, a1, b0);
      x3 = fmaf(x3, a0, b1);
      x4 = fmaf(x4, b1, a1);
      x5 = fmaf(x5, b0, a0);
      x6 = fmaf(x6, b1, a0);
      x7 = fmaf(x7, b0, a1);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456f) out[0] = s;
}

// FP64 DFMA chains: the denominator of the EKF kernels' roofline.
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0,
         x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;
}

// TMEM -> register read rate (tcgen05.ld): the bound of a GEMM whose every fp32 accumulator must be looked at
// by the epilogue (k_tc_gemm_top2: 4 bytes of TMEM per 256 flops at K = 128).  One CTA per SM, 8 warps (two per
// TMEM lane quarter), each sweeping the 512 columns with 32x32b.x32 loads, two in flight.
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tm_wait(uint32_t (&a)[32], uint32_t (&b)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[7]), "+r"(a[15]), "+r"(a[23]), "+r"(a[31]), "+r"(b[0]), "+r"(b[7]), "+r"(b[15]),
                 "+r"(b[23]), "+r"(b[31])
               :
               : "memory");
}

__global__ void __launch_bounds__(256, 1) k_tmem_read(uint32_t* out, int iters) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&s_base))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_base;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t col0 = (uint32_t)((warp >> 2) * 256);  // the two warps of a quarter split the columns
  uint32_t acc = 0;
  uint32_t a[32], b[32];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 256; c += 64) {
      tm_ld32(lane_base + col0 + c, a);
      tm_ld32(lane_base + col0 + c + 32, b);
      tm_wait(a, b);
      acc ^= a[0] ^ a[31] ^ b[0] ^ b[31];
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace pre3

// TMEM read bandwidth: GB/s over the whole chip and bytes per clock per SM (at sm_mhz, 0 = skip)
extern "C" int pre3_measure_tmem_read(pre3_ctx* ctx, double* gbs) {
  using namespace pre3;
  if (!ctx || ctx->device < 0 || !gbs) return PRE3_ERR_CUDA;
  PRE3_CUDA(cudaSetDevice(ctx->device));
  PRE3_TRY(ws_reserve(ctx, 4096));
  uint32_t* out = ws_take<uint32_t>(ctx, 16);
  const int blocks = ctx->sm_count, iters = 2000;
  cudaEvent_t a, b;
  PRE3_CUDA(cudaEventCreate(&a));
  PRE3_CUDA(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    PRE3_CUDA(cudaEventRecord(a, ctx->stream));
    k_tmem_read<<<blocks, 256, 0, ctx->stream>>>(out, iters);
    PRE3_CUDA(cudaEventRecord(b, ctx->stream));
    PRE3_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    PRE3_CUDA(cudaEventElapsedTime(&ms, a, b));
    const double bytes = 128.0 * 512.0 * 4.0 * (double)iters * (double)blocks;
    if (rep > 0 && ms > 0.f) best = std::max(best, bytes / (ms * 1e-3) / 1e9);
  }
  count_launch(ctx, 4);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  PRE3_CUDA(cudaGetLastError());
  *gbs = best;
  return PRE3_OK;
}

extern "C" int pre3_measure_fp64_peak(pre3_ctx* ctx, double* tflops) {
  using namespace pre3;
  if (!ctx || ctx->device < 0 || !tflops) return PRE3_ERR_CUDA;
  PRE3_CUDA(cudaSetDevice(ctx->device));
  PRE3_TRY(ws_reserve(ctx, 4096));
  double* out = ws_take<double>(ctx, 16);
  const int blocks = ctx->sm_count * 8, iters = 1024;
  cudaEvent_t a, b;
  PRE3_CUDA(cudaEventCreate(&a));
  PRE3_CUDA(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    PRE3_CUDA(cudaEventRecord(a, ctx->stream));
    k_fp64_peak<<<blocks, 256, 0, ctx->stream>>>(out, iters, 0.999, 0.001);
    PRE3_CUDA(cudaEventRecord(b, ctx->stream));
    PRE3_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    PRE3_CUDA(cudaEventElapsedTime(&ms, a, b));
    const double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
    if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  count_launch(ctx, 5);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *tflops = best;
  return PRE3_OK;
}

// mode 0: constant-bank operands (the classic peak); mode 1: three register operands
extern "C" int pre3_measure_fp32_peak_mode(pre3_ctx* ctx, int mode, double* tflops) {
  using namespace pre3;
  if (!ctx || ctx->device < 0 || !tflops) return PRE3_ERR_CUDA;
  PRE3_CUDA(cudaSetDevice(ctx->device));
  PRE3_TRY(ws_reserve(ctx, 8192));
  float* out = ws_take<float>(ctx, 16);
  float* ab = ws_take<float>(ctx, 32);
  float hab[32];
  for (int i = 0; i < 32; ++i) hab[i] = (i & 8) ? 0.001f + 1e-6f * i : 0.999f - 1e-6f * i;
  PRE3_CUDA(cudaMemcpyAsync(ab, hab, sizeof hab, cudaMemcpyHostToDevice, ctx->stream));
  const int blocks = ctx->sm_count * 8, iters = 4096;
  cudaEvent_t a, b;
  PRE3_CUDA(cudaEventCreate(&a));
  PRE3_CUDA(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    PRE3_CUDA(cudaEventRecord(a, ctx->stream));
    if (mode == 0)
      k_fp32_peak<<<blocks, 256, 0, ctx->stream>>>(out, iters, 0.999f, 0.001f);
    else
      k_fp32_peak_3reg<<<blocks, 256, 0, ctx->stream>>>(out, iters, ab);
    PRE3_CUDA(cudaEventRecord(b, ctx->stream));
    PRE3_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    PRE3_CUDA(cudaEventElapsedTime(&ms, a, b));
    const double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
    if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  count_launch(ctx, 5);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *tflops = best;
  return PRE3_OK;
}

extern "C" int pre3_measure_fp32_peak(pre3_ctx* ctx, double* tflops) {
  return pre3_measure_fp32_peak_mode(ctx, 0, tflops);
}
