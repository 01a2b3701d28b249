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

}  // namespace pre3

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

extern "C" int pre3_measure_fp32_peak(pre3_ctx* ctx, double* tflops) {
  using namespace pre3;
  if (!ctx || ctx->device < 0 || !tflops) return PRE3_ERR_CUDA;
  PRE3_CUDA(cudaSetDevice(ctx->device));
  PRE3_TRY(ws_reserve(ctx, 4096));
  float* out = ws_take<float>(ctx, 16);
  const int blocks = ctx->sm_count * 8, iters = 4096;
  cudaEvent_t a, b;
  PRE3_CUDA(cudaEventCreate(&a));
  PRE3_CUDA(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    PRE3_CUDA(cudaEventRecord(a, ctx->stream));
    k_fp32_peak<<<blocks, 256, 0, ctx->stream>>>(out, iters, 0.999f, 0.001f);
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
