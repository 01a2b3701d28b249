// pipebench.cu -- per-instruction throughput on B200 for the selection epilogue of the matcher (which pipe, which rate):
// each kernel runs 8 independent dependency chains per thread of ONE instruction kind, 1024 threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipebench.bin tools/pipebench.cu
// Prints warp-instructions per clock per SM sub-partition (1.0 = one per cycle = 32 lanes/clk).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048
#define NCH 8

template <int OP>
__device__ __forceinline__ void step(uint32_t (&v)[NCH], uint32_t a, uint32_t b) {
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    uint32_t x = v[i];
    // asm volatile pins the instruction: nothing is merged or folded
    if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 1) asm volatile("max.f32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 11) asm volatile("min.f32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 2) x = __float_as_uint(fmaxf(fmaxf(__uint_as_float(x), __uint_as_float(a)), __uint_as_float(b)));
    if (OP == 3) asm volatile("max.u32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 4) x = max(max(x, a), b);
    if (OP == 5) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 6) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 8) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
    if (OP == 9) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(x) : "r"(a));
    if (OP == 10) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(x) : "r"(a));
    if (OP == 12) asm volatile("shr.u32 %0, %0, 1;" : "+r"(x));
    v[i] = x;
  }
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) k_op(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t v[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) v[i] = threadIdx.x * 7 + i;
  for (int it = 0; it < ITERS; ++it) {
    step<OP>(v, a, b);
    a += 0x10003u;
    b ^= a;  // operands change every iteration: nothing folds (one extra integer add per 8 measured instructions)
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s ^= v[i];
  out[blockIdx.x * 1024 + threadIdx.x] = s;
}

// packed pairs: 4 chains of float2
template <int OP>
__global__ void __launch_bounds__(1024, 1) k_op2(float2* out, float2 a, float2 b) {
  float2 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      unsigned long long& x = *reinterpret_cast<unsigned long long*>(&v[i]);
      const unsigned long long a2 = *reinterpret_cast<const unsigned long long*>(&a);
      const unsigned long long b2 = *reinterpret_cast<const unsigned long long*>(&b);
      if (OP == 0) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(a2));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(a2), "l"(b2));
      if (OP == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(a2));
    }
    a.x += 1.0e-7f;
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) s.x += v[i].x, s.y += v[i].y;
  out[blockIdx.x * 1024 + threadIdx.x] = s;
}

// fp64 pipe and the 64-bit conversions the fused sequence matcher's converter warps are made of: 4 chains per thread,
// THREADS / 128 warps per sub-partition (the converters run 2 per sub-partition)
template <int OP, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_opd(double* out, double a, double b) {
  double v[4];
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = threadIdx.x * 0.001 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (OP == 0) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(v[i]) : "d"(a));
      if (OP == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(v[i]) : "d"(a));
      if (OP == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[i]) : "d"(a), "d"(b));
      if (OP == 3) {
        uint32_t f;
        asm volatile("cvt.rn.f32.f64 %0, %1;" : "=r"(f) : "d"(v[i]));
        acc ^= f;
      }
      if (OP == 4) {
        uint16_t h;
        asm volatile("cvt.rn.f16.f64 %0, %1;" : "=h"(h) : "d"(v[i]));
        acc ^= h;
      }
      if (OP == 5) {
        double d;
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "r"(acc + i));
        acc ^= (uint32_t)__double_as_longlong(d);
      }
    }
    a += 1.0e-9;
  }
  double s = (double)acc;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += v[i];
  out[blockIdx.x * THREADS + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F launch, double ops_per_thread) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) launch();
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  int mhz = 0;
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  const double clk = 1.965e9;  // B200 max SM clock (no throttling in these short ALU loops)
  const double warp_instr_per_smsp = ops_per_thread * 1024 / 32 / 4;   // per SM: 32 warps over 4 sub-partitions
  printf("%-34s %7.3f ms  %.3f warp-instr/clk/SMSP (%.1f lanes/clk/SM)\n", name, ms, warp_instr_per_smsp / (ms * 1e-3 * clk),
         ops_per_thread * 1024 / (ms * 1e-3 * clk));
}

int main() {
  uint32_t* out;
  cudaMalloc(&out, 148 * 1024 * 8);
  const double n = (double)ITERS * NCH;
#define RUN(OP, NAME) run(NAME, [&]() { k_op<OP><<<148, 1024>>>(out, 0x3f800123u, 0x12345u); }, n)
  RUN(0, "LOP3 (and/or)");
  RUN(1, "FMNMX (fmaxf)");
  RUN(11, "FMNMX (fminf)");
  RUN(2, "FMNMX3 (3-input fmaxf)");
  RUN(3, "VIMNMX.U32 (max)");
  RUN(4, "VIMNMX3.U32 (3-input max)");
  RUN(5, "FADD");
  RUN(6, "FFMA");
  RUN(7, "IADD3");
  RUN(8, "IMAD");
  RUN(9, "PRMT");
  RUN(10, "HMNMX2 (half2 max)");
  RUN(12, "SHF (shift)");
  const double n2 = (double)ITERS * 4;
  run("FADD2 (packed, per instr)", [&]() { k_op2<0><<<148, 1024>>>((float2*)out, make_float2(1.0f, 2.0f), make_float2(0.5f, 0.25f)); }, n2);
  run("FFMA2 (packed, per instr)", [&]() { k_op2<1><<<148, 1024>>>((float2*)out, make_float2(1.0f, 0.99f), make_float2(0.5f, 0.25f)); }, n2);
  run("FMUL2 (packed, per instr)", [&]() { k_op2<2><<<148, 1024>>>((float2*)out, make_float2(1.0f, 0.99f), make_float2(0.5f, 0.25f)); }, n2);
  // run() assumes 1024 threads per SM: scale the per-thread count by THREADS / 1024
#define RUND(OP, T, NAME) run(NAME, [&]() { k_opd<OP, T><<<148, T>>>((double*)out, 1.0000001, 0.5); }, n2 * T / 1024.0)
  RUND(0, 1024, "DMUL 8 warps/SMSP");
  RUND(0, 256, "DMUL 2 warps/SMSP");
  RUND(1, 1024, "DADD 8 warps/SMSP");
  RUND(1, 256, "DADD 2 warps/SMSP");
  RUND(2, 1024, "DFMA 8 warps/SMSP");
  RUND(2, 256, "DFMA 2 warps/SMSP");
  RUND(3, 1024, "F2F.F32.F64 (+LOP3) 8 warps/SMSP");
  RUND(3, 256, "F2F.F32.F64 (+LOP3) 2 warps/SMSP");
  RUND(4, 1024, "F2F.F16.F64 (+LOP3) 8 warps/SMSP");
  RUND(5, 1024, "F2F.F64.F32 (+2 ALU) 8 warps/SMSP");
  return 0;
}
