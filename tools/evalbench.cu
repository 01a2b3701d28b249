// evalbench.cu -- stand-alone timing of candidate inner loops for the support scorer (k_eval, csrc/ransac.cu).
// Not part of libpre3: a scratch harness to choose the thread / register mapping by measurement.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/evalbench.bin tools/evalbench.cu
// Every variant scores H hypotheses (R, t in fp32) against N correspondences tiled through shared memory and
// counts r^2 < thr^2 with the borderline test |r^2 - thr^2| <= delta (rare slow path: atomic append to a list),
// i.e. the work of the hot loop of k_eval without the fp64 fit in front.  All variants must give the same
// total count.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

constexpr int THREADS = 128;
constexpr int TILE = 512;
constexpr int LIST = 1024;

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__);       \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

struct Hyp {
  float r[9], t[3];
};

// ---------------------------------------------------------------- V0: the scalar loop of round 1
__global__ void __launch_bounds__(THREADS, 6)
k_v0(const Hyp* __restrict__ hyp, const float4* __restrict__ Ya4, const float4* __restrict__ Yb4, int N, float thr2,
     float delta, int* __restrict__ counts, int* __restrict__ nborder) {
  __shared__ float4 sA[TILE], sB[TILE];
  __shared__ uint32_t sList[LIST];
  __shared__ int sListN;
  const int tid = threadIdx.x;
  const int h = blockIdx.x * THREADS + tid;
  if (tid == 0) sListN = 0;
  const Hyp f = hyp[h];
  const float r0 = f.r[0], r1 = f.r[1], r2 = f.r[2], r3 = f.r[3], r4 = f.r[4], r5 = f.r[5], r6 = f.r[6], r7 = f.r[7],
              r8 = f.r[8], t0 = f.t[0], t1 = f.t[1], t2 = f.t[2];
  int cnt = 0;
  for (int base = 0; base < N; base += TILE) {
    const int tn = min(TILE, N - base);
    __syncthreads();
    for (int i = tid; i < tn; i += THREADS) {
      sA[i] = Ya4[base + i];
      sB[i] = Yb4[base + i];
    }
    __syncthreads();
    auto resid2 = [&](int i) {
      const float4 b = sB[i];
      const float4 a = sA[i];
      const float ex = fmaf(r0, b.x, fmaf(r1, b.y, fmaf(r2, b.z, t0))) - a.x;
      const float ey = fmaf(r3, b.x, fmaf(r4, b.y, fmaf(r5, b.z, t1))) - a.y;
      const float ez = fmaf(r6, b.x, fmaf(r7, b.y, fmaf(r8, b.z, t2))) - a.z;
      return fmaf(ex, ex, fmaf(ey, ey, ez * ez));
    };
    auto queue = [&](float q, int i) {
      const int slot = atomicAdd(&sListN, 1);
      if (slot < LIST) sList[slot] = ((uint32_t)tid << 24) | (uint32_t)(base + i);
    };
    int i = 0;
    for (; i + 4 <= tn; i += 4) {
      const float q0 = resid2(i), q1 = resid2(i + 1), q2 = resid2(i + 2), q3 = resid2(i + 3);
      cnt += (q0 < thr2 ? 1 : 0) + (q1 < thr2 ? 1 : 0) + (q2 < thr2 ? 1 : 0) + (q3 < thr2 ? 1 : 0);
      const bool b0 = fabsf(q0 - thr2) <= delta, b1 = fabsf(q1 - thr2) <= delta, b2 = fabsf(q2 - thr2) <= delta,
                 b3 = fabsf(q3 - thr2) <= delta;
      if (b0 | b1 | b2 | b3) {
        if (b0) queue(q0, i);
        if (b1) queue(q1, i + 1);
        if (b2) queue(q2, i + 2);
        if (b3) queue(q3, i + 3);
      }
    }
    for (; i < tn; ++i) {
      const float q = resid2(i);
      cnt += q < thr2 ? 1 : 0;
      if (fabsf(q - thr2) <= delta) queue(q, i);
    }
  }
  __syncthreads();
  counts[h] = cnt;
  if (tid == 0 && sListN) atomicAdd(nborder, sListN);
}

// ---------------------------------------------------------------- packed helpers
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// ---------------------------------------------------------------- V1: thread = 1 hypothesis, f32x2 over 2 matches
// smem SoA: bx[], by[], bz[], ax[], ay[], az[]; one LDS.128 brings one coordinate of 4 matches.
template <int NH>
struct Dummy {};

__global__ void __launch_bounds__(THREADS, 6)
k_v1(const Hyp* __restrict__ hyp, const float4* __restrict__ Ya4, const float4* __restrict__ Yb4, int N, float thr2,
     float delta, int* __restrict__ counts, int* __restrict__ nborder) {
  __shared__ __align__(16) float s[6][TILE];
  __shared__ uint32_t sList[LIST];
  __shared__ int sListN;
  const int tid = threadIdx.x;
  const int h = blockIdx.x * THREADS + tid;
  if (tid == 0) sListN = 0;
  const Hyp f = hyp[h];
  float2 r[9], t[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) r[i] = f2(f.r[i], f.r[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = f2(f.t[i], f.t[i]);
  const float2 nthr = f2(-thr2, -thr2);
  int cnt = 0;
  for (int base = 0; base < N; base += TILE) {
    const int tn = min(TILE, N - base);
    __syncthreads();
    for (int i = tid; i < TILE; i += THREADS) {
      float4 a = make_float4(1e9f, 1e9f, 1e9f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < tn) {
        a = Ya4[base + i];
        b = Yb4[base + i];
      }
      s[0][i] = b.x, s[1][i] = b.y, s[2][i] = b.z, s[3][i] = a.x, s[4][i] = a.y, s[5][i] = a.z;
    }
    __syncthreads();
    const int tn4 = (tn + 3) & ~3;  // padded matches are far away: never inliers, never borderline
    for (int i = 0; i < tn4; i += 4) {
      const float4 bx = *reinterpret_cast<const float4*>(&s[0][i]);
      const float4 by = *reinterpret_cast<const float4*>(&s[1][i]);
      const float4 bz = *reinterpret_cast<const float4*>(&s[2][i]);
      const float4 ax = *reinterpret_cast<const float4*>(&s[3][i]);
      const float4 ay = *reinterpret_cast<const float4*>(&s[4][i]);
      const float4 az = *reinterpret_cast<const float4*>(&s[5][i]);
      float2 d[2];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float2 x = g ? f2(bx.z, bx.w) : f2(bx.x, bx.y);
        const float2 y = g ? f2(by.z, by.w) : f2(by.x, by.y);
        const float2 z = g ? f2(bz.z, bz.w) : f2(bz.x, bz.y);
        const float2 px = g ? f2(ax.z, ax.w) : f2(ax.x, ax.y);
        const float2 py = g ? f2(ay.z, ay.w) : f2(ay.x, ay.y);
        const float2 pz = g ? f2(az.z, az.w) : f2(az.x, az.y);
        const float2 ex = add2(fma2(r[0], x, fma2(r[1], y, fma2(r[2], z, t[0]))), neg2(px));
        const float2 ey = add2(fma2(r[3], x, fma2(r[4], y, fma2(r[5], z, t[1]))), neg2(py));
        const float2 ez = add2(fma2(r[6], x, fma2(r[7], y, fma2(r[8], z, t[2]))), neg2(pz));
        const float2 q = fma2(ex, ex, fma2(ey, ey, mul2(ez, ez)));
        d[g] = add2(q, nthr);
      }
      cnt += (__float_as_uint(d[0].x) >> 31) + (__float_as_uint(d[0].y) >> 31) + (__float_as_uint(d[1].x) >> 31) +
             (__float_as_uint(d[1].y) >> 31);
      const float m = fminf(fminf(fabsf(d[0].x), fabsf(d[0].y)), fminf(fabsf(d[1].x), fabsf(d[1].y)));
      if (m <= delta) {
        const float dd[4] = {d[0].x, d[0].y, d[1].x, d[1].y};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (fabsf(dd[e]) <= delta) {
            const int slot = atomicAdd(&sListN, 1);
            if (slot < LIST) sList[slot] = ((uint32_t)tid << 24) | (uint32_t)(base + i + e);
          }
      }
    }
  }
  __syncthreads();
  counts[h] = cnt;
  if (tid == 0 && sListN) atomicAdd(nborder, sListN);
}

// ---------------------------------------------------------------- V2 / V3: thread = 2*NP hypotheses (NP packed pairs);
// match coordinates DUPLICATED in shared memory so that one LDS.128 yields two packed operands:
//   s0[i] = {bx,bx,by,by}  s1[i] = {bz,bz,ax,ax}  s2[i] = {ay,ay,az,az}
template <int NP, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_v23(const Hyp* __restrict__ hyp, const float4* __restrict__ Ya4, const float4* __restrict__ Yb4, int N, float thr2,
      float delta, int* __restrict__ counts, int* __restrict__ nborder) {
  __shared__ float4 s0[TILE], s1[TILE], s2[TILE];
  __shared__ uint32_t sList[LIST];
  __shared__ int sListN;
  const int tid = threadIdx.x;
  const int h0 = (blockIdx.x * THREADS + tid) * 2 * NP;
  if (tid == 0) sListN = 0;
  float2 r[NP][9], t[NP][3];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const Hyp fa = hyp[h0 + 2 * p], fb = hyp[h0 + 2 * p + 1];
#pragma unroll
    for (int i = 0; i < 9; ++i) r[p][i] = f2(fa.r[i], fb.r[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[p][i] = f2(fa.t[i], fb.t[i]);
  }
  const float2 nthr = f2(-thr2, -thr2);
  int cnt[NP][2];
#pragma unroll
  for (int p = 0; p < NP; ++p) cnt[p][0] = cnt[p][1] = 0;
  for (int base = 0; base < N; base += TILE) {
    const int tn = min(TILE, N - base);
    __syncthreads();
    for (int i = tid; i < TILE; i += THREADS) {
      float4 a = make_float4(1e9f, 1e9f, 1e9f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < tn) {
        a = Ya4[base + i];
        b = Yb4[base + i];
      }
      s0[i] = make_float4(b.x, b.x, b.y, b.y);
      s1[i] = make_float4(b.z, b.z, a.x, a.x);
      s2[i] = make_float4(a.y, a.y, a.z, a.z);
    }
    __syncthreads();
    const int tn2 = (tn + 1) & ~1;
#pragma unroll 2
    for (int i = 0; i < tn2; i += 2) {
      float2 d[2][NP];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float4 u0 = s0[i + e], u1 = s1[i + e], u2 = s2[i + e];
        const float2 x = f2(u0.x, u0.y), y = f2(u0.z, u0.w), z = f2(u1.x, u1.y);
        const float2 px = f2(-u1.z, -u1.w), py = f2(-u2.x, -u2.y), pz = f2(-u2.z, -u2.w);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const float2 ex = add2(fma2(r[p][0], x, fma2(r[p][1], y, fma2(r[p][2], z, t[p][0]))), px);
          const float2 ey = add2(fma2(r[p][3], x, fma2(r[p][4], y, fma2(r[p][5], z, t[p][1]))), py);
          const float2 ez = add2(fma2(r[p][6], x, fma2(r[p][7], y, fma2(r[p][8], z, t[p][2]))), pz);
          const float2 q = fma2(ex, ex, fma2(ey, ey, mul2(ez, ez)));
          d[e][p] = add2(q, nthr);
        }
      }
      float m = 3.0e38f;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        cnt[p][0] += (__float_as_uint(d[0][p].x) >> 31) + (__float_as_uint(d[1][p].x) >> 31);
        cnt[p][1] += (__float_as_uint(d[0][p].y) >> 31) + (__float_as_uint(d[1][p].y) >> 31);
        m = fminf(m, fminf(fminf(fabsf(d[0][p].x), fabsf(d[0][p].y)), fminf(fabsf(d[1][p].x), fabsf(d[1][p].y))));
      }
      if (m <= delta) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            if (fabsf(d[e][p].x) <= delta) {
              const int slot = atomicAdd(&sListN, 1);
              if (slot < LIST) sList[slot] = ((uint32_t)tid << 24) | (uint32_t)(base + i + e);
            }
            if (fabsf(d[e][p].y) <= delta) {
              const int slot = atomicAdd(&sListN, 1);
              if (slot < LIST) sList[slot] = ((uint32_t)tid << 24) | (uint32_t)(base + i + e) | 0x800000u;
            }
          }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    counts[h0 + 2 * p] = cnt[p][0];
    counts[h0 + 2 * p + 1] = cnt[p][1];
  }
  if (tid == 0 && sListN) atomicAdd(nborder, sListN);
}

// ---------------------------------------------------------------- V4: scalar, 2 hypotheses per thread (register tile)
__global__ void __launch_bounds__(THREADS, 4)
k_v4(const Hyp* __restrict__ hyp, const float4* __restrict__ Ya4, const float4* __restrict__ Yb4, int N, float thr2,
     float delta, int* __restrict__ counts, int* __restrict__ nborder) {
  __shared__ float4 sA[TILE], sB[TILE];
  __shared__ uint32_t sList[LIST];
  __shared__ int sListN;
  const int tid = threadIdx.x;
  const int h0 = (blockIdx.x * THREADS + tid) * 2;
  if (tid == 0) sListN = 0;
  const Hyp f = hyp[h0], g = hyp[h0 + 1];
  int c0 = 0, c1 = 0;
  for (int base = 0; base < N; base += TILE) {
    const int tn = min(TILE, N - base);
    __syncthreads();
    for (int i = tid; i < TILE; i += THREADS) {
      float4 a = make_float4(1e9f, 1e9f, 1e9f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < tn) {
        a = Ya4[base + i];
        b = Yb4[base + i];
      }
      sA[i] = a;
      sB[i] = b;
    }
    __syncthreads();
    const int tn2 = (tn + 1) & ~1;
#pragma unroll 2
    for (int i = 0; i < tn2; i += 2) {
      float d[2][2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float4 b = sB[i + e], a = sA[i + e];
        {
          const float ex = fmaf(f.r[0], b.x, fmaf(f.r[1], b.y, fmaf(f.r[2], b.z, f.t[0]))) - a.x;
          const float ey = fmaf(f.r[3], b.x, fmaf(f.r[4], b.y, fmaf(f.r[5], b.z, f.t[1]))) - a.y;
          const float ez = fmaf(f.r[6], b.x, fmaf(f.r[7], b.y, fmaf(f.r[8], b.z, f.t[2]))) - a.z;
          d[e][0] = fmaf(ex, ex, fmaf(ey, ey, ez * ez)) - thr2;
        }
        {
          const float ex = fmaf(g.r[0], b.x, fmaf(g.r[1], b.y, fmaf(g.r[2], b.z, g.t[0]))) - a.x;
          const float ey = fmaf(g.r[3], b.x, fmaf(g.r[4], b.y, fmaf(g.r[5], b.z, g.t[1]))) - a.y;
          const float ez = fmaf(g.r[6], b.x, fmaf(g.r[7], b.y, fmaf(g.r[8], b.z, g.t[2]))) - a.z;
          d[e][1] = fmaf(ex, ex, fmaf(ey, ey, ez * ez)) - thr2;
        }
      }
      c0 += (__float_as_uint(d[0][0]) >> 31) + (__float_as_uint(d[1][0]) >> 31);
      c1 += (__float_as_uint(d[0][1]) >> 31) + (__float_as_uint(d[1][1]) >> 31);
      const float m = fminf(fminf(fabsf(d[0][0]), fabsf(d[0][1])), fminf(fabsf(d[1][0]), fabsf(d[1][1])));
      if (m <= delta) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int p = 0; p < 2; ++p)
            if (fabsf(d[e][p]) <= delta) {
              const int slot = atomicAdd(&sListN, 1);
              if (slot < LIST) sList[slot] = ((uint32_t)tid << 24) | (uint32_t)(base + i + e) | (p ? 0x800000u : 0u);
            }
      }
    }
  }
  __syncthreads();
  counts[h0] = c0;
  counts[h0 + 1] = c1;
  if (tid == 0 && sListN) atomicAdd(nborder, sListN);
}

// ---------------------------------------------------------------- harness
typedef void (*kern_t)(const Hyp*, const float4*, const float4*, int, float, float, int*, int*);

static double run(const char* name, kern_t k, int hyp_per_thread, const Hyp* dh, const float4* dA, const float4* dB, int N,
                  int H, float thr2, float delta, int* dcounts, int* dnb, std::vector<int>& counts) {
  const int blocks = H / (THREADS * hyp_per_thread);
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  CK(cudaMemset(dcounts, 0, sizeof(int) * H));
  CK(cudaMemset(dnb, 0, sizeof(int)));
  for (int w = 0; w < 2; ++w) k<<<blocks, THREADS>>>(dh, dA, dB, N, thr2, delta, dcounts, dnb);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(dnb, 0, sizeof(int)));
  CK(cudaEventRecord(a));
  const int reps = 5;
  for (int r = 0; r < reps; ++r) k<<<blocks, THREADS>>>(dh, dA, dB, N, thr2, delta, dcounts, dnb);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  ms /= reps;
  counts.resize(H);
  CK(cudaMemcpy(counts.data(), dcounts, sizeof(int) * H, cudaMemcpyDeviceToHost));
  int nb;
  CK(cudaMemcpy(&nb, dnb, sizeof(int), cudaMemcpyDeviceToHost));
  long long tot = 0;
  for (int c : counts) tot += c;
  const double evals = (double)H * N;
  printf("%-28s N=%6d H=%8d  %8.3f ms  %7.3f Tevals/s  %6.2f TFLOP/s(27)  frac74.4=%.3f  count=%lld border=%d\n", name, N,
         H, ms, evals / ms * 1e-9, 27.0 * evals / ms * 1e-9, 27.0 * evals / ms * 1e-9 / 74.4, tot, nb / reps);
  return ms;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 20000;
  const int H = argc > 2 ? atoi(argv[2]) : 148 * 6 * 128 * 4;  // multiple of 128*4
  srand(1);
  std::vector<float4> A(N), B(N);
  std::vector<Hyp> hy(H);
  auto rnd = []() { return (float)rand() / RAND_MAX; };
  for (int i = 0; i < N; ++i) {
    B[i] = make_float4(4 * rnd() - 2, 3 * rnd() - 1.5f, 0.8f + 4.2f * rnd(), 0.f);
    const bool inl = rnd() < 0.4f;
    A[i] = inl ? make_float4(B[i].x + 0.02f + 0.002f * rnd(), B[i].y - 0.01f, B[i].z + 0.03f, 0.f)
               : make_float4(4 * rnd() - 2, 3 * rnd() - 1.5f, 0.8f + 4.2f * rnd(), 0.f);
  }
  for (int h = 0; h < H; ++h) {
    const float ax = 0.02f * (rnd() - 0.5f), ay = 0.02f * (rnd() - 0.5f), az = 0.02f * (rnd() - 0.5f);
    Hyp f;
    f.r[0] = 1, f.r[1] = -az, f.r[2] = ay, f.r[3] = az, f.r[4] = 1, f.r[5] = -ax, f.r[6] = -ay, f.r[7] = ax, f.r[8] = 1;
    f.t[0] = 0.02f + 0.01f * (rnd() - 0.5f), f.t[1] = -0.01f + 0.01f * (rnd() - 0.5f), f.t[2] = 0.03f + 0.01f * (rnd() - 0.5f);
    hy[h] = f;
  }
  const float thr = 0.012f, thr2 = thr * thr, delta = 4e-9f;
  Hyp* dh;
  float4 *dA, *dB;
  int *dc, *dnb;
  CK(cudaMalloc(&dh, sizeof(Hyp) * H));
  CK(cudaMalloc(&dA, sizeof(float4) * N));
  CK(cudaMalloc(&dB, sizeof(float4) * N));
  CK(cudaMalloc(&dc, sizeof(int) * H));
  CK(cudaMalloc(&dnb, sizeof(int)));
  CK(cudaMemcpy(dh, hy.data(), sizeof(Hyp) * H, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dA, A.data(), sizeof(float4) * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), sizeof(float4) * N, cudaMemcpyHostToDevice));
  std::vector<int> c0, c;
  run("v0 scalar 1hyp/thr", k_v0, 1, dh, dA, dB, N, H, thr2, delta, dc, dnb, c0);
  auto cmp = [&](const char* n) {
    long long diff = 0;
    for (int i = 0; i < H; ++i) diff += abs(c[i] - c0[i]);
    if (diff) printf("   !! %s differs from v0 in %lld counts (fp32 contraction order differs: only borderline evals may)\n", n, diff);
  };
  run("v1 f32x2 1hyp x 2match", k_v1, 1, dh, dA, dB, N, H, thr2, delta, dc, dnb, c); cmp("v1");
  run("v2 f32x2 2hyp/thr", k_v23<1, 6>, 2, dh, dA, dB, N, H, thr2, delta, dc, dnb, c); cmp("v2");
  run("v3 f32x2 4hyp/thr (minb4)", k_v23<2, 4>, 4, dh, dA, dB, N, H, thr2, delta, dc, dnb, c); cmp("v3");
  run("v3b f32x2 4hyp/thr (minb3)", k_v23<2, 3>, 4, dh, dA, dB, N, H, thr2, delta, dc, dnb, c); cmp("v3b");
  run("v4 scalar 2hyp/thr", k_v4, 2, dh, dA, dB, N, H, thr2, delta, dc, dnb, c); cmp("v4");
  return 0;
}
