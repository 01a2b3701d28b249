// match_tc.cu -- stage 1 on the 5th-generation tensor cores (sm_100a only).
//
// siftmatch (M/sift/siftmatch.c:83-132) is a dense contraction: d2(k1,k2) = |a|^2 + |b|^2 - 2 a.b.
// Three kernels:
//   k_tc_convert   descriptors (class double / single, 128 x K column-major = K rows of 128)
//                  -> fp16 operand image, pre-tiled in the exact shared-memory layout the UMMA
//                  descriptors expect (128-row blocks, two 64-element K halves, 128-byte swizzle),
//                  so that one TMA bulk copy (cp.async.bulk) lands a ready operand tile; plus
//                  |x|^2 per descriptor (fp64 -> fp32), the pair's max |b| and a "values do not
//                  fit fp16" flag.  HBM-bound.
//   k_tc_gemm_pair persistent CTA PAIRS (2-CTA clusters, one per SM pair), warp-specialised: TMA producer
//                  warp, single-thread tcgen05.mma.cta_group::2 issuer in the leader CTA (M256 x N256 x K16,
//                  kind::f16: each CTA holds its own 128 rows of A and HALF of the B tile, so the operand
//                  stream per SM halves; fp32 accumulators in TMEM, two 256-column stages), two epilogue
//                  warpgroups per CTA that pull accumulators with tcgen05.ld and keep a running
//                  (argmax, second largest) of  a.b - |b|^2/2  per L1 row.  The column norm rides in the
//                  contraction itself as a ninth K = 16 step (four fp16 slots holding an exact split of
//                  -|b|^2/2 against a constant A tile), so the epilogue is pure selection: 3.5 instructions
//                  per accumulator.  The distance matrix never exists in memory.
//   k_tc_gemm_top2 the round-1 kernel (cta_group::1, M128 x N128, norms added in the epilogue); kept
//                  as PRE3_TC_V1=1 for A/B timing.
//   k_tc_rescore   the proposal only PROPOSES: the candidate's distance is recomputed exactly in
//                  the reference's arithmetic (sequential `acc += delta*delta` in the class's
//                  accumulation type, siftmatch.c:101-107), the runner-up is bracketed with a
//                  certified error margin, and the float-cast ratio test (:122-123) is decided
//                  only when the bracket decides it; every other row goes to the exact
//                  brute-force kernel (match_exact.cu).  NN indices, scores and accept decisions
//                  are therefore bit-identical to the reference.
// Compile with -fmad=false (the exact recomputation must not contract).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "match.cuh"
#include "tc_ptx.cuh"

namespace pre3 {

namespace tc {

constexpr int ND = 128;            // descriptor length this engine is built for
constexpr int BLK = 128;           // rows per operand block (UMMA M and N)
constexpr int BLK_BYTES = BLK * ND * 2;       // 32 KB: [khalf 2][row 128][128 B]
constexpr int HALF_BYTES = BLK * 128;         // 16 KB
constexpr int A_BUF_BYTES = 2 * BLK_BYTES;    // two row blocks
constexpr int NSTAGE = 3;                     // B ring
constexpr int THREADS = 352;                  // 11 warps: 0-3 / 4-7 epilogue, 8 TMA, 9 MMA, 10 TMEM alloc
// The issuing warps carry the HIGHEST warp ids of their SM sub-partitions (warp id mod 4): the scheduler serves the
// highest ready warp id first, so the single-thread TMA / MMA issuers are not starved by the compute-heavy
// epilogue warps they share a sub-partition with.
constexpr int W_EPI0 = 0, W_TMA = 8, W_MMA = 9, W_ALLOC = 10;
constexpr int SMEM_BYTES = 2 * A_BUF_BYTES + NSTAGE * BLK_BYTES + 256 /*barriers*/ + 4 * BLK * 4 /*norms*/;
// ---- CTA-pair kernel ----
constexpr int EXT_BYTES = BLK * 32;           // 4 KB: the ninth K step of a 128-row block, [row 128][16 halves]
constexpr int P2_NSTAGE = 5;                  // B ring
constexpr int P2_TILE_N = 128;                // columns per B tile (64 per CTA)
constexpr int P2_BHALF = 64 * 128;            // one K half of a CTA's 64 B rows: 8 KB
constexpr int P2_BEXT = 64 * 32;              // their K-extension rows: 2 KB
constexpr int P2_STAGE_BYTES = 2 * P2_BHALF + P2_BEXT;        // 18 KB per stage and CTA
constexpr int P2_OFF_A = 0;                                   // 2 buffers x 2 row blocks x 32 KB
constexpr int P2_OFF_AX = 4 * BLK_BYTES;                      // constant A extension tile (4 KB)
constexpr int P2_OFF_B = P2_OFF_AX + EXT_BYTES;
constexpr int P2_OFF_BAR = P2_OFF_B + P2_NSTAGE * P2_STAGE_BYTES;
constexpr int P2_SMEM_BYTES = P2_OFF_BAR + 512;
static_assert(P2_SMEM_BYTES <= 227 * 1024, "CTA-pair matcher: shared memory");
static_assert(P2_STAGE_BYTES % 1024 == 0, "B stages must keep the 1024-byte swizzle alignment");
// exact split of g = -|b|^2/2 over four fp16 slots:  g = 4096 h1 + 4096 h2 + h3 + h4  (A side: 4096, 4096, 1, 1)
constexpr float EXT_SCALE = 4096.0f;
constexpr double NORM_MAX = 2.0e8;            // |x|^2 above this does not fit the split -> pair flagged bad (exact kernel)

// The epilogue orders  v'(j) = |b_j|^2 + C - 2 a~.b~_j  (C = 1.0625 max_i |a_i|^2 of the pair keeps v' > 0, so
// float bits order like unsigned integers); the low 7 bits of the key carry the column within the tile.
// byte offset of 16-byte chunk `chunk` (0: halves 0-7, 1: halves 8-15) of row r inside a 128-row K-extension tile.
//   layout 0  SWIZZLE_32B: 32-byte rows, the two chunks XOR-swizzled with bit 2 of the row (address bit 7)
//   layout 1  no swizzle ("interleaved"): 8 x 16-byte core matrices, [row group][chunk][row & 7]
__host__ __device__ __forceinline__ int ext_off(int layout, int r, int chunk) {
  return layout == 0 ? r * 32 + ((chunk ^ ((r >> 2) & 1)) << 4) : (r >> 3) * 256 + chunk * 128 + (r & 7) * 16;
}

struct Prop {      // per L1 row, output of the proposal GEMM
  float best;      // smallest v' (low 7 mantissa bits truncated)
  float second;    // second smallest v' (+inf if < 2 columns)
  int32_t idx;     // argmin column (-1 if none)
};

struct PairInfo {
  unsigned amax_bits;  // max |a|^2 (float bits) over the valid rows
  unsigned bmax_bits;  // max |b|^2 (float bits) over the valid columns
  int bad;             // some value is non-finite or too large for fp16
  int pad;
};

// kind::f16, A = B = F16, D = F32, both K-major, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BLK >> 3) << 17) | ((uint32_t)(BLK >> 4) << 24);

// ---------------------------------------------------------------------------------------------
// k_tc_convert: one warp per descriptor row
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load4(const double* src, double* v) {  // 32-byte aligned by construction
  const double2 x = __ldg(reinterpret_cast<const double2*>(src));
  const double2 y = __ldg(reinterpret_cast<const double2*>(src) + 1);
  v[0] = x.x, v[1] = x.y, v[2] = y.x, v[3] = y.y;
}
__device__ __forceinline__ void load4(const float* src, double* v) {
  const float4 x = __ldg(reinterpret_cast<const float4*>(src));
  v[0] = x.x, v[1] = x.y, v[2] = x.z, v[3] = x.w;
}

constexpr int CV_ROWS = 32;  // rows per block (4 per warp)

struct FrameInfo {       // per converted descriptor set ("frame")
  unsigned max_bits;     // max |x|^2 (float bits) over the valid rows
  int bad;               // some value is non-finite or too large for fp16
};

// One launch converts `gridDim.y` descriptor sets.  Norms are stored RAW (|x|^2 rounded up to float, +inf for
// padded rows, which therefore can never win as columns); the bias that keeps the keys positive is added by the
// GEMM epilogue, so one converted set can serve as the A operand of one pair and the B operand of another
// (consecutive frames of a sequence).
template <typename T>
__global__ void __launch_bounds__(256)
k_tc_convert(const T* __restrict__ L, int K, int Kp, const int32_t* __restrict__ kc, unsigned char* __restrict__ img,
             unsigned char* __restrict__ ext, int ext_layout, float* __restrict__ nrm, FrameInfo* __restrict__ finfo) {
  __shared__ float s_max[8];
  __shared__ int s_bad[8];
  const int p = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = kc ? max(0, min(kc[p], K)) : K;
  float wmax = 0.f;
  bool wbad = false;
  double v[CV_ROWS / 8][4];
#pragma unroll
  for (int i = 0; i < CV_ROWS / 8; ++i) {  // all loads first
    const int row = blockIdx.x * CV_ROWS + warp * (CV_ROWS / 8) + i;
    v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.0;
    if (row < n) load4(L + ((size_t)p * K + row) * ND + 4 * lane, v[i]);
  }
#pragma unroll
  for (int i = 0; i < CV_ROWS / 8; ++i) {
    const int row = blockIdx.x * CV_ROWS + warp * (CV_ROWS / 8) + i;
    if (row >= Kp) break;
    // fp16 image: block rb, K half kh, row r, 16-byte chunk c of the 128-byte row XOR-swizzled with r & 7
    const int rb = row >> 7, r = row & 127, kh = lane >> 4, c = (lane & 15) >> 1;
    __half2 h01 = __floats2half2_rn((float)v[i][0], (float)v[i][1]);
    __half2 h23 = __floats2half2_rn((float)v[i][2], (float)v[i][3]);
    uint2 packed;
    packed.x = *reinterpret_cast<unsigned*>(&h01);
    packed.y = *reinterpret_cast<unsigned*>(&h23);
    unsigned char* dst = img + (size_t)p * Kp * (ND * 2) + (size_t)rb * BLK_BYTES + (size_t)kh * HALF_BYTES +
                         (size_t)r * 128 + (size_t)((c ^ (r & 7)) << 4) + (size_t)((lane & 1) << 3);
    *reinterpret_cast<uint2*>(dst) = packed;
    // |x|^2 in fp64 (fixed shuffle tree), range check
    double sq = (v[i][0] * v[i][0] + v[i][1] * v[i][1]) + (v[i][2] * v[i][2] + v[i][3] * v[i][3]);
    bool bad = false;
#pragma unroll
    for (int e = 0; e < 4; ++e) bad = bad || !(fabs(v[i][e]) <= 60000.0);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    bad = __any_sync(0xffffffffu, bad) || !(sq <= NORM_MAX);
    if (ext && lane == 0) {
      // ninth K step of the CTA-pair GEMM: g = -|x|^2/2 = 4096 h1 + 4096 h2 + h3 + h4 (exact in fp64 down to the fp16
      // denormal floor 2^-25); padded rows get a value no real column can reach.  16 halves per row (ext_off).
      __half h[4];
      if (row < n && !bad) {
        double g = -0.5 * sq;
        h[0] = __double2half(g * (1.0 / 4096.0));
        g = g - 4096.0 * (double)__half2float(h[0]);
        h[1] = __double2half(g * (1.0 / 4096.0));
        g = g - 4096.0 * (double)__half2float(h[1]);
        h[2] = __double2half(g);
        g = g - (double)__half2float(h[2]);
        h[3] = __double2half(g);
      } else {
        h[0] = __float2half(-60000.0f);
        h[1] = h[2] = h[3] = __float2half(0.0f);
      }
      uint4 w;
      w.x = (unsigned)__half_as_ushort(h[0]) | ((unsigned)__half_as_ushort(h[1]) << 16);
      w.y = (unsigned)__half_as_ushort(h[2]) | ((unsigned)__half_as_ushort(h[3]) << 16);
      w.z = w.w = 0u;
      unsigned char* e0 = ext + ((size_t)p * Kp + (size_t)(row & ~127)) * 32;  // the row's 4 KB tile
      *reinterpret_cast<uint4*>(e0 + ext_off(ext_layout, row & 127, 0)) = w;
      *reinterpret_cast<uint4*>(e0 + ext_off(ext_layout, row & 127, 1)) = make_uint4(0u, 0u, 0u, 0u);
    }
    float f = __double2float_ru(sq);
    if (row < n) {
      wmax = fmaxf(wmax, f);
      wbad = wbad || bad;
    }
    if (lane == 0) nrm[(size_t)p * Kp + row] = (row < n) ? f : INFINITY;
  }
  if (lane == 0) {
    s_max[warp] = wmax;
    s_bad[warp] = wbad ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = 0.f;
    int bd = 0;
    for (int w = 0; w < 8; ++w) {
      mx = fmaxf(mx, s_max[w]);
      bd |= s_bad[w];
    }
    if (mx > 0.f) atomicMax(&finfo[p].max_bits, __float_as_uint(mx));
    if (bd) atomicOr(&finfo[p].bad, 1);
  }
}

// per pair: bounds of its two descriptor sets (fa[p], fb[p])
__global__ void k_tc_pair_info(int P, const FrameInfo* __restrict__ fa, const FrameInfo* __restrict__ fb,
                               PairInfo* __restrict__ info, int a_shared) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int pa = a_shared ? 0 : p;
  PairInfo o;
  o.amax_bits = fa[pa].max_bits;
  o.bmax_bits = fb[p].max_bits;
  o.bad = fa[pa].bad | fb[p].bad;
  o.pad = 0;
  info[p] = o;
}

// ---------------------------------------------------------------------------------------------
// k_tc_gemm_top2
// ---------------------------------------------------------------------------------------------
struct Barriers {
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[NSTAGE], b_empty[NSTAGE];
  uint64_t t_full[4], t_empty[4];  // accumulator slot = set * 2 + row block
  uint32_t tmem_base;
};

// EXP selects an ablation of the pipeline (PRE3_TC_EXP in the environment, tools/match_bench.py; 0 = the product;
// the others produce no usable proposal and only exist to time the halves of the kernel):
//   1  TMA + MMA only: the epilogue waits for and releases the accumulators without reading them
//   2  TMA + epilogue only: the issuer commits without issuing tcgen05.mma
//   4  TMA only          5  MMA only (no B loads, no epilogue)
//   6  MMA only, N = 256 per instruction (half as many)      7  MMA only, A operand read from tensor memory
template <int EXP>
__global__ void __launch_bounds__(THREADS, 1)
k_tc_gemm_top2(const unsigned char* __restrict__ imgA, const unsigned char* __restrict__ imgB,
               const float* __restrict__ nrmB, const PairInfo* __restrict__ pinfo, int P, int K1p, int K2p,
               Prop* __restrict__ prop, int a_shared) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  if (base & 1023u) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t sA = base;                                  // 2 buffers x 64 KB
  const uint32_t sB = base + 2 * A_BUF_BYTES;                // NSTAGE x 32 KB
  Barriers* bars = reinterpret_cast<Barriers*>(smem + 2 * A_BUF_BYTES + NSTAGE * BLK_BYTES);
  // [row block][set][128] biased column norms of the tile in flight
  float* sNB = reinterpret_cast<float*>(smem + 2 * A_BUF_BYTES + NSTAGE * BLK_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = (K1p / BLK + 1) / 2;  // row groups (<= 2 row blocks each) per pair
  const int ntile = K2p / BLK;             // B tiles per pair
  const long long nunits = (long long)P * groups;

  if (warp == W_MMA && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->a_full[i]), 1);
      mbar_init(smem_u32(&bars->a_empty[i]), 1);
    }
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(smem_u32(&bars->b_full[i]), 1);
      mbar_init(smem_u32(&bars->b_empty[i]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&bars->t_full[i]), 1);
      mbar_init(smem_u32(&bars->t_empty[i]), 128);  // every thread of the epilogue warpgroup arrives
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == W_TMA) {
    // ===== TMA producer ========================================================================
    if (lane == 0) {
      long long t = 0;  // B tiles issued so far
      int uc = 0;       // units handled by this CTA so far
      for (long long u = blockIdx.x; u < nunits; u += gridDim.x, ++uc) {
        const int p = (int)(u / groups), g = (int)(u % groups);
        const int nrb = min(2, K1p / BLK - 2 * g);
        const int ab = uc & 1;
        mbar_wait(smem_u32(&bars->a_empty[ab]), ((uc >> 1) & 1) ^ 1);
        mbar_expect_tx(smem_u32(&bars->a_full[ab]), (uint32_t)(nrb * BLK_BYTES));
        tma_bulk_g2s(sA + ab * A_BUF_BYTES, imgA + ((size_t)(a_shared ? 0 : p) * K1p + (size_t)g * 2 * BLK) * (ND * 2),
                     (uint32_t)(nrb * BLK_BYTES), smem_u32(&bars->a_full[ab]));
        for (int j = 0; j < ntile; ++j, ++t) {
          if (EXP >= 5) continue;
          const int st = (int)(t % NSTAGE);
          mbar_wait(smem_u32(&bars->b_empty[st]), (uint32_t)(((t / NSTAGE) & 1) ^ 1));
          mbar_expect_tx(smem_u32(&bars->b_full[st]), BLK_BYTES);
          tma_bulk_g2s(sB + st * BLK_BYTES, imgB + ((size_t)p * K2p + (size_t)j * BLK) * (ND * 2), BLK_BYTES,
                       smem_u32(&bars->b_full[st]));
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer (one thread) =============================================================
    if (lane == 0) {
      long long t = 0;
      int uc = 0;
      uint32_t use0 = 0, use1 = 0, use2 = 0, use3 = 0;  // times each accumulator slot has been filled
      for (long long u = blockIdx.x; u < nunits; u += gridDim.x, ++uc) {
        const int g = (int)(u % groups);
        const int nrb = min(2, K1p / BLK - 2 * g);
        const int ab = uc & 1;
        mbar_wait(smem_u32(&bars->a_full[ab]), (uc >> 1) & 1);
        for (int j = 0; j < ntile; ++j, ++t) {
          const int st = (int)(t % NSTAGE);
          const int set = (int)(t & 1);
          if (EXP < 5) mbar_wait(smem_u32(&bars->b_full[st]), (uint32_t)((t / NSTAGE) & 1));
          tc_fence_after();
          for (int rb = 0; rb < nrb; ++rb) {
            const int slot = set * 2 + rb;
            uint32_t& use = slot == 0 ? use0 : (slot == 1 ? use1 : (slot == 2 ? use2 : use3));
            mbar_wait(smem_u32(&bars->t_empty[slot]), (use & 1) ^ 1);
            ++use;
            tc_fence_after();
            const uint32_t d = tmem + (uint32_t)(slot * BLK);
#pragma unroll
            for (int k = 0; k < ND / 16; ++k) {
              const uint32_t koff = (uint32_t)((k >> 2) * HALF_BYTES + (k & 3) * 32);
              const uint64_t da = umma_desc(sA + ab * A_BUF_BYTES + rb * BLK_BYTES + koff);
              const uint64_t db = umma_desc(sB + st * BLK_BYTES + koff);
              if (EXP == 6) {  // timing only: one N = 256 instruction per two tiles (operands are garbage)
                if ((t & 1) == 0)
                  tc_mma_f16(tmem + (uint32_t)(rb * 2 * BLK), da, umma_desc(sB + koff),
                             (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(BLK >> 4) << 24), k > 0 ? 1u : 0u);
              } else if (EXP == 7) {
                tc_mma_f16_ts(d, tmem + (uint32_t)(448 + k * 8), db, IDESC, k > 0 ? 1u : 0u);
              } else if (EXP != 2 && EXP != 4) {
                tc_mma_f16(d, da, db, IDESC, k > 0 ? 1u : 0u);
              }
            }
            tc_commit(smem_u32(&bars->t_full[slot]));  // accumulator ready for the epilogue
          }
          if (EXP < 5) tc_commit(smem_u32(&bars->b_empty[st]));  // B stage may be refilled
        }
        tc_commit(smem_u32(&bars->a_empty[ab]));  // A buffer may be refilled
      }
    }
  } else if (warp < W_TMA) {
    // ===== epilogue: warpgroup rb owns row block rb ============================================
    const int rb = (warp - W_EPI0) >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    long long t = 0;
    uint32_t useA = 0, useB = 0;  // times this warpgroup has drained its slot of set 0 / set 1
    for (long long u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int p = (int)(u / groups), g = (int)(u % groups);
      const int nrb = min(2, K1p / BLK - 2 * g);
      if (rb >= nrb) {
        t += ntile;
        continue;
      }
      uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;  // two smallest packed keys of this row
      uint32_t keymask;
      asm volatile("mov.u32 %0, 0xFFFFFF80;" : "=r"(keymask));
      int btile = -1;
      const float* nb = nrmB + (size_t)p * K2p;
      const int wtid = threadIdx.x - (W_EPI0 + 4 * rb) * 32;  // 0..127 within the warpgroup
      // the column norm of the NEXT tile is fetched one tile ahead: its global-load latency (exposed, it cost a
      // quarter of the epilogue's time) hides behind the current tile's selection
      float nb_next = __ldg(nb + wtid);
      // C = 1.0625 max_i |a_i|^2 of the pair: keeps every v' = |b|^2 + C - 2 a.b positive, so float bits order
      // like unsigned integers (k_tc_rescore subtracts the same C)
      const float bias = 1.0625f * __uint_as_float(pinfo[p].amax_bits);
      for (int j = 0; j < ntile; ++j, ++t) {
        const int set = (int)(t & 1);
        const int slot = set * 2 + rb;
        // stage the tile's (biased) column norms for broadcast reads
        float* snb = sNB + (rb * 2 + set) * BLK;
        snb[wtid] = nb_next + bias;
        if (j + 1 < ntile) nb_next = __ldg(nb + (size_t)(j + 1) * BLK + wtid);
        named_bar_sync(1 + rb, 128);
        uint32_t& use = set == 0 ? useA : useB;
        mbar_wait(smem_u32(&bars->t_full[slot]), use & 1);
        ++use;
        tc_fence_after();
        const uint32_t m1_in = m1;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * BLK);
        uint32_t buf[2][32];
        if (EXP != 1 && EXP < 4) tc_ld_32x32(taddr, buf[0]);
#pragma unroll
        for (int c = 0; c < BLK / 32; ++c) {
          if (EXP == 1 || EXP >= 4) break;
          tc_ld_wait(buf[c & 1]);
          if (c + 1 < BLK / 32) tc_ld_32x32(taddr + (uint32_t)((c + 1) * 32), buf[(c + 1) & 1]);
          const float4* nb4 = reinterpret_cast<const float4*>(snb + c * 32);
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 n4 = nb4[i4];
            const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float v = fmaf(-2.0f, __uint_as_float(buf[c & 1][4 * i4 + e]), nn[e]);
              uint32_t key;  // (bits & ~127) | column: one LOP3 (mask kept in a register)
              asm("lop3.b32 %0, %1, %2, %3, 0xEA;"
                  : "=r"(key)
                  : "r"(__float_as_uint(v)), "r"(keymask), "r"((uint32_t)(c * 32 + 4 * i4 + e)));
              m2 = min(m2, max(key, m1));
              m1 = min(m1, key);
            }
          }
        }
        if (m1 != m1_in) btile = j;
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->t_empty[slot]));
      }
      const int row = (g * 2 + rb) * BLK + q * 32 + lane;
      Prop o;
      o.best = __uint_as_float(m1 & 0xFFFFFF80u);
      o.second = __uint_as_float(m2 & 0xFFFFFF80u);
      o.idx = btile < 0 ? -1 : btile * BLK + (int)(m1 & 0x7Fu);
      prop[(size_t)p * K1p + row] = o;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// k_tc_gemm_pair: the CTA-pair form (cta_group::2).  Cluster = 2 CTAs on one SM pair; unit of work = (pair p,
// 512-row group g of L1): CTA r holds TWO 128-row A tiles (s = 0, 1: row blocks 4g + 2s + r), resident for the whole
// unit.  Per B tile of 128 columns CTA r loads the 64 columns 128 j + 64 r (+ their K-extension rows); for each s the
// leader issues nine M256 x N128 x K16 MMAs that read A and B from BOTH CTAs' shared memory and write a 128 x 128 fp32
// accumulator into EACH CTA's tensor memory (slot = 2 stage + s; four 128-column slots).  Every operand byte crosses
// L2 -> SM once per 512 rows of L1 (measured: the 256-row form was bound by that stream, not by the MMAs or the
// selection: 0.227 ms at the 256 x 2048 x 2048 shape with the MMAs and the selection switched off).
//   barriers (per CTA unless noted):
//     a_full / b_full    TMA bytes of this CTA landed (expect_tx)
//     pa_full / pb_full  LEADER only: the peer's relay thread saw the peer's a_full / b_full (remote arrive)
//     a_empty / b_empty  tcgen05.commit multicast to both CTAs: the MMAs reading the buffer have completed
//     t_full[slot]       tcgen05.commit multicast: accumulator slot ready in both CTAs' tensor memory
//     t_empty[slot]      LEADER only: 8 warp arrivals (4 local, 4 remote): both CTAs' warpgroup s drained the slot
// ---------------------------------------------------------------------------------------------
struct PairBarriers {
  uint64_t a_full[2], a_empty[2], pa_full[2];
  uint64_t b_full[P2_NSTAGE], b_empty[P2_NSTAGE], pb_full[P2_NSTAGE];
  uint64_t t_full[4], t_empty[4];
  uint32_t tmem_base;
};
static_assert(sizeof(PairBarriers) <= 512, "PairBarriers");

__device__ __forceinline__ void tc_mma_f16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major K = 16 operand tile of the ninth step.  layout 0: SWIZZLE_32B (32-byte rows, 8-row groups 256 B apart);
// layout 1: no swizzle, core matrices 128 B apart along K (LBO) and 256 B apart along the rows (SBO)
__device__ __forceinline__ uint64_t umma_desc_ext(uint32_t saddr, int layout) {
  const uint64_t common = (uint64_t)((saddr >> 4) & 0x3FFFu) | (16ull << 32) | (1ull << 46);
  return layout == 0 ? (common | (1ull << 16) | (6ull << 61)) : (common | (8ull << 16));
}
// kind::f16, A = B = F16, D = F32, both K-major, M = 256 (pair), N = 128
constexpr uint32_t IDESC_PAIR = (1u << 4) | ((uint32_t)(P2_TILE_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

struct Prop2 {     // per L1 row: s(j) = a~.b~_j - |b_j|^2/2  (d2(j) = |a|^2 - 2 s(j)); larger = closer
  float best;      // largest s (low 7 mantissa bits replaced)
  float second;    // second largest s (-inf if < 2 columns)
  int32_t idx;     // argmax column (-1 if none)
};

// EXP: 0 product; 1 TMA + MMA only (accumulators released unread); 2 TMA + epilogue only (no MMA issued);
//      3 TMA + tcgen05.ld only (no MMA, accumulators read into registers but not looked at)
// EPI: 1 product (selection split over the ALU and FMA pipes); 0 all-ALU selection (3.5 instructions per accumulator)
template <int EXP, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_tc_gemm_pair(const unsigned char* __restrict__ imgA, const unsigned char* __restrict__ imgB,
               const unsigned char* __restrict__ extB, int ext_layout, int P, int K1p, int K2p,
               Prop2* __restrict__ prop, int a_shared) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t sA = base + P2_OFF_A, sAX = base + P2_OFF_AX, sB = base + P2_OFF_B;
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(smem + P2_OFF_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int nblk = K1p / BLK;               // K1p, K2p are multiples of 256
  const int groups = (nblk + 3) / 4;        // 512-row groups; the last one may hold only two row blocks (s = 0)
  const int ntile = K2p / P2_TILE_N;
  const long long nunits = (long long)P * groups;

  if (warp == W_MMA && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->a_full[i]), 1);
      mbar_init(smem_u32(&bars->a_empty[i]), 1);
      mbar_init(smem_u32(&bars->pa_full[i]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&bars->t_full[i]), 1);
      mbar_init(smem_u32(&bars->t_empty[i]), 8);
    }
    for (int i = 0; i < P2_NSTAGE; ++i) {
      mbar_init(smem_u32(&bars->b_full[i]), 1);
      mbar_init(smem_u32(&bars->b_empty[i]), 1);
      mbar_init(smem_u32(&bars->pb_full[i]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // constant A tile of the ninth K step: every row = (4096, 4096, 1, 1, 0 ...)
  for (int r = threadIdx.x; r < BLK; r += THREADS) {
    const __half one = __float2half(1.0f), sc = __float2half(EXT_SCALE);
    uint4 w;
    w.x = (unsigned)__half_as_ushort(sc) | ((unsigned)__half_as_ushort(sc) << 16);
    w.y = (unsigned)__half_as_ushort(one) | ((unsigned)__half_as_ushort(one) << 16);
    w.z = w.w = 0u;
    *reinterpret_cast<uint4*>(smem + P2_OFF_AX + ext_off(ext_layout, r, 0)) = w;
    *reinterpret_cast<uint4*>(smem + P2_OFF_AX + ext_off(ext_layout, r, 1)) = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == W_TMA) {
    // ===== TMA producer (each CTA loads its own operand parts) =====================================
    if (lane == 0) {
      long long t = 0;
      int uc = 0;
      for (long long u = cluster_id; u < nunits; u += nclusters, ++uc) {
        const int p = (int)(u / groups), g = (int)(u % groups);
        const int ns = min(2, (nblk - 4 * g) / 2);  // A tiles per CTA in this unit
        const int ab = uc & 1;
        mbar_wait(smem_u32(&bars->a_empty[ab]), ((uc >> 1) & 1) ^ 1);
        mbar_expect_tx(smem_u32(&bars->a_full[ab]), (uint32_t)(ns * BLK_BYTES));
        for (int sidx = 0; sidx < ns; ++sidx)
          tma_bulk_g2s(sA + (ab * 2 + sidx) * BLK_BYTES,
                       imgA + ((size_t)(a_shared ? 0 : p) * K1p + (size_t)(4 * g + 2 * sidx + (int)rank) * BLK) * (ND * 2),
                       (uint32_t)BLK_BYTES, smem_u32(&bars->a_full[ab]));
        for (int j = 0; j < ntile; ++j, ++t) {
          const int st = (int)(t % P2_NSTAGE);
          mbar_wait(smem_u32(&bars->b_empty[st]), (uint32_t)(((t / P2_NSTAGE) & 1) ^ 1));
          mbar_expect_tx(smem_u32(&bars->b_full[st]), (uint32_t)P2_STAGE_BYTES);
          // this CTA's 64 columns: rows 64 rank .. 64 rank + 63 of the 128-row block j of the B image
          const unsigned char* blk = imgB + ((size_t)p * K2p + (size_t)j * BLK) * (ND * 2) + (size_t)rank * P2_BHALF;
          const uint32_t dst = sB + st * P2_STAGE_BYTES;
          tma_bulk_g2s(dst, blk, P2_BHALF, smem_u32(&bars->b_full[st]));                          // K half 0
          tma_bulk_g2s(dst + P2_BHALF, blk + HALF_BYTES, P2_BHALF, smem_u32(&bars->b_full[st]));  // K half 1
          tma_bulk_g2s(dst + 2 * P2_BHALF, extB + ((size_t)p * K2p + (size_t)j * BLK) * 32 + (size_t)rank * P2_BEXT,
                       P2_BEXT, smem_u32(&bars->b_full[st]));
        }
      }
      // tail: the multicast commits that release the last buffers target THIS CTA's barriers too; stay until they
      // have arrived (a commit must never land in the shared memory of a CTA that has exited)
      for (int i = 0; i < P2_NSTAGE; ++i, ++t)
        mbar_wait(smem_u32(&bars->b_empty[t % P2_NSTAGE]), (uint32_t)(((t / P2_NSTAGE) & 1) ^ 1));
      for (int i = 0; i < 2; ++i, ++uc) mbar_wait(smem_u32(&bars->a_empty[uc & 1]), ((uc >> 1) & 1) ^ 1);
    }
  } else if (warp == W_MMA) {
    if (lane == 0) {
      long long t = 0;
      int uc = 0;
      if (leader) {
        // ===== MMA issuer (one thread of the leader CTA, for both SMs) ===============================
        uint32_t use00 = 0, use01 = 0, use10 = 0, use11 = 0;  // fills of slot (stage, s)
        for (long long u = cluster_id; u < nunits; u += nclusters, ++uc) {
          const int g = (int)(u % groups);
          const int ns = min(2, (nblk - 4 * g) / 2);
          const int ab = uc & 1;
          mbar_wait(smem_u32(&bars->a_full[ab]), (uc >> 1) & 1);
          mbar_wait_cl(smem_u32(&bars->pa_full[ab]), (uc >> 1) & 1);
          for (int j = 0; j < ntile; ++j, ++t) {
            const int st = (int)(t % P2_NSTAGE);
            const uint32_t ph = (uint32_t)((t / P2_NSTAGE) & 1);
            mbar_wait(smem_u32(&bars->b_full[st]), ph);
            mbar_wait_cl(smem_u32(&bars->pb_full[st]), ph);
            tc_fence_after();
            const uint32_t sBst = sB + st * P2_STAGE_BYTES;
#pragma unroll
            for (int sidx = 0; sidx < 2; ++sidx) {
              if (sidx >= ns) break;
              const int slot = (int)(t & 1) * 2 + sidx;
              uint32_t& use = (t & 1) ? (sidx ? use11 : use10) : (sidx ? use01 : use00);
              mbar_wait_cl(smem_u32(&bars->t_empty[slot]), (use & 1) ^ 1);
              ++use;
              tc_fence_after();
              const uint32_t d = tmem + (uint32_t)(slot * P2_TILE_N);
              if (EXP != 2 && EXP != 3 && EXP != 4) {
#pragma unroll
                for (int k = 0; k < ND / 16; ++k) {
                  const uint32_t koa = (uint32_t)((k >> 2) * HALF_BYTES + (k & 3) * 32);
                  const uint32_t kob = (uint32_t)((k >> 2) * P2_BHALF + (k & 3) * 32);
                  tc_mma_f16_2cta(d, umma_desc(sA + (ab * 2 + sidx) * BLK_BYTES + koa), umma_desc(sBst + kob), IDESC_PAIR,
                                  k > 0 ? 1u : 0u);
                }
                tc_mma_f16_2cta(d, umma_desc_ext(sAX, ext_layout), umma_desc_ext(sBst + 2 * P2_BHALF, ext_layout),
                                IDESC_PAIR, 1u);
              }
              tc_commit_mc2(smem_u32(&bars->t_full[slot]));
            }
            tc_commit_mc2(smem_u32(&bars->b_empty[st]));
          }
          tc_commit_mc2(smem_u32(&bars->a_empty[ab]));
        }
      } else {
        // ===== relay (peer CTA): tells the leader's issuer that this CTA's operand bytes have landed ==
        for (long long u = cluster_id; u < nunits; u += nclusters, ++uc) {
          const int ab = uc & 1;
          mbar_wait(smem_u32(&bars->a_full[ab]), (uc >> 1) & 1);
          mbar_arrive_cluster(mapa_u32(smem_u32(&bars->pa_full[ab]), 0));
          for (int j = 0; j < ntile; ++j, ++t) {
            const int st = (int)(t % P2_NSTAGE);
            mbar_wait(smem_u32(&bars->b_full[st]), (uint32_t)((t / P2_NSTAGE) & 1));
            mbar_arrive_cluster(mapa_u32(smem_u32(&bars->pb_full[st]), 0));
          }
        }
      }
    }
  } else if (warp < W_TMA) {
    // ===== epilogue: warpgroup s owns A tile s of this CTA (row block 4g + 2s + rank) ================
    // The selection runs on the ALU pipe (LOP3 and the 3-input FMNMX3 issue at half rate, the 2-input FMNMX at full
    // rate: tools/pipebench.cu); what does not have to be exact is moved to the FMA pipe: per two accumulators the ALU
    // packs the column into the keys (2 LOP3), takes their maximum (carries the index exactly), advances the running
    // maximum and the running second (one 3-input max); the pair's minimum and min(m1, hi) -- which only feed the
    // SECOND best, a value that is bracketed with a margin anyway -- are formed arithmetically, lo = (k0 + k1) - hi and
    // x = (m1 + hi) - max(m1, hi), as packed FADD2 on two independent chains (even / odd accumulators), 2 ulp off at
    // most.  The next tile's first tcgen05.ld is issued, and the accumulator slot handed back to the MMA issuer, as
    // soon as the current tile's last load has landed in registers.
    const int sidx = (warp - W_EPI0) >> 2;
    const int q = warp & 3;
    const uint32_t tel0 = mapa_u32(smem_u32(&bars->t_empty[sidx]), 0);      // the leader's t_empty of (stage 0, s)
    const uint32_t tel1 = mapa_u32(smem_u32(&bars->t_empty[2 + sidx]), 0);  // (stage 1, s)
    long long t = 0;
    uint32_t use0 = 0, use1 = 0;  // times this warpgroup has waited for its slot of stage 0 / 1
    uint32_t keymask;
    asm volatile("mov.u32 %0, 0xFFFFFF80;" : "=r"(keymask));
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sidx * P2_TILE_N);
    uint32_t buf[2][32];
    auto wait_full = [&](int stg) {
      uint32_t& use = stg == 0 ? use0 : use1;
      mbar_wait(smem_u32(&bars->t_full[stg * 2 + sidx]), use & 1);
      ++use;
      tc_fence_after();
    };
    auto release = [&](int stg) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(smem_u32(&bars->t_empty[stg * 2 + sidx]));
        else mbar_arrive_cluster(stg == 0 ? tel0 : tel1);
      }
    };
    bool primed = false;
    // units in which this warpgroup has a tile: all but a trailing half group for s = 1
    for (long long u = cluster_id; u < nunits; u += nclusters) {
      const int p = (int)(u / groups), g = (int)(u % groups);
      const int ns = min(2, (nblk - 4 * g) / 2);
      if (sidx >= ns) {  // no A tile for this warpgroup: the issuer skips the slot as well
        t += ntile;
        continue;
      }
      // does the next unit this warpgroup takes part in exist?  (needed for the prefetch of its first tile)
      bool next_unit = false;
      for (long long v = u + nclusters; v < nunits; v += nclusters)
        if (sidx < min(2, (nblk - 4 * (int)(v % groups)) / 2)) {
          next_unit = true;
          break;
        }
      float2 m1 = make_float2(-INFINITY, -INFINITY), m2 = make_float2(-INFINITY, -INFINITY);  // chains: even / odd columns
      int btA = -1, btB = -1;
      for (int j = 0; j < ntile; ++j, ++t) {
        const int stg = (int)(t & 1);
        if (!primed) {
          wait_full(stg);
          if (EXP != 1 && EXP != 4) tc_ld_32x32(tbase + (uint32_t)(stg * 2 * P2_TILE_N), buf[0]);
          primed = true;
        }
        const float2 m1_in = m1;
        const bool has_next = (j + 1 < ntile) || next_unit;
#pragma unroll
        for (int c = 0; c < P2_TILE_N / 32; ++c) {
          if (EXP != 1 && EXP != 4) tc_ld_wait(buf[c & 1]);
          if (c + 1 < P2_TILE_N / 32) {
            if (EXP != 1 && EXP != 4) tc_ld_32x32(tbase + (uint32_t)(stg * 2 * P2_TILE_N + (c + 1) * 32), buf[(c + 1) & 1]);
          } else {
            release(stg);  // every accumulator of this slot is in registers
            if (has_next) {
              // the next tile of this warpgroup: stage t + 1 when it is in this unit; in the next unit the tile counter
              // has skipped the units without an s = 1 tile, so its stage is that of the unit's first tile
              int nstg = stg ^ 1;
              if (j + 1 >= ntile) {
                long long tt = t + 1;
                for (long long v = u + nclusters; v < nunits; v += nclusters) {
                  if (sidx < min(2, (nblk - 4 * (int)(v % groups)) / 2)) break;
                  tt += ntile;
                }
                nstg = (int)(tt & 1);
              }
              wait_full(nstg);
              if (EXP != 1 && EXP != 4) tc_ld_32x32(tbase + (uint32_t)(nstg * 2 * P2_TILE_N), buf[0]);
            }
          }
          if (EXP == 1 || EXP == 4) continue;
          if (EXP == 3) {  // timing only: the accumulators reach the registers, no selection
            m1.x = fmaxf(m1.x, __uint_as_float(buf[c & 1][0] ^ buf[c & 1][31]));
            continue;
          }
          if (EPI == 0) {
#pragma unroll
            for (int i2 = 0; i2 < 16; ++i2) {
              uint32_t k0, k1;  // (bits & ~127) | column within the tile: one LOP3 each
              asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(k0) : "r"(buf[c & 1][2 * i2]), "r"(keymask), "r"((uint32_t)(c * 32 + 2 * i2)));
              asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(k1) : "r"(buf[c & 1][2 * i2 + 1]), "r"(keymask), "r"((uint32_t)(c * 32 + 2 * i2 + 1)));
              const float f0 = __uint_as_float(k0), f1 = __uint_as_float(k1);
              const float hi = fmaxf(f0, f1), lo = fminf(f0, f1);
              m2.x = fmaxf(fmaxf(m2.x, fminf(m1.x, hi)), lo);  // second largest of {m1, m2, f0, f1}
              m1.x = fmaxf(m1.x, hi);
            }
          } else {
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              // four consecutive accumulators v0..v3: chain A takes (v0, v2), chain B (v1, v3); (v0,v1) and (v2,v3) are
              // adjacent registers = packed operands
              uint32_t k[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(k[e]) : "r"(buf[c & 1][4 * i4 + e]), "r"(keymask), "r"((uint32_t)(c * 32 + 4 * i4 + e)));
              const float2 ka = make_float2(__uint_as_float(k[0]), __uint_as_float(k[1]));
              const float2 kb = make_float2(__uint_as_float(k[2]), __uint_as_float(k[3]));
              const float2 hi = make_float2(fmaxf(ka.x, kb.x), fmaxf(ka.y, kb.y));
              const float2 nhi = make_float2(-hi.x, -hi.y);
              const float2 lo = __fadd2_rn(__fadd2_rn(ka, kb), nhi);            // min(ka, kb) up to rounding
              const float2 m1n = make_float2(fmaxf(m1.x, hi.x), fmaxf(m1.y, hi.y));
              const float2 x = __fadd2_rn(__fadd2_rn(m1, hi), make_float2(-m1n.x, -m1n.y));  // min(m1, hi) up to rounding
              m2.x = fmaxf(fmaxf(m2.x, x.x), lo.x);
              m2.y = fmaxf(fmaxf(m2.y, x.y), lo.y);
              m1 = m1n;
            }
          }
        }
        if (m1.x != m1_in.x) btA = j;
        if (m1.y != m1_in.y) btB = j;
      }
      // merge the two chains and store the proposal of this row
      float b1, b2;
      int col;
      {
        const int cA = btA < 0 ? -1 : btA * P2_TILE_N + (int)(__float_as_uint(m1.x) & 0x7Fu);
        const int cB = btB < 0 ? -1 : btB * P2_TILE_N + (int)(__float_as_uint(m1.y) & 0x7Fu);
        if (cB >= 0 && (cA < 0 || m1.y > m1.x || (m1.y == m1.x && cB < cA))) {
          b1 = m1.y, col = cB, b2 = fmaxf(m1.x, fmaxf(m2.x, m2.y));
        } else {
          b1 = m1.x, col = cA, b2 = fmaxf(m1.y, fmaxf(m2.x, m2.y));
        }
      }
      const int row = (4 * g + 2 * sidx + (int)rank) * BLK + q * 32 + lane;
      Prop2 out;
      out.best = __uint_as_float(__float_as_uint(b1) & 0xFFFFFF80u);
      out.second = __uint_as_float(__float_as_uint(b2) & 0xFFFFFF80u);
      out.idx = col;
      prop[(size_t)p * K1p + row] = out;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's tensor / shared memory stay alive until the leader's MMAs have drained
  if (warp == W_ALLOC) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// k_tc_seq_fused: conversion AND proposal GEMM of a SEQUENCE in one kernel (K1p = K2p = 512).
//
// k_tc_convert is bound by HBM (it reads every descriptor as fp64), k_tc_gemm_pair by the tensor pipe and the
// epilogue's issue slots; run one after the other each leaves the other unit idle, and the fp16 image makes a round
// trip through HBM / L2 in between.  Here every CTA pair takes a CONTIGUOUS run of pairs u0 .. u1-1, i.e. frames
// u0 .. u1, and keeps three frames resident in shared memory as fp16 operand images: frame i is the B operand of pair
// i-1 and then -- the same bytes, untouched -- the A operand of pair i, while converter warps read frame i+1 from HBM
// (fp64 or fp32), convert it in registers and store it swizzled into the third buffer.  The fp16 image never exists in
// global memory; the descriptors are read from HBM once per sequence (plus one frame per CTA pair).
//
// One frame image per CTA (64 KB + 8 KB of K-extension rows): [sidx 2][K half 2][row 128][128 B], where the 128 rows of
// part sidx are rows 64 rank .. 64 rank + 63 of the frame's 128-row block 2 sidx followed by the same rows of block
// 2 sidx + 1.  As A operand part sidx is one M = 128 tile of this CTA; as B operand the 64 rows of block j are this
// CTA's half of the N = 128 tile j.  Arithmetic (fp16 rounding, MMA order, selection) is that of k_tc_convert +
// k_tc_gemm_pair: the proposals are identical.
//   barriers:  f_full[buf][j]  LEADER only: one arrival per converted quad of rows, 16 local + 16 remote (release.cluster):
//                              tile j of the frame in `buf` is in both CTAs' shared memory
//              f_empty[buf]    tcgen05.commit multicast: the MMAs of the pair that used `buf` as A have completed
//              t_full / t_empty as in k_tc_gemm_pair
// ---------------------------------------------------------------------------------------------
constexpr int FZ_NCW = 7;                                  // converter warps per CTA (16 warps: 4 per sub-partition)
constexpr int FZ_W_CONV = 8, FZ_W_ALLOC = FZ_W_CONV, FZ_W_MMA = 8 + FZ_NCW;  // 0-7: the two epilogue warpgroups
constexpr int FZ_THREADS = 32 * (9 + FZ_NCW);
constexpr int FZ_KP = 512;                                 // padded descriptor count this kernel is built for
constexpr int FZ_IMG_BYTES = 2 * BLK_BYTES;                // 64 KB per frame and CTA
constexpr int FZ_EXT_BYTES = 4 * P2_BEXT;                  // 8 KB
constexpr int FZ_BUF_BYTES = FZ_IMG_BYTES + FZ_EXT_BYTES;
constexpr int FZ_NBUF = 3;
constexpr int FZ_OFF_AX = FZ_NBUF * FZ_BUF_BYTES;
constexpr int FZ_OFF_BAR = FZ_OFF_AX + EXT_BYTES;
constexpr int FZ_SMEM_BYTES = FZ_OFF_BAR + 512;
constexpr int FZ_QROWS = 4;                                // rows a converter warp converts at a time ("quad")
constexpr int FZ_TQUADS = 64 / FZ_QROWS;                   // quads per tile and CTA
constexpr int FZ_FQUADS = 4 * FZ_TQUADS;                   // quads per frame and CTA
static_assert(FZ_SMEM_BYTES + 1024 <= 227 * 1024, "fused sequence matcher: shared memory");
static_assert(FZ_BUF_BYTES % 1024 == 0, "fused sequence matcher: layout");

struct FusedBarriers {
  uint64_t f_full[FZ_NBUF][4], f_empty[FZ_NBUF];
  uint64_t t_full[4], t_empty[4];
  uint32_t tmem_base;
};
static_assert(sizeof(FusedBarriers) <= 512, "FusedBarriers");

__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FZ_THREADS, 1)
k_tc_seq_fused(const T* __restrict__ L, int K, const int32_t* __restrict__ kc, int P, int ext_layout,
               float* __restrict__ nrm, FrameInfo* __restrict__ finfo, Prop2* __restrict__ prop) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t sAX = base + FZ_OFF_AX;
  FusedBarriers* bars = reinterpret_cast<FusedBarriers*>(smem + FZ_OFF_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int u0 = (int)((long long)P * cluster_id / nclusters), u1 = (int)((long long)P * (cluster_id + 1) / nclusters);
  const int npairs = u1 - u0;  // >= 1: the grid never has more CTA pairs than frame pairs

  if (warp == FZ_W_MMA && lane == 0) {
    for (int b = 0; b < FZ_NBUF; ++b) {
      for (int j = 0; j < 4; ++j) mbar_init(smem_u32(&bars->f_full[b][j]), 2 * FZ_TQUADS);
      mbar_init(smem_u32(&bars->f_empty[b]), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&bars->t_full[i]), 1);
      mbar_init(smem_u32(&bars->t_empty[i]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == FZ_W_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int r = threadIdx.x; r < BLK; r += FZ_THREADS) {  // constant A tile of the ninth K step
    const __half one = __float2half(1.0f), sc = __float2half(EXT_SCALE);
    uint4 w;
    w.x = (unsigned)__half_as_ushort(sc) | ((unsigned)__half_as_ushort(sc) << 16);
    w.y = (unsigned)__half_as_ushort(one) | ((unsigned)__half_as_ushort(one) << 16);
    w.z = w.w = 0u;
    *reinterpret_cast<uint4*>(smem + FZ_OFF_AX + ext_off(ext_layout, r, 0)) = w;
    *reinterpret_cast<uint4*>(smem + FZ_OFF_AX + ext_off(ext_layout, r, 1)) = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp >= FZ_W_CONV && warp < FZ_W_CONV + FZ_NCW) {
    // ===== converters: frame u0 + i -> buffer i % 3, this CTA's 64 rows of every 128-row block ========
    // Work item = a quad of 4 rows (1 KB each: lane l holds elements 4l .. 4l+3).  The loads of the next quad are in
    // flight while the current one is converted (two register sets), across tile and frame boundaries; only the
    // shared-memory writes wait for the buffer.  |x|^2 (fp32, see `convert`): the per-lane partial sums of the rows of
    // two quads are reduced together -- the xor-16, 8 and 4 steps of the butterfly as a reduce-scatter (a lane keeps the
    // rows its lane bits select), then xor 2, 1 -- and lane l ends with the norm of row (l >> 2) & 3 of quad l >> 4;
    // lanes 0, 4, .. 28 then write the K-extension rows and the norms of their rows in parallel.
    const int cw = warp - FZ_W_CONV;
    const int total = (npairs + 1) * FZ_FQUADS;  // quads of this CTA's stream; warp cw takes quads cw, cw + NCW, ...
    // this lane's byte offset inside a 128-byte image row group: K half, swizzled 16-byte chunk, 8-byte half of it
    // (the swizzle key tr & 7 = e & 3 | (quad & 1) << 2 is added per row below)
    const uint32_t lane_img = (uint32_t)((lane >> 4) * HALF_BYTES + ((lane & 1) << 3));
    const uint32_t lane_c = (uint32_t)((lane & 15) >> 1);
    // bounds of the frame a stream of quads is in: max |x|^2, "some |x|^2 does not fit the split".  (k_tc_convert also
    // tests every value against 60000; that test is implied by the one on |x|^2: a value above 14143 -- or a NaN / inf --
    // puts |x|^2 above NORM_MAX = 2e8, or makes it NaN.)
    struct Stats {
      float wmax;
      int wbad, fi;
    };
    auto flush_bounds = [&](Stats& st) {  // idempotent: a boundary frame is converted by two CTA pairs
      const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(st.wmax));  // non-negative floats order as uints
      const unsigned bd = __reduce_or_sync(0xffffffffu, (unsigned)st.wbad);
      if (lane == 0) {
        if (mx != 0u) atomicMax(&finfo[u0 + st.fi].max_bits, mx);
        if (bd) atomicOr(&finfo[u0 + st.fi].bad, 1);
      }
      st.wmax = 0.f;
      st.wbad = 0;
    };
    int cur_i = -1;  // newest frame whose buffer this warp has acquired
    // position of a quad in the stream, advanced by 2 FZ_NCW quads at a time without divisions
    struct Pos {
      int i, hq;  // frame (relative to u0), quad within the frame (tile j = hq >> 4, quad of the tile = hq & 15)
      int b;      // i % FZ_NBUF: the frame's buffer (kept incrementally: a division by 3 per use otherwise)
      __device__ __forceinline__ void advance(int by) {
        hq += by;
        if (hq >= FZ_FQUADS) {
          hq -= FZ_FQUADS;
          ++i;
          b = b == FZ_NBUF - 1 ? 0 : b + 1;
        }
      }
    };
    // valid rows of frame u0 + i (kc: per-frame descriptor counts), looked up once per frame
    int nf_i = -1, nf_n = K;
    auto frame_rows = [&](int i) -> int {
      if (kc && i != nf_i) {
        nf_i = i;
        nf_n = max(0, min(__ldg(kc + u0 + i), K));
      }
      return nf_n;
    };
    // (An L2 bulk prefetch of the quad a warp converts 2 .. 9 trips later was tried here and measured slower at every
    // distance, 0.602 -> 0.634 .. 0.685 ms per 4096 pairs, profiles/r02_fused_prefetch.log: the converters do not wait
    // for memory.)
    auto issue = [&](const Pos& q, double (&v)[FZ_QROWS][4]) {
      const int n = frame_rows(q.i);
      const int row0 = (q.hq >> 4) * BLK + 64 * (int)rank + (q.hq & 15) * FZ_QROWS;
      const T* src = L + ((size_t)(u0 + q.i) * K + row0) * ND + 4 * lane;
      if (row0 + FZ_QROWS <= n) {  // the usual quad: all four rows exist, no predicates, nothing to clear
#pragma unroll
        for (int e = 0; e < FZ_QROWS; ++e) load4(src + (size_t)e * ND, v[e]);
      } else {
#pragma unroll
        for (int e = 0; e < FZ_QROWS; ++e) {
          v[e][0] = v[e][1] = v[e][2] = v[e][3] = 0.0;
          if (row0 + e < n) load4(src + (size_t)e * ND, v[e]);
        }
      }
    };
    auto acquire = [&](const Pos& q) {  // first write of this warp into the buffer for frame q.i: the pair that used it as A must be done
      if (q.i > cur_i) {
        cur_i = q.i;
        mbar_wait(smem_u32(&bars->f_empty[q.b]), (uint32_t)(((q.i / FZ_NBUF) & 1) ^ 1));
        tc_fence_after();
      }
    };
    // part 1 of a quad: fp16 image rows into the buffer, this lane's partial |x|^2 of the four rows; frees the raw values.
    // |x|^2 is summed in FP32 from the float-rounded values (the ones the fp16 image is made of): FMA-pipe instructions
    // instead of 7 per row on the fp64 pipe, 32-bit shuffles, 4-cycle dependent adds.  ncu of the fp64 form
    // (profiles/r02_g_fused_source.md): the converter warps were never waiting for memory (long scoreboard 3 % of their
    // samples) -- they were busy, a third of the time throttled on the fp64 pipe, and the MMA issuer waited for them.
    // Error of the fp32 sum: the input rounding (2^-23 on x^2), one rounding per product and per level of the pairwise
    // tree (2 in the lane + 5 shuffle levels; all terms are non-negative, so every partial sum is below the total):
    // |t - |x|^2| <= 2^-20 |x|^2, carried by the rescore's bound (rescore_group / rescore_decide).
    auto convert = [&](const Pos& q, double (&v)[FZ_QROWS][4], float (&sq)[FZ_QROWS], Stats& st) {
      if (q.i != st.fi) {
        if (st.fi >= 0) flush_bounds(st);
        st.fi = q.i;
      }
      const int j = q.hq >> 4, qd = q.hq & 15;
      const int tr0 = (j & 1) * 64 + qd * FZ_QROWS;  // first row of the quad within the 128-row tile of part j >> 1
      unsigned char* dst0 = smem + q.b * FZ_BUF_BYTES + (j >> 1) * BLK_BYTES + tr0 * 128 + lane_img;
#pragma unroll
      for (int e = 0; e < FZ_QROWS; ++e) {
        const float f0 = (float)v[e][0], f1 = (float)v[e][1], f2 = (float)v[e][2], f3 = (float)v[e][3];
        __half2 h01 = __floats2half2_rn(f0, f1);
        __half2 h23 = __floats2half2_rn(f2, f3);
        uint2 packed;
        packed.x = *reinterpret_cast<unsigned*>(&h01);
        packed.y = *reinterpret_cast<unsigned*>(&h23);
        const uint32_t key = (uint32_t)(e | ((qd & 1) << 2));  // (tr0 + e) & 7
        *reinterpret_cast<uint2*>(dst0 + e * 128 + ((lane_c ^ key) << 4)) = packed;
        sq[e] = __fadd_rn(__fmaf_rn(f0, f0, __fmul_rn(f1, f1)), __fmaf_rn(f2, f2, __fmul_rn(f3, f3)));
      }
    };
    // part 2: |x|^2 of the eight rows of two quads (a: rows 0-3, b: rows 4-7) -- reduce-scatter over lane bits 4, 3, 2
    // (a lane keeps the rows its bits select), butterfly over bits 1, 0: lane l ends with the sum of row
    // (l >> 2) & 3 of quad l >> 4, a pairwise tree over the 32 lanes
    auto reduce8 = [&](const float (&sa4)[FZ_QROWS], const float (&sb4)[FZ_QROWS]) -> float {
      const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
      float w[4], u[2];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const float snd = b4 ? sa4[x] : sb4[x], kp = b4 ? sb4[x] : sa4[x];
        w[x] = __fadd_rn(kp, __shfl_xor_sync(0xffffffffu, snd, 16));
      }
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const float snd = b3 ? w[x] : w[x + 2], kp = b3 ? w[x + 2] : w[x];
        u[x] = __fadd_rn(kp, __shfl_xor_sync(0xffffffffu, snd, 8));
      }
      const float snd = b2 ? u[0] : u[1], kp = b2 ? u[1] : u[0];
      float t = __fadd_rn(kp, __shfl_xor_sync(0xffffffffu, snd, 4));
      t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 2));
      t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, 1));
      return t;
    };
    // part 3 (lanes 0, 4, .. 28: one row each): K-extension row = the exact four-slot split of -t/2 (the split of
    // k_tc_convert in fp32: t has 24 bits, every residual is exact), the norm t itself.  A frame with an out-of-range
    // VALUE is flagged as a whole (flush_bounds): its pairs never use the proposal
    auto finish = [&](const Pos& q, float t, float& fn_out, int& bad_out) {
      const int j = q.hq >> 4;
      const int rl = (q.hq & 15) * FZ_QROWS + ((lane >> 2) & 3);
      const int row = j * BLK + 64 * (int)rank + rl;
      const int f = u0 + q.i;
      const int n = frame_rows(q.i);
      const bool bad = !(t <= (float)NORM_MAX);
      __half h[4];
      if (row < n && !bad) {
        float gq = -0.5f * t;
        h[0] = __float2half_rn(gq * (1.0f / 4096.0f));
        gq = __fmaf_rn(-4096.0f, __half2float(h[0]), gq);
        h[1] = __float2half_rn(gq * (1.0f / 4096.0f));
        gq = __fmaf_rn(-4096.0f, __half2float(h[1]), gq);
        h[2] = __float2half_rn(gq);
        gq = __fsub_rn(gq, __half2float(h[2]));
        h[3] = __float2half_rn(gq);
      } else {
        h[0] = __float2half(-60000.0f);
        h[1] = h[2] = h[3] = __float2half(0.0f);
      }
      uint4 w;
      w.x = (unsigned)__half_as_ushort(h[0]) | ((unsigned)__half_as_ushort(h[1]) << 16);
      w.y = (unsigned)__half_as_ushort(h[2]) | ((unsigned)__half_as_ushort(h[3]) << 16);
      w.z = w.w = 0u;
      unsigned char* e0 = smem + q.b * FZ_BUF_BYTES + FZ_IMG_BYTES + j * P2_BEXT;  // SWIZZLE_32B rows (ext_off)
      *reinterpret_cast<uint4*>(e0 + ext_off(0, rl, 0)) = w;
      *reinterpret_cast<uint4*>(e0 + ext_off(0, rl, 1)) = make_uint4(0u, 0u, 0u, 0u);
      nrm[(size_t)f * FZ_KP + row] = (row < n) ? t : INFINITY;
      fn_out = row < n ? t : 0.f;
      bad_out = (row < n && bad) ? 1 : 0;
    };
    auto arrive = [&](const Pos& q) {
      const uint32_t bar = smem_u32(&bars->f_full[q.b][q.hq >> 4]);
      if (leader) mbar_arrive(bar);
      else mbar_arrive_cluster(mapa_u32(bar, 0));
    };
    // Two quads per trip (streams a and b, FZ_NCW quads apart).  The raw values of a quad are dead once `convert` has
    // run, so the loads of the quad two trips ahead are issued right behind it and stay in flight during the serial
    // rest of the trip (the shuffle chains and the dependent conversions of the split), which the two quads run
    // interleaved.
    double va[FZ_QROWS][4], vb[FZ_QROWS][4];
    Stats sa{0.f, 0, -1}, sb{0.f, 0, -1};
    Pos qa{0, cw, 0}, qb{0, cw, 0};
    qb.advance(FZ_NCW);
    int Q = cw;  // stream index of qa
    if (Q < total) issue(qa, va);
    if (Q + FZ_NCW < total) issue(qb, vb);
#pragma unroll 1
    for (; Q < total; Q += 2 * FZ_NCW) {
      const bool has_b = Q + FZ_NCW < total;
      const Pos ca = qa, cb = qb;
      float sqa[FZ_QROWS], sqb[FZ_QROWS] = {0.f, 0.f, 0.f, 0.f};
      acquire(ca);
      convert(ca, va, sqa, sa);
      qa.advance(2 * FZ_NCW);
      if (Q + 2 * FZ_NCW < total) issue(qa, va);
      if (has_b) {
        acquire(cb);
        convert(cb, vb, sqb, sb);
      }
      qb.advance(2 * FZ_NCW);
      if (Q + 3 * FZ_NCW < total) issue(qb, vb);
      const float t = reduce8(sqa, sqb);
      if ((lane & 3) == 0 && (lane < 16 || has_b)) {  // lanes 0-15 hold the rows of quad a, 16-31 those of quad b
        Pos cq;
        cq.i = lane < 16 ? ca.i : cb.i;
        cq.hq = lane < 16 ? ca.hq : cb.hq;
        cq.b = lane < 16 ? ca.b : cb.b;
        float fn;
        int bd;
        finish(cq, t, fn, bd);
        if (lane < 16) sa.wmax = fmaxf(sa.wmax, fn), sa.wbad |= bd;
        else sb.wmax = fmaxf(sb.wmax, fn), sb.wbad |= bd;
      }
      // the quads are written: generic-proxy writes -> async proxy, then tell the issuer (CTA-scope release, as every
      // other remote arrive of these kernels: what the arrive publishes are this SM's shared-memory writes, already
      // pushed to the async proxy by the fence; a cluster-scope release would also wait for the global norm stores:
      // measured 1.00 -> 0.78 ms per 4096 pairs)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        arrive(ca);
        if (has_b) arrive(cb);
      }
    }
    if (sa.fi >= 0) flush_bounds(sa);
    if (sb.fi >= 0) flush_bounds(sb);
    // the commit that releases the last pair's A buffer targets THIS CTA's barrier too: stay until it has arrived
    if (cw == 0) mbar_wait(smem_u32(&bars->f_empty[(npairs - 1) % FZ_NBUF]), (uint32_t)(((npairs - 1) / FZ_NBUF) & 1));
  } else if (warp == FZ_W_MMA) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (one thread of the leader CTA, for both SMs) =================================
      uint32_t use[4] = {0, 0, 0, 0};
      for (int i = 0; i < npairs; ++i) {
        const int bA = i % FZ_NBUF, bB = (i + 1) % FZ_NBUF;
        const uint32_t phA = (uint32_t)((i / FZ_NBUF) & 1), phB = (uint32_t)(((i + 1) / FZ_NBUF) & 1);
        const uint32_t sAf = base + bA * FZ_BUF_BYTES, sBf = base + bB * FZ_BUF_BYTES;
        for (int j = 0; j < 4; ++j) mbar_wait_cl(smem_u32(&bars->f_full[bA][j]), phA);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mbar_wait_cl(smem_u32(&bars->f_full[bB][j]), phB);
          tc_fence_after();
          const uint32_t sBt = sBf + (j >> 1) * BLK_BYTES + (j & 1) * P2_BHALF;
#pragma unroll
          for (int sidx = 0; sidx < 2; ++sidx) {
            const int slot = (j & 1) * 2 + sidx;
            mbar_wait_cl(smem_u32(&bars->t_empty[slot]), (use[slot] & 1) ^ 1);
            ++use[slot];
            tc_fence_after();
            const uint32_t d = tmem + (uint32_t)(slot * P2_TILE_N);
#pragma unroll
            for (int k = 0; k < ND / 16; ++k) {
              const uint32_t ko = (uint32_t)((k >> 2) * HALF_BYTES + (k & 3) * 32);
              tc_mma_f16_2cta(d, umma_desc(sAf + sidx * BLK_BYTES + ko), umma_desc(sBt + ko), IDESC_PAIR, k > 0 ? 1u : 0u);
            }
            tc_mma_f16_2cta(d, umma_desc_ext(sAX, ext_layout),
                            umma_desc_ext(sBf + FZ_IMG_BYTES + j * P2_BEXT, ext_layout), IDESC_PAIR, 1u);
            tc_commit_mc2(smem_u32(&bars->t_full[slot]));
          }
        }
        tc_commit_mc2(smem_u32(&bars->f_empty[bA]));
      }
    }
  } else if (warp < FZ_W_CONV) {
    // ===== epilogue: the selection of k_tc_gemm_pair (EPI = 1); warpgroup s owns part s of this CTA ==
    const int sidx = warp >> 2;
    const int q = warp & 3;
    const uint32_t tel0 = mapa_u32(smem_u32(&bars->t_empty[sidx]), 0);
    const uint32_t tel1 = mapa_u32(smem_u32(&bars->t_empty[2 + sidx]), 0);
    uint32_t use0 = 0, use1 = 0;
    uint32_t keymask;
    asm volatile("mov.u32 %0, 0xFFFFFF80;" : "=r"(keymask));
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sidx * P2_TILE_N);
    uint32_t buf[2][32];
    auto wait_full = [&](int stg) {
      uint32_t& use = stg == 0 ? use0 : use1;
      mbar_wait(smem_u32(&bars->t_full[stg * 2 + sidx]), use & 1);
      ++use;
      tc_fence_after();
    };
    auto release = [&](int stg) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(smem_u32(&bars->t_empty[stg * 2 + sidx]));
        else mbar_arrive_cluster(stg == 0 ? tel0 : tel1);
      }
    };
    wait_full(0);
    tc_ld_32x32(tbase, buf[0]);
    for (int i = 0; i < npairs; ++i) {
      float2 m1 = make_float2(-INFINITY, -INFINITY), m2 = make_float2(-INFINITY, -INFINITY);
      int btA = -1, btB = -1;
      for (int j = 0; j < 4; ++j) {
        const int stg = j & 1;
        const float2 m1_in = m1;
        const bool has_next = (j + 1 < 4) || (i + 1 < npairs);
#pragma unroll
        for (int c = 0; c < P2_TILE_N / 32; ++c) {
          tc_ld_wait(buf[c & 1]);
          if (c + 1 < P2_TILE_N / 32) {
            tc_ld_32x32(tbase + (uint32_t)(stg * 2 * P2_TILE_N + (c + 1) * 32), buf[(c + 1) & 1]);
          } else {
            release(stg);
            if (has_next) {
              wait_full(stg ^ 1);
              tc_ld_32x32(tbase + (uint32_t)((stg ^ 1) * 2 * P2_TILE_N), buf[0]);
            }
          }
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            uint32_t k[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(k[e]) : "r"(buf[c & 1][4 * i4 + e]), "r"(keymask), "r"((uint32_t)(c * 32 + 4 * i4 + e)));
            const float2 ka = make_float2(__uint_as_float(k[0]), __uint_as_float(k[1]));
            const float2 kb = make_float2(__uint_as_float(k[2]), __uint_as_float(k[3]));
            const float2 hi = make_float2(fmaxf(ka.x, kb.x), fmaxf(ka.y, kb.y));
            const float2 nhi = make_float2(-hi.x, -hi.y);
            const float2 lo = __fadd2_rn(__fadd2_rn(ka, kb), nhi);
            const float2 m1n = make_float2(fmaxf(m1.x, hi.x), fmaxf(m1.y, hi.y));
            const float2 x = __fadd2_rn(__fadd2_rn(m1, hi), make_float2(-m1n.x, -m1n.y));
            m2.x = fmaxf(fmaxf(m2.x, x.x), lo.x);
            m2.y = fmaxf(fmaxf(m2.y, x.y), lo.y);
            m1 = m1n;
          }
        }
        if (m1.x != m1_in.x) btA = j;
        if (m1.y != m1_in.y) btB = j;
      }
      float b1, b2;
      int col;
      {
        const int cA = btA < 0 ? -1 : btA * P2_TILE_N + (int)(__float_as_uint(m1.x) & 0x7Fu);
        const int cB = btB < 0 ? -1 : btB * P2_TILE_N + (int)(__float_as_uint(m1.y) & 0x7Fu);
        if (cB >= 0 && (cA < 0 || m1.y > m1.x || (m1.y == m1.x && cB < cA))) {
          b1 = m1.y, col = cB, b2 = fmaxf(m1.x, fmaxf(m2.x, m2.y));
        } else {
          b1 = m1.x, col = cA, b2 = fmaxf(m1.y, fmaxf(m2.x, m2.y));
        }
      }
      const int r = q * 32 + lane;  // row within part sidx: rows 64 rank .. of block 2 sidx, then of block 2 sidx + 1
      const int row = (2 * sidx + (r >> 6)) * BLK + 64 * (int)rank + (r & 63);
      Prop2 out;
      out.best = __uint_as_float(__float_as_uint(b1) & 0xFFFFFF80u);
      out.second = __uint_as_float(__float_as_uint(b2) & 0xFFFFFF80u);
      out.idx = col;
      prop[(size_t)(u0 + i) * FZ_KP + row] = out;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == FZ_W_ALLOC) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// k_tc_rescore: exact candidate distance + certified ratio-test decision.  One warp handles 32
// rows at a time: coalesced loads of the row and of its candidate column, delta*delta per
// element into shared memory, then lane r sums row r strictly in bin order.
// ---------------------------------------------------------------------------------------------
constexpr int RS_ROWS = 16;  // rows per warp pass: 16 x 129 products of the accumulation type = 16.5 KB (double)

template <typename T, typename ACC>
__device__ __forceinline__ void load_row4(const T* src, ACC* v);
template <>
__device__ __forceinline__ void load_row4<double, double>(const double* src, double* v) {
  const double2 x = __ldg(reinterpret_cast<const double2*>(src));
  const double2 y = __ldg(reinterpret_cast<const double2*>(src) + 1);
  v[0] = x.x, v[1] = x.y, v[2] = y.x, v[3] = y.y;
}
template <>
__device__ __forceinline__ void load_row4<float, float>(const float* src, float* v) {
  const float4 x = __ldg(reinterpret_cast<const float4*>(src));
  v[0] = x.x, v[1] = x.y, v[2] = x.z, v[3] = x.w;
}
template <>
__device__ __forceinline__ void load_row4<float, double>(const float* src, double* v) {
  const float4 x = __ldg(reinterpret_cast<const float4*>(src));
  v[0] = x.x, v[1] = x.y, v[2] = x.z, v[3] = x.w;
}

constexpr int RS_GROUPS = 4;  // 16-row groups per (one-warp) block: 4x fewer blocks to schedule

template <typename T, typename ACC>
__device__ __forceinline__ void rescore_group(ACC (*sprod)[ND + 1], int p, int row0, int n1, int n2,
                                              const T* __restrict__ L1, const T* __restrict__ L2, int K1, int K2,
                                              int K1p, float thresh, const Prop* __restrict__ prop,
                                              const float* __restrict__ nrmA, const PairInfo& pi, int need_score,
                                              int v2, int a_shared, MatchRow* __restrict__ rows,
                                              int32_t* __restrict__ row_list, int32_t* __restrict__ row_list_n) {
  const int lane = threadIdx.x & 31;
  const int pa = a_shared ? 0 : p;  // one L1 set for every problem
  const int nr = min(RS_ROWS, n1 - row0);
  const int k1 = row0 + lane;
  const bool mine = lane < nr;

  // ---- what the approximate keys already decide ------------------------------------------------
  // v'(j) = d2~(j) + C - |a|^2; |d2~(j) - d2(j)| <= m for every column j of the pair:
  // fp16 rounding of both operands (2^-11 relative each, 2^-25 absolute below the normal range), fp32
  // accumulation of 128 products, fp32 norms / bias / epilogue, the 7 truncated key bits, the
  // reference's own rounding when it accumulates in float; 25 % slack on top.
  Prop my;
  my.best = my.second = INFINITY;
  my.idx = -1;
  double na = 0.0, m = 0.0, C = 0.0;
  bool ambiguous = false, need_exact = false;
  MatchRow out;
  out.best = INFINITY;
  out.bestk = -1;
  out.accept = 0;
  if (mine && n2 > 0) {
    my = prop[(size_t)p * K1p + k1];
    na = (double)nrmA[(size_t)pa * K1p + k1];
    const double nbm = (double)__uint_as_float(pi.bmax_bits);
    C = (double)(1.0625f * __uint_as_float(pi.amax_bits));
    const double ra = sqrt(na), rbm = sqrt(nbm);
    m = 1.25 * ((1.0 / 512 + 1.0 / 16384) * ra * rbm + (na + nbm + C) * (1.0 / 16384) + (ra + rbm) * (1.0 / 1048576));
    // CTA-pair proposal: d2~ = |a|^2 - 2 s with s = a~.b~ - |b|^2/2 accumulated in the contraction (the fp16 split of
    // the norm is exact down to 2^-25 absolute, the 7 replaced key bits of s cost |s| 2^-16 <= (na + nbm) 2^-16).
    // k_tc_seq_fused sums |x|^2 in fp32 from the float-rounded values: both norms are within 2^-20 relative of the
    // exact ones (derivation at its `convert`), the (na + nbm) 2^-19 term
    if (v2) m = 1.25 * ((1.0 / 512 + 1.0 / 16384) * ra * rbm + (na + nbm) * (1.0 / 16384 + 1.0 / 524288) + (ra + rbm) * (1.0 / 1048576) + 1.0 / 4194304);
    const bool keys_ok = v2 ? (my.best > -INFINITY) : ((C - na > m) && (my.best < INFINITY));
    if (pi.bad || my.idx < 0 || my.idx >= n2 || !(thresh > 0.f) || !keys_ok) {
      ambiguous = true;  // keys may be meaningless (range, sign) -> exact kernel
    } else {
      const double b1 = v2 ? na - 2.0 * (double)my.best : (double)my.best - C + na;
      const double s2 = v2 ? na - 2.0 * (double)my.second : (double)my.second - C + na;
      const double L1b = b1 - m;  // exact best  >= L1b
      const double U2b = s2 + m;  // exact second_best <= U2b
      if (L1b > 0.0 && __fmul_rn(thresh, (float)L1b) > (float)U2b) {
        out.accept = 0;  // ratio test fails for sure (siftmatch.c:122): row never output
      } else if (!need_score && b1 + m < s2 - m && __fmul_rn(thresh, (float)(b1 + m)) <= (float)(s2 - m)) {
        // the candidate is the unique exact minimum (its distance <= b1 + m < s2 - m <= every other
        // column's) and the test passes for every (best, second_best) in the brackets: accepted
        // without recomputing the distance (the caller does not ask for the scores)
        out.accept = 1;
        out.bestk = my.idx;
        out.best = b1;
      } else {
        need_exact = true;
      }
    }
  }

  // ---- exact distance to the candidate for the rows that may be accepted -------------------------
  unsigned todo = __ballot_sync(0xffffffffu, need_exact);
  while (todo) {
    int rr[4], ii[4];
    ACC va[4][4], vb[4][4];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      rr[u] = -1;
      ii[u] = 0;
      if (todo) {
        rr[u] = __ffs(todo) - 1;
        todo &= todo - 1;
        ++cnt;
      }
      if (rr[u] >= 0) ii[u] = __shfl_sync(0xffffffffu, my.idx, rr[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)  // all loads of up to four rows in flight together
      if (rr[u] >= 0) {
        load_row4<T, ACC>(L1 + ((size_t)pa * K1 + row0 + rr[u]) * ND + 4 * lane, va[u]);
        load_row4<T, ACC>(L2 + ((size_t)p * K2 + ii[u]) * ND + 4 * lane, vb[u]);
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (rr[u] >= 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const ACC d = va[u][e] - vb[u][e];
          sprod[rr[u]][4 * lane + e] = d * d;
        }
      }
  }
  __syncwarp();
  if (need_exact) {
    ACC acc = 0;
    for (int bin = 0; bin < ND; ++bin) acc += sprod[lane][bin];  // strictly in order (siftmatch.c:101-107)
    const double d1 = (double)acc;
    const double s2 = v2 ? na - 2.0 * (double)my.second : (double)my.second - C + na;
    const double L2b = s2 - m, U2b = s2 + m;  // every other column's exact distance is >= L2b; one is <= U2b
    if (d1 < L2b) {
      // unique exact minimum: best = d1, bestk = idx; second_best in [L2b, U2b]
      const float lhs = __fmul_rn(thresh, (float)acc);
      out.best = d1;
      out.bestk = my.idx;
      if (lhs <= (float)L2b)
        out.accept = 1;
      else if (lhs > (float)U2b)
        out.accept = 0;
      else
        ambiguous = true;
    } else {
      // best >= L2b, second_best <= max(d1, U2b): rejected for sure when even that fails the test
      const double hi = d1 > U2b ? d1 : U2b;
      if (L2b > 0.0 && __fmul_rn(thresh, (float)L2b) > (float)hi)
        out.accept = 0;
      else
        ambiguous = true;
    }
  }
  if (!mine) return;
  if (ambiguous) {
    const int slot = atomicAdd(row_list_n, 1);
    row_list[slot] = p * K1 + k1;
  }
  rows[(size_t)p * K1 + k1] = out;
}

template <typename T, typename ACC>
__global__ void __launch_bounds__(32)
k_tc_rescore(const T* __restrict__ L1, const T* __restrict__ L2, int K1, int K2, int K1p, int K2p,
             const int32_t* __restrict__ k1c, const int32_t* __restrict__ k2c, float thresh,
             const Prop* __restrict__ prop, const float* __restrict__ nrmA, const PairInfo* __restrict__ info,
             int need_score, int v2, int a_shared, MatchRow* __restrict__ rows, int32_t* __restrict__ row_list,
             int32_t* __restrict__ row_list_n) {
  __shared__ ACC sprod[RS_ROWS][ND + 1];
  const int p = blockIdx.y;
  const int n1 = k1c ? max(0, min(k1c[a_shared ? 0 : p], K1)) : K1;
  const int n2 = k2c ? max(0, min(k2c[p], K2)) : K2;
  const PairInfo pi = info[p];
  for (int grp = 0; grp < RS_GROUPS; ++grp) {
    const int row0 = (blockIdx.x * RS_GROUPS + grp) * RS_ROWS;
    if (row0 >= n1) return;
    rescore_group<T, ACC>(sprod, p, row0, n1, n2, L1, L2, K1, K2, K1p, thresh, prop, nrmA, pi, need_score, v2, a_shared,
                          rows, row_list, row_list_n);
    __syncwarp();  // sprod is reused by the next group
  }
}

// ---------------------------------------------------------------------------------------------
// k_tc_rescore2: the same decisions as k_tc_rescore in two phases per 256-row block.  Phase A: one THREAD per row
// decides what the proposal's brackets decide (rejected for sure / accepted without recomputing) and lists the rows
// that need the exact candidate distance; phase B: two warps take the listed rows 32 at a time (coalesced loads of the
// row and of its candidate column, products in shared memory, lane r sums row r strictly in bin order).  k_tc_rescore
// ran one warp per 16 rows with 16.5 KB of shared memory per warp: 13 warps per SM, half of their lanes idle, for a
// phase-B population of a few per cent of the rows (ncu r02_p: 85 us per 4096 pairs, warps active 18 %).
// ---------------------------------------------------------------------------------------------
constexpr int RS2_THREADS = 256;
constexpr int RS2_BW = 2;      // warps of the exact pass
constexpr int RS2_ROWS = 32;   // rows per warp pass

struct RsDecision {
  Prop my;
  double na, m;
  MatchRow out;
  bool need_exact, ambiguous;
};

__device__ __forceinline__ RsDecision rescore_decide(int p, int pa, int k1, int n2, int K1p, float thresh,
                                                     const Prop* __restrict__ prop, const float* __restrict__ nrmA,
                                                     const PairInfo& pi, int need_score) {
  RsDecision d;
  d.out.best = INFINITY;
  d.out.bestk = -1;
  d.out.accept = 0;
  d.need_exact = d.ambiguous = false;
  d.my.best = d.my.second = INFINITY;
  d.my.idx = -1;
  d.na = d.m = 0.0;
  if (n2 <= 0) return d;
  d.my = prop[(size_t)p * K1p + k1];
  const double na = (double)nrmA[(size_t)pa * K1p + k1];
  const double nbm = (double)__uint_as_float(pi.bmax_bits);
  const double ra = sqrt(na), rbm = sqrt(nbm);
  // error bound of the CTA-pair proposal (see rescore_group)
  const double m = 1.25 * ((1.0 / 512 + 1.0 / 16384) * ra * rbm + (na + nbm) * (1.0 / 16384 + 1.0 / 524288) + (ra + rbm) * (1.0 / 1048576) + 1.0 / 4194304);
  d.na = na;
  d.m = m;
  if (pi.bad || d.my.idx < 0 || d.my.idx >= n2 || !(thresh > 0.f) || !(d.my.best > -INFINITY)) {
    d.ambiguous = true;
    return d;
  }
  const double b1 = na - 2.0 * (double)d.my.best, s2 = na - 2.0 * (double)d.my.second;
  const double L1b = b1 - m, U2b = s2 + m;
  if (L1b > 0.0 && __fmul_rn(thresh, (float)L1b) > (float)U2b) {
    d.out.accept = 0;
  } else if (!need_score && b1 + m < s2 - m && __fmul_rn(thresh, (float)(b1 + m)) <= (float)(s2 - m)) {
    d.out.accept = 1;
    d.out.bestk = d.my.idx;
    d.out.best = b1;
  } else {
    d.need_exact = true;
  }
  return d;
}

template <typename T, typename ACC>
__global__ void __launch_bounds__(RS2_THREADS)
k_tc_rescore2(const T* __restrict__ L1, const T* __restrict__ L2, int K1, int K2, int K1p, int K2p,
              const int32_t* __restrict__ k1c, const int32_t* __restrict__ k2c, float thresh,
              const Prop* __restrict__ prop, const float* __restrict__ nrmA, const PairInfo* __restrict__ info,
              int need_score, int a_shared, MatchRow* __restrict__ rows, int32_t* __restrict__ row_list,
              int32_t* __restrict__ row_list_n) {
  extern __shared__ __align__(16) unsigned char rs2_smem[];
  ACC(*sprod)[ND + 1] = reinterpret_cast<ACC(*)[ND + 1]>(rs2_smem);  // [warp * 32 + row][bin]
  __shared__ int s_list[RS2_THREADS];
  __shared__ int s_n;
  const int p = blockIdx.y, pa = a_shared ? 0 : p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n1 = k1c ? max(0, min(k1c[pa], K1)) : K1;
  const int n2 = k2c ? max(0, min(k2c[p], K2)) : K2;
  const int row0 = blockIdx.x * RS2_THREADS;
  if (row0 >= n1) return;
  const PairInfo pi = info[p];
  if (tid == 0) s_n = 0;
  __syncthreads();
  // ---- phase A ------------------------------------------------------------------------------------
  const int k1 = row0 + tid;
  if (k1 < n1) {
    const RsDecision d = rescore_decide(p, pa, k1, n2, K1p, thresh, prop, nrmA, pi, need_score);
    if (d.need_exact) {
      s_list[atomicAdd(&s_n, 1)] = k1;
    } else {
      if (d.ambiguous) row_list[atomicAdd(row_list_n, 1)] = p * K1 + k1;
      rows[(size_t)p * K1 + k1] = d.out;
    }
  }
  __syncthreads();
  // ---- phase B ------------------------------------------------------------------------------------
  const int n = s_n;
  if (warp >= RS2_BW) return;
  ACC(*mine)[ND + 1] = sprod + warp * RS2_ROWS;
  for (int base = warp * RS2_ROWS; base < n; base += RS2_BW * RS2_ROWS) {
    const int nr = min(RS2_ROWS, n - base);
    const int r1 = lane < nr ? s_list[base + lane] : -1;
    RsDecision d;
    d.my.idx = 0;
    if (r1 >= 0) d = rescore_decide(p, pa, r1, n2, K1p, thresh, prop, nrmA, pi, need_score);
    for (int r = 0; r < nr; r += 4) {  // up to four rows' loads in flight together
      ACC va[4][4], vb[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r + u < nr) {
          const int rr = __shfl_sync(0xffffffffu, r1, r + u), ii = __shfl_sync(0xffffffffu, d.my.idx, r + u);
          load_row4<T, ACC>(L1 + ((size_t)pa * K1 + rr) * ND + 4 * lane, va[u]);
          load_row4<T, ACC>(L2 + ((size_t)p * K2 + ii) * ND + 4 * lane, vb[u]);
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r + u < nr) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const ACC dd = va[u][e] - vb[u][e];
            mine[r + u][4 * lane + e] = dd * dd;
          }
        }
    }
    __syncwarp();
    if (r1 >= 0) {
      ACC acc = 0;
      for (int bin = 0; bin < ND; ++bin) acc += mine[lane][bin];  // strictly in order (siftmatch.c:101-107)
      const double d1 = (double)acc;
      const double s2 = d.na - 2.0 * (double)d.my.second;
      const double L2b = s2 - d.m, U2b = s2 + d.m;
      MatchRow out = d.out;
      bool ambiguous = false;
      if (d1 < L2b) {
        const float lhs = __fmul_rn(thresh, (float)acc);
        out.best = d1;
        out.bestk = d.my.idx;
        if (lhs <= (float)L2b)
          out.accept = 1;
        else if (lhs > (float)U2b)
          out.accept = 0;
        else
          ambiguous = true;
      } else {
        const double hi = d1 > U2b ? d1 : U2b;
        if (L2b > 0.0 && __fmul_rn(thresh, (float)L2b) > (float)hi)
          out.accept = 0;
        else
          ambiguous = true;
      }
      if (ambiguous) row_list[atomicAdd(row_list_n, 1)] = p * K1 + r1;
      rows[(size_t)p * K1 + r1] = out;
    }
    __syncwarp();
  }
}

template <typename ACC>
constexpr size_t rs2_smem_bytes() {
  return sizeof(ACC) * (size_t)RS2_BW * RS2_ROWS * (ND + 1);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// descriptor counts are padded to whole CTA-pair tiles (256 rows / columns)
static inline int pad128(int k) { return (k + 255) / 256 * 256; }

bool match_tc_supported(int cls, int K1, int K2, int ND) {
  return (cls == PRE3_CLASS_DOUBLE || cls == PRE3_CLASS_SINGLE || cls == PRE3_CLASS_DOUBLE_F32) && ND == tc::ND &&
         K1 >= 1 && K2 >= 1;
}

size_t match_tc_workspace_bytes(int P, int K1, int K2) {
  const size_t K1p = pad128(K1), K2p = pad128(K2);
  size_t b = 0;
  b += align_up((size_t)(P + 1) * K1p * tc::ND * 2, 1024) + 1024;
  b += align_up((size_t)P * K2p * tc::ND * 2, 1024) + 1024;
  b += align_up((size_t)(P + 1) * K1p * 4) + align_up((size_t)P * K2p * 4);
  b += align_up((size_t)(P + 1) * K1p * 32, 1024) + align_up((size_t)P * K2p * 32, 1024) + 2048;  // K-extension tiles
  b += align_up((size_t)P * K1p * sizeof(tc::Prop));
  b += align_up((size_t)P * sizeof(tc::PairInfo)) + 2 * align_up((size_t)(P + 1) * sizeof(tc::FrameInfo));
  b += align_up((size_t)P * K1 * 4) + 256;
  return b + 4096;
}

int launch_match_tc(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int P, int K1, int K2, int ND,
                    const int32_t* dk1, const int32_t* dk2, float thresh, int need_score, MatchRow* drows) {
  using namespace tc;
  if (!match_tc_supported(cls, K1, K2, ND)) return fail(ctx, PRE3_ERR_ARG, "tensor-core matcher: unsupported shape");
  if (P <= 0) return PRE3_OK;
  const int K1p = pad128(K1), K2p = pad128(K2);
  // dL2 == nullptr: SEQUENCE mode.  dL1 holds P + 1 descriptor sets (consecutive frames); pair p matches set p
  // against set p + 1.  Every set is converted once and serves as the A operand of pair p and the B operand of
  // pair p - 1: half the conversion traffic of 2 P independent sets.
  const bool seq = dL2 == nullptr;
  if (seq && K1 != K2) return fail(ctx, PRE3_ERR_ARG, "sequence mode needs the same descriptor count per frame");
  const int shared = (!seq && ctx->l1_shared) ? 1 : 0;  // ONE L1 set for all P problems (the sweep entry points)
  const int FA = seq ? P + 1 : (shared ? 1 : P);  // descriptor sets behind dL1
  static const int ext_layout = getenv("PRE3_TC_EXTLAYOUT") ? atoi(getenv("PRE3_TC_EXTLAYOUT")) : 0;
  // carve (1024-byte aligned operand images: TMA bulk copies need 16, the smem tiles 1024)
  auto take1k = [&](size_t bytes) {
    ctx->ws_off = align_up(ctx->ws_off, 1024);
    return ws_take<unsigned char>(ctx, bytes);
  };
  const size_t set_bytes = (size_t)K1p * ND * 2;
  unsigned char* imgA = take1k((size_t)FA * set_bytes);
  unsigned char* imgB = seq ? imgA + set_bytes : take1k((size_t)P * K2p * ND * 2);
  unsigned char* extA = take1k((size_t)FA * K1p * 32);
  unsigned char* extB = seq ? extA + (size_t)K1p * 32 : take1k((size_t)P * K2p * 32);
  float* nrmA = ws_take<float>(ctx, (size_t)FA * K1p);
  float* nrmB = seq ? nrmA + K1p : ws_take<float>(ctx, (size_t)P * K2p);
  Prop* prop = ws_take<Prop>(ctx, (size_t)P * K1p);
  PairInfo* info = ws_take<PairInfo>(ctx, P);
  FrameInfo* fa = ws_take<FrameInfo>(ctx, (size_t)FA);
  FrameInfo* fb = seq ? fa + 1 : ws_take<FrameInfo>(ctx, (size_t)P);
  int32_t* list = ws_take<int32_t>(ctx, (size_t)P * K1);
  int32_t* list_n = ws_take<int32_t>(ctx, 64);
  PRE3_TRY(pipe_enter(ctx, PS_CONVERT));
  PRE3_CUDA(cudaMemsetAsync(fa, 0, sizeof(FrameInfo) * (size_t)FA, ctx->stream));
  if (!seq) PRE3_CUDA(cudaMemsetAsync(fb, 0, sizeof(FrameInfo) * (size_t)P, ctx->stream));
  PRE3_CUDA(cudaMemsetAsync(list_n, 0, sizeof(int32_t), ctx->stream));
  if (seq) {
    dL2 = (const char*)dL1 + (size_t)K1 * ND * (cls == PRE3_CLASS_DOUBLE ? 8 : 4);
    if (dk1) dk2 = dk1 + 1;
  }
  static const int use_v1 = getenv("PRE3_TC_V1") ? atoi(getenv("PRE3_TC_V1")) : 0;
  static const int exp_mode = getenv("PRE3_TC_EXP") ? atoi(getenv("PRE3_TC_EXP")) : 0;
  // A sequence of 512-row frames: conversion and GEMM in ONE kernel (k_tc_seq_fused; PRE3_TC_FUSED=0 keeps them apart)
  const int fused_env = getenv("PRE3_TC_FUSED") ? atoi(getenv("PRE3_TC_FUSED")) : 1;
  static const int epi_env = getenv("PRE3_TC_EPI") ? atoi(getenv("PRE3_TC_EPI")) : 1;
  const bool fused = seq && fused_env && !use_v1 && exp_mode == 0 && epi_env && ext_layout == 0 && K1p == FZ_KP && K2p == FZ_KP;
  if (fused) {
    {
      Span span__(ctx, T_MATCH_FUSED);
      const int grid = 2 * std::min(P, ctx->sm_count / 2);
#define PRE3_FUSED(T)                                                                                                \
  do {                                                                                                               \
    static bool attr_done = false;                                                                                   \
    if (!attr_done) {                                                                                                \
      PRE3_CUDA(cudaFuncSetAttribute(k_tc_seq_fused<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM_BYTES)); \
      attr_done = true;                                                                                              \
    }                                                                                                                \
    k_tc_seq_fused<T><<<grid, FZ_THREADS, FZ_SMEM_BYTES, ctx->stream>>>((const T*)dL1, K1, dk1, P, ext_layout, nrmA, fa, \
                                                                       reinterpret_cast<Prop2*>(prop));             \
  } while (0)
      if (cls == PRE3_CLASS_DOUBLE) PRE3_FUSED(double);
      else PRE3_FUSED(float);
#undef PRE3_FUSED
      count_launch(ctx);
    }
    PRE3_TRY(pipe_enter(ctx, PS_MATCH));
    {
      Span span__(ctx, T_RESCORE);  // the pair bounds are an input of the rescore
      k_tc_pair_info<<<(P + 255) / 256, 256, 0, ctx->stream>>>(P, fa, fb, info, shared);
      count_launch(ctx);
    }
  } else
  {
    Span span__(ctx, T_CONVERT);
    const dim3 g1(K1p / CV_ROWS, FA), g2(K2p / CV_ROWS, P);
    if (cls == PRE3_CLASS_DOUBLE) {
      k_tc_convert<double><<<g1, 256, 0, ctx->stream>>>((const double*)dL1, K1, K1p, dk1, imgA, extA, ext_layout, nrmA, fa);
      if (!seq) k_tc_convert<double><<<g2, 256, 0, ctx->stream>>>((const double*)dL2, K2, K2p, dk2, imgB, extB, ext_layout, nrmB, fb);
    } else {
      k_tc_convert<float><<<g1, 256, 0, ctx->stream>>>((const float*)dL1, K1, K1p, dk1, imgA, extA, ext_layout, nrmA, fa);
      if (!seq) k_tc_convert<float><<<g2, 256, 0, ctx->stream>>>((const float*)dL2, K2, K2p, dk2, imgB, extB, ext_layout, nrmB, fb);
    }
    k_tc_pair_info<<<(P + 255) / 256, 256, 0, ctx->stream>>>(P, fa, fb, info, shared);
    count_launch(ctx, seq ? 2 : 3);
  }
  PRE3_TRY(pipe_enter(ctx, PS_MATCH));
  if (fused) {
    // proposals already written
  } else if (!use_v1) {
    // CTA pairs: one 2-CTA cluster per SM pair, unit = (pair, 256-row group)
    Span span__(ctx, T_MATCH_TC);
    const long long units = (long long)P * ((K1p / BLK + 3) / 4);
    const int grid = 2 * (int)std::min<long long>(units, ctx->sm_count / 2);
#define PRE3_GEMM2(E, EP)                                                                                            \
  do {                                                                                                             \
    static bool attr_done = false;                                                                                 \
    if (!attr_done) {                                                                                              \
      PRE3_CUDA(cudaFuncSetAttribute(k_tc_gemm_pair<E, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES)); \
      attr_done = true;                                                                                            \
    }                                                                                                              \
    k_tc_gemm_pair<E, EP><<<grid, THREADS, P2_SMEM_BYTES, ctx->stream>>>(imgA, imgB, extB, ext_layout, P, K1p, K2p, \
                                                                     reinterpret_cast<Prop2*>(prop), shared);      \
  } while (0)
    static const int epi_mode = getenv("PRE3_TC_EPI") ? atoi(getenv("PRE3_TC_EPI")) : 1;
    switch (exp_mode * 2 + (epi_mode ? 1 : 0)) {
      case 0: PRE3_GEMM2(0, 0); break;
      case 1: PRE3_GEMM2(0, 1); break;
      case 2: case 3: PRE3_GEMM2(1, 1); break;
      case 4: PRE3_GEMM2(2, 0); break;
      case 5: PRE3_GEMM2(2, 1); break;
      case 6: case 7: PRE3_GEMM2(3, 1); break;
      case 8: case 9: PRE3_GEMM2(4, 1); break;
      default: return fail(ctx, PRE3_ERR_ARG, "PRE3_TC_EXP: ablation not built for the CTA-pair kernel");
    }
#undef PRE3_GEMM2
    count_launch(ctx);
  } else {
    Span span__(ctx, T_MATCH_TC);
    const long long units = (long long)P * ((K1p / BLK + 1) / 2);
    const int grid = (int)std::min<long long>(units, ctx->sm_count);
#define PRE3_GEMM(E)                                                                                             \
  do {                                                                                                           \
    static bool attr_done = false;                                                                               \
    if (!attr_done) {                                                                                            \
      PRE3_CUDA(cudaFuncSetAttribute(k_tc_gemm_top2<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); \
      attr_done = true;                                                                                          \
    }                                                                                                            \
    k_tc_gemm_top2<E><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(imgA, imgB, nrmB, info, P, K1p, K2p, prop, shared);   \
  } while (0)
    switch (exp_mode) {
      case 0: PRE3_GEMM(0); break;
      case 1: PRE3_GEMM(1); break;
      case 2: PRE3_GEMM(2); break;
      case 4: PRE3_GEMM(4); break;
      case 5: PRE3_GEMM(5); break;
      case 6: PRE3_GEMM(6); break;
      case 7: PRE3_GEMM(7); break;
      default: return fail(ctx, PRE3_ERR_ARG, "PRE3_TC_EXP: ablation not built");
    }
#undef PRE3_GEMM
    count_launch(ctx);
  }
  const int v2 = use_v1 ? 0 : 1;
  static const int rescore_v1 = getenv("PRE3_RESCORE_V1") ? atoi(getenv("PRE3_RESCORE_V1")) : 0;
  if (v2 && !rescore_v1) {
    Span span__(ctx, T_RESCORE);
    const dim3 g((K1 + RS2_THREADS - 1) / RS2_THREADS, P);
#define PRE3_RS2(T, ACC)                                                                                              \
  do {                                                                                                                \
    static bool attr_done = false;                                                                                    \
    if (!attr_done) {                                                                                                 \
      PRE3_CUDA(cudaFuncSetAttribute(k_tc_rescore2<T, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                     (int)rs2_smem_bytes<ACC>()));                                                    \
      attr_done = true;                                                                                               \
    }                                                                                                                 \
    k_tc_rescore2<T, ACC><<<g, RS2_THREADS, rs2_smem_bytes<ACC>(), ctx->stream>>>(                                    \
        (const T*)dL1, (const T*)dL2, K1, K2, K1p, K2p, dk1, dk2, thresh, prop, nrmA, info, need_score, shared, drows, \
        list, list_n);                                                                                                \
  } while (0)
    if (cls == PRE3_CLASS_DOUBLE)
      PRE3_RS2(double, double);
    else if (cls == PRE3_CLASS_DOUBLE_F32)
      PRE3_RS2(float, double);
    else
      PRE3_RS2(float, float);
#undef PRE3_RS2
    count_launch(ctx);
  } else {
    Span span__(ctx, T_RESCORE);
    const dim3 g((K1 + RS_ROWS * RS_GROUPS - 1) / (RS_ROWS * RS_GROUPS), P);
    if (cls == PRE3_CLASS_DOUBLE)
      k_tc_rescore<double, double><<<g, 32, 0, ctx->stream>>>((const double*)dL1, (const double*)dL2, K1, K2,
                                                                          K1p, K2p, dk1, dk2, thresh, prop, nrmA, info,
                                                                          need_score, v2, shared, drows, list, list_n);
    else if (cls == PRE3_CLASS_DOUBLE_F32)
      k_tc_rescore<float, double><<<g, 32, 0, ctx->stream>>>((const float*)dL1, (const float*)dL2, K1, K2,
                                                                         K1p, K2p, dk1, dk2, thresh, prop, nrmA, info,
                                                                         need_score, v2, shared, drows, list, list_n);
    else
      k_tc_rescore<float, float><<<g, 32, 0, ctx->stream>>>((const float*)dL1, (const float*)dL2, K1, K2,
                                                                        K1p, K2p, dk1, dk2, thresh, prop, nrmA, info,
                                                                        need_score, v2, shared, drows, list, list_n);
    count_launch(ctx);
  }
  PRE3_CUDA(cudaGetLastError());
  if (getenv("PRE3_DEBUG") && atoi(getenv("PRE3_DEBUG")) >= 2) {  // diagnostics only: how many rows fall back to the exact kernel
    int32_t n = 0;
    PRE3_CUDA(cudaMemcpyAsync(&n, list_n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PRE3_CUDA(cudaStreamSynchronize(ctx->stream));
    fprintf(stderr, "[pre3] tc matcher: P=%d K1=%d K2=%d uncertified rows %d of %lld\n", P, K1, K2, n,
            (long long)P * K1);
  }
  // rows the proposal could not certify: exact brute force, one warp per row
  return launch_match_rows_exact(ctx, dL1, dL2, cls, K1, K2, ND, dk2, thresh, list, list_n, P * K1, drows);
}

}  // namespace pre3
