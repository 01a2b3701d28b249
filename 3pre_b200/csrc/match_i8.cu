// match_i8.cu -- stage 1 for class int8 / uint8 descriptors on the integer tensor cores (sm_100a, tcgen05 kind::i8).
//
// compare_mxINT8_CLASS / compare_mxUINT8_CLASS of M/sift/siftmatch.c:83-132 promote to `int` (:61-64): every squared
// distance is an exact integer below 2^24, so  d2(k1,k2) = |a|^2 + |b|^2 - 2 a.b  evaluated with an s32-accumulating
// GEMM IS the reference's value -- no proposal, no rescore, no error margins (`sift_demo2.m:93-96` matches
// uint8(512*descr); SIFT_extract_save stores uint8 descriptors the same way).
//   k_i8_convert    descriptors (128 x K column-major = K rows of 128 bytes) -> operand image in the shared-memory layout
//                   the UMMA descriptors expect (128-row blocks, ONE 128-byte swizzled row per descriptor), |x|^2 per
//                   descriptor and the per-column key constant (|b|^2 << 7) | (column & 127).
//   k_i8_gemm_pair  persistent CTA pairs, same pipeline as k_tc_gemm_pair (match_tc.cu): TMA producer warp, one-thread
//                   tcgen05.mma.cta_group::2.kind::i8 issuer (M256 x N128 x K32, four per accumulator tile), s32
//                   accumulators in tensor memory, two epilogue warpgroups per CTA.  Per accumulator the epilogue forms
//                   key = (|b|^2 - 2 a.b) * 128 + column  with one IMAD (FMA pipe) and keeps the two smallest keys with
//                   three integer min / max (ALU pipe); tiles are merged with strict `<` on the values so the first
//                   index wins ties exactly as :110-116.  The kernel writes the final MatchRow (best, bestk, Lowe's test
//                   with the float casts of :122-123).
#include <stdlib.h>

#include "match.cuh"
#include "tc_ptx.cuh"

namespace pre3 {

namespace i8 {

using namespace tc;

constexpr int ND = 128;
constexpr int BLK = 128;                  // rows per operand block
constexpr int BLK_BYTES = BLK * ND;       // 16 KB: [row 128][128 B], 128-byte swizzle
constexpr int TILE_N = 128;               // columns per B tile (64 per CTA)
constexpr int BHALF = 64 * ND;            // a CTA's 64 B rows: 8 KB
constexpr int NSTAGE = 8;                 // B ring
constexpr int THREADS = 352;              // warps 0-3 / 4-7 epilogue, 8 TMA, 9 MMA, 10 TMEM alloc
constexpr int W_EPI0 = 0, W_TMA = 8, W_MMA = 9, W_ALLOC = 10;
constexpr int OFF_A = 0;                  // 2 buffers x 2 row blocks x 16 KB
constexpr int OFF_B = 4 * BLK_BYTES;
constexpr int OFF_K = OFF_B + NSTAGE * BHALF;   // ring of key-constant tiles (128 x int32), same stage index as B
constexpr int KEY_BYTES = TILE_N * 4;
constexpr int OFF_BAR = OFF_K + NSTAGE * KEY_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 512;
constexpr int PADV = 0xFFFFFF;            // value of "no column": (INT_MAX >> 7); real values are < 2^24 / 2
constexpr int CV_ROWS = 32;

struct Bars {
  uint64_t a_full[2], a_empty[2], pa_full[2];
  uint64_t b_full[NSTAGE], b_empty[NSTAGE], pb_full[NSTAGE];
  uint64_t t_full[4], t_empty[4];
  uint64_t k_full[NSTAGE], k_empty[NSTAGE];  // key tiles: local to each CTA (8 epilogue-warp arrivals free a stage)
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 512, "Bars");

// One launch converts gridDim.y descriptor sets; 8 threads per row (16 bytes each).
template <bool SIGNED>
__global__ void __launch_bounds__(256)
k_i8_convert(const unsigned char* __restrict__ L, int K, int Kp, const int32_t* __restrict__ kc,
             unsigned char* __restrict__ img, int32_t* __restrict__ nrm, int32_t* __restrict__ ckey) {
  const int p = blockIdx.y;
  const int n = kc ? max(0, min(kc[p], K)) : K;
  const int row = blockIdx.x * CV_ROWS + (threadIdx.x >> 3), chunk = threadIdx.x & 7;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (row < n) v = __ldg(reinterpret_cast<const uint4*>(L + ((size_t)p * K + row) * ND) + chunk);
  int sq = 0;
  if (SIGNED) {
    sq = __dp4a((int)v.x, (int)v.x, sq);
    sq = __dp4a((int)v.y, (int)v.y, sq);
    sq = __dp4a((int)v.z, (int)v.z, sq);
    sq = __dp4a((int)v.w, (int)v.w, sq);
  } else {
    unsigned u = 0;
    u = __dp4a(v.x, v.x, u);
    u = __dp4a(v.y, v.y, u);
    u = __dp4a(v.z, v.z, u);
    u = __dp4a(v.w, v.w, u);
    sq = (int)u;
  }
  sq += __shfl_xor_sync(0xffffffffu, sq, 1);
  sq += __shfl_xor_sync(0xffffffffu, sq, 2);
  sq += __shfl_xor_sync(0xffffffffu, sq, 4);
  if (row >= Kp) return;
  const int rb = row >> 7, r = row & 127;
  *reinterpret_cast<uint4*>(img + (size_t)p * Kp * ND + (size_t)rb * BLK_BYTES + (size_t)r * 128 +
                            (size_t)((chunk ^ (r & 7)) << 4)) = v;
  if (chunk == 0) {
    nrm[(size_t)p * Kp + row] = row < n ? sq : 0;
    ckey[(size_t)p * Kp + row] = row < n ? ((sq << 7) | r) : (int)((unsigned)PADV << 7 | (unsigned)r);
  }
}

__device__ __forceinline__ void tc_mma_i8_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// unit = (pair, 512-row group of L1); CTA `rank` of the pair holds the row blocks 4g + 2s + rank (s = 0, 1) and half
// of every B tile.  idesc selects the signedness of A and B (kind::i8, D = S32, M = 256, N = 128, both K-major).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_i8_gemm_pair(const unsigned char* __restrict__ imgA, const unsigned char* __restrict__ imgB,
               const int32_t* __restrict__ nrmA, const int32_t* __restrict__ ckeyB, uint32_t idesc, int P, int K1,
               int K1p, int K2p, const int32_t* __restrict__ k1c, float thresh, MatchRow* __restrict__ rows,
               int a_shared) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t sA = base + OFF_A, sB = base + OFF_B;
  Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int nblk = K1p / BLK;               // K1p, K2p are multiples of 256
  const int groups = (nblk + 3) / 4;
  const int ntile = K2p / TILE_N;
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
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(smem_u32(&bars->b_full[i]), 1);
      mbar_init(smem_u32(&bars->b_empty[i]), 1);
      mbar_init(smem_u32(&bars->pb_full[i]), 1);
      mbar_init(smem_u32(&bars->k_full[i]), 1);
      mbar_init(smem_u32(&bars->k_empty[i]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ALLOC) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars->tmem_base))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == W_TMA) {
    if (lane == 0) {
      long long t = 0;
      int uc = 0;
      for (long long u = cluster_id; u < nunits; u += nclusters, ++uc) {
        const int p = (int)(u / groups), g = (int)(u % groups);
        const int ns = min(2, (nblk - 4 * g) / 2);
        const int ab = uc & 1;
        mbar_wait(smem_u32(&bars->a_empty[ab]), ((uc >> 1) & 1) ^ 1);
        mbar_expect_tx(smem_u32(&bars->a_full[ab]), (uint32_t)(ns * BLK_BYTES));
        for (int sidx = 0; sidx < ns; ++sidx)
          tma_bulk_g2s(sA + (ab * 2 + sidx) * BLK_BYTES,
                       imgA + ((size_t)(a_shared ? 0 : p) * K1p + (size_t)(4 * g + 2 * sidx + (int)rank) * BLK) * ND,
                       (uint32_t)BLK_BYTES, smem_u32(&bars->a_full[ab]));
        for (int j = 0; j < ntile; ++j, ++t) {
          const int st = (int)(t % NSTAGE);
          mbar_wait(smem_u32(&bars->b_empty[st]), (uint32_t)(((t / NSTAGE) & 1) ^ 1));
          mbar_expect_tx(smem_u32(&bars->b_full[st]), (uint32_t)BHALF);
          tma_bulk_g2s(sB + st * BHALF, imgB + ((size_t)p * K2p + (size_t)j * BLK) * ND + (size_t)rank * BHALF, BHALF,
                       smem_u32(&bars->b_full[st]));
          // the tile's 128 key constants, read by the epilogue long after the MMA has released the B stage
          mbar_wait(smem_u32(&bars->k_empty[st]), (uint32_t)(((t / NSTAGE) & 1) ^ 1));
          mbar_expect_tx(smem_u32(&bars->k_full[st]), (uint32_t)KEY_BYTES);
          tma_bulk_g2s(base + OFF_K + st * KEY_BYTES, ckeyB + (size_t)p * K2p + (size_t)j * TILE_N, KEY_BYTES,
                       smem_u32(&bars->k_full[st]));
        }
      }
      // the multicast commits that release the last buffers target this CTA's barriers too: stay until they have arrived
      for (int i = 0; i < NSTAGE; ++i, ++t)
        mbar_wait(smem_u32(&bars->b_empty[t % NSTAGE]), (uint32_t)(((t / NSTAGE) & 1) ^ 1));
      for (int i = 0; i < 2; ++i, ++uc) mbar_wait(smem_u32(&bars->a_empty[uc & 1]), ((uc >> 1) & 1) ^ 1);
    }
  } else if (warp == W_MMA) {
    if (lane == 0) {
      long long t = 0;
      int uc = 0;
      if (leader) {
        uint32_t use00 = 0, use01 = 0, use10 = 0, use11 = 0;
        for (long long u = cluster_id; u < nunits; u += nclusters, ++uc) {
          const int g = (int)(u % groups);
          const int ns = min(2, (nblk - 4 * g) / 2);
          const int ab = uc & 1;
          mbar_wait(smem_u32(&bars->a_full[ab]), (uc >> 1) & 1);
          mbar_wait_cl(smem_u32(&bars->pa_full[ab]), (uc >> 1) & 1);
          for (int j = 0; j < ntile; ++j, ++t) {
            const int st = (int)(t % NSTAGE);
            const uint32_t ph = (uint32_t)((t / NSTAGE) & 1);
            mbar_wait(smem_u32(&bars->b_full[st]), ph);
            mbar_wait_cl(smem_u32(&bars->pb_full[st]), ph);
            tc_fence_after();
            const uint32_t sBst = sB + st * BHALF;
#pragma unroll
            for (int sidx = 0; sidx < 2; ++sidx) {
              if (sidx >= ns) break;
              const int slot = (int)(t & 1) * 2 + sidx;
              uint32_t& use = (t & 1) ? (sidx ? use11 : use10) : (sidx ? use01 : use00);
              mbar_wait_cl(smem_u32(&bars->t_empty[slot]), (use & 1) ^ 1);
              ++use;
              tc_fence_after();
              const uint32_t d = tmem + (uint32_t)(slot * TILE_N);
#pragma unroll
              for (int k = 0; k < ND / 32; ++k)
                tc_mma_i8_2cta(d, umma_desc(sA + (ab * 2 + sidx) * BLK_BYTES + k * 32), umma_desc(sBst + k * 32), idesc,
                               k > 0 ? 1u : 0u);
              tc_commit_mc2(smem_u32(&bars->t_full[slot]));
            }
            tc_commit_mc2(smem_u32(&bars->b_empty[st]));
          }
          tc_commit_mc2(smem_u32(&bars->a_empty[ab]));
        }
      } else {
        for (long long u = cluster_id; u < nunits; u += nclusters, ++uc) {
          const int ab = uc & 1;
          mbar_wait(smem_u32(&bars->a_full[ab]), (uc >> 1) & 1);
          mbar_arrive_cluster(mapa_u32(smem_u32(&bars->pa_full[ab]), 0));
          for (int j = 0; j < ntile; ++j, ++t) {
            const int st = (int)(t % NSTAGE);
            mbar_wait(smem_u32(&bars->b_full[st]), (uint32_t)((t / NSTAGE) & 1));
            mbar_arrive_cluster(mapa_u32(smem_u32(&bars->pb_full[st]), 0));
          }
        }
      }
    }
  } else if (warp < W_TMA) {
    // ===== epilogue: warpgroup s owns A tile s of this CTA; thread = one L1 row =====================
    const int sidx = (warp - W_EPI0) >> 2;
    const int q = warp & 3;
    const uint32_t tel0 = mapa_u32(smem_u32(&bars->t_empty[sidx]), 0);
    const uint32_t tel1 = mapa_u32(smem_u32(&bars->t_empty[2 + sidx]), 0);
    long long t = 0;
    uint32_t use0 = 0, use1 = 0;
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(sidx * TILE_N);
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
    for (long long u = cluster_id; u < nunits; u += nclusters) {
      const int p = (int)(u / groups), g = (int)(u % groups);
      const int ns = min(2, (nblk - 4 * g) / 2);
      if (sidx >= ns) {
        t += ntile;
        continue;
      }
      bool next_unit = false;
      for (long long v = u + nclusters; v < nunits; v += nclusters)
        if (sidx < min(2, (nblk - 4 * (int)(v % groups)) / 2)) {
          next_unit = true;
          break;
        }
      int gb = PADV, gs = PADV, gidx = -1;  // running best / second VALUES (|b|^2 - 2 a.b) and the best's column
      for (int j = 0; j < ntile; ++j, ++t) {
        const int stg = (int)(t & 1);
        if (!primed) {
          wait_full(stg);
          tc_ld_32x32(tbase + (uint32_t)(stg * 2 * TILE_N), buf[0]);
          primed = true;
        }
        const bool has_next = (j + 1 < ntile) || next_unit;
        int a1 = 0x7fffffff, a2 = 0x7fffffff, b1 = 0x7fffffff, b2 = 0x7fffffff;  // two chains: even / odd columns
        const int kst = (int)(t % NSTAGE);
        mbar_wait(smem_u32(&bars->k_full[kst]), (uint32_t)((t / NSTAGE) & 1));
        const int4* ck = reinterpret_cast<const int4*>(smem + OFF_K + kst * KEY_BYTES);
#pragma unroll
        for (int c = 0; c < TILE_N / 32; ++c) {
          tc_ld_wait(buf[c & 1]);
          if (c + 1 < TILE_N / 32) {
            tc_ld_32x32(tbase + (uint32_t)(stg * 2 * TILE_N + (c + 1) * 32), buf[(c + 1) & 1]);
          } else {
            release(stg);
            if (has_next) {
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
              tc_ld_32x32(tbase + (uint32_t)(nstg * 2 * TILE_N), buf[0]);
            }
          }
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const int4 cv = ck[c * 8 + i4];  // warp-uniform address: one broadcast
            const int k0 = (int)buf[c & 1][4 * i4 + 0] * -256 + cv.x;
            const int k1 = (int)buf[c & 1][4 * i4 + 1] * -256 + cv.y;
            const int k2 = (int)buf[c & 1][4 * i4 + 2] * -256 + cv.z;
            const int k3 = (int)buf[c & 1][4 * i4 + 3] * -256 + cv.w;
            int x;
            x = max(a1, k0), a1 = min(a1, k0), a2 = min(a2, x);
            x = max(b1, k1), b1 = min(b1, k1), b2 = min(b2, x);
            x = max(a1, k2), a1 = min(a1, k2), a2 = min(a2, x);
            x = max(b1, k3), b1 = min(b1, k3), b2 = min(b2, x);
          }
        }
        __syncwarp();
        if (lane == 0) {  // key tile read; a unit with one A tile per CTA has no second warpgroup: arrive for it too
          mbar_arrive(smem_u32(&bars->k_empty[kst]));
          if (ns == 1) mbar_arrive(smem_u32(&bars->k_empty[kst]));
        }
        // the tile's two smallest keys, then into the running state with strict '<' on the VALUES (earlier tile wins ties)
        const int tb = min(a1, b1), ts = min(max(a1, b1), min(a2, b2));
        const int tv = tb >> 7, tsv = ts >> 7;
        if (tv < gb) {
          gs = min(gb, tsv);
          gb = tv;
          gidx = j * TILE_N + (tb & 127);
        } else {
          gs = min(gs, tv);
        }
      }
      const int row = (4 * g + 2 * sidx + (int)rank) * BLK + q * 32 + lane;
      const int pa = a_shared ? 0 : p;
      const int n1 = k1c ? min(k1c[pa], K1) : K1;
      if (row < n1) {
        const int na = nrmA[(size_t)pa * K1p + row];
        MatchRow r;
        const int best = gb >= PADV ? 0x7fffffff : na + gb;
        const int second = gs >= PADV ? 0x7fffffff : na + gs;
        r.best = (double)best;
        r.bestk = gb >= PADV ? -1 : gidx;
        r.accept = (__fmul_rn(thresh, (float)best) <= (float)second && r.bestk != -1) ? 1 : 0;  // siftmatch.c:122-123
        rows[(size_t)p * K1 + row] = r;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == W_ALLOC) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace i8

static inline int pad256(int k) { return (k + 255) / 256 * 256; }

bool match_i8_supported(int cls, int K1, int K2, int ND) {
  return (cls == PRE3_CLASS_INT8 || cls == PRE3_CLASS_UINT8) && ND == i8::ND && K1 >= 1 && K2 >= 1;
}

size_t match_i8_workspace_bytes(int P, int K1, int K2) {
  const size_t K1p = pad256(K1), K2p = pad256(K2);
  size_t b = 0;
  b += align_up((size_t)(P + 1) * K1p * i8::ND, 1024) + 1024;
  b += align_up((size_t)P * K2p * i8::ND, 1024) + 1024;
  b += 2 * (align_up((size_t)(P + 1) * K1p * 4) + align_up((size_t)P * K2p * 4));
  return b + 4096;
}

int launch_match_i8(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int P, int K1, int K2, int ND,
                    const int32_t* dk1, const int32_t* dk2, float thresh, MatchRow* drows) {
  using namespace i8;
  if (!match_i8_supported(cls, K1, K2, ND)) return fail(ctx, PRE3_ERR_ARG, "integer tensor-core matcher: unsupported shape");
  if (P <= 0) return PRE3_OK;
  const int K1p = pad256(K1), K2p = pad256(K2);
  const bool seq = dL2 == nullptr;  // P + 1 consecutive sets behind dL1, pair p = (set p, set p + 1)
  if (seq && K1 != K2) return fail(ctx, PRE3_ERR_ARG, "sequence mode needs the same descriptor count per frame");
  const int shared = (!seq && ctx->l1_shared) ? 1 : 0;
  const int FA = seq ? P + 1 : (shared ? 1 : P);
  auto take1k = [&](size_t bytes) {
    ctx->ws_off = align_up(ctx->ws_off, 1024);
    return ws_take<unsigned char>(ctx, bytes);
  };
  const size_t set_bytes = (size_t)K1p * ND;
  unsigned char* imgA = take1k((size_t)FA * set_bytes);
  unsigned char* imgB = seq ? imgA + set_bytes : take1k((size_t)P * K2p * ND);
  int32_t* nrmA = ws_take<int32_t>(ctx, (size_t)FA * K1p);
  int32_t* ckA = ws_take<int32_t>(ctx, (size_t)FA * K1p);
  int32_t* nrmB = seq ? nrmA + K1p : ws_take<int32_t>(ctx, (size_t)P * K2p);
  int32_t* ckB = seq ? ckA + K1p : ws_take<int32_t>(ctx, (size_t)P * K2p);
  (void)nrmB;
  if (seq) {
    dL2 = (const char*)dL1 + (size_t)K1 * ND;
    if (dk1) dk2 = dk1 + 1;
  }
  const bool sgn = cls == PRE3_CLASS_INT8;
  PRE3_TRY(pipe_enter(ctx, PS_CONVERT));
  {
    Span span__(ctx, T_CONVERT);
    const dim3 g1(K1p / CV_ROWS, FA), g2(K2p / CV_ROWS, P);
    if (sgn) {
      k_i8_convert<true><<<g1, 256, 0, ctx->stream>>>((const unsigned char*)dL1, K1, K1p, dk1, imgA, nrmA, ckA);
      if (!seq) k_i8_convert<true><<<g2, 256, 0, ctx->stream>>>((const unsigned char*)dL2, K2, K2p, dk2, imgB, nrmB, ckB);
    } else {
      k_i8_convert<false><<<g1, 256, 0, ctx->stream>>>((const unsigned char*)dL1, K1, K1p, dk1, imgA, nrmA, ckA);
      if (!seq) k_i8_convert<false><<<g2, 256, 0, ctx->stream>>>((const unsigned char*)dL2, K2, K2p, dk2, imgB, nrmB, ckB);
    }
    count_launch(ctx, seq ? 1 : 2);
  }
  PRE3_TRY(pipe_enter(ctx, PS_MATCH));
  {
    Span span__(ctx, T_MATCH_TC);
    const long long units = (long long)P * ((K1p / BLK + 3) / 4);
    const int grid = 2 * (int)std::min<long long>(units, ctx->sm_count / 2);
    static bool attr_done = false;
    if (!attr_done) {
      PRE3_CUDA(cudaFuncSetAttribute(k_i8_gemm_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_done = true;
    }
    // kind::i8: D = S32 (bits 4-5 = 2), A / B format 0 = unsigned, 1 = signed (bits 7-9 / 10-12), K-major, N, M = 256
    const uint32_t fmt = sgn ? 1u : 0u;
    const uint32_t idesc = (2u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    k_i8_gemm_pair<<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(imgA, imgB, nrmA, ckB, idesc, P, K1, K1p, K2p, dk1, thresh,
                                                               drows, shared);
    count_launch(ctx);
  }
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

}  // namespace pre3
