// match_exact.cu -- stage 1 in the reference's own arithmetic, brute force.
//
// compare_mx*_CLASS of M/sift/siftmatch.c:83-132: for every column of L1 the squared L2
// distance to every column of L2, accumulated bin by bin in the promoted type (:101-107,
// double / float / int), best + second best with strict '<' (:110-116, first index wins),
// Lowe's test with float casts (:122-123).  This kernel is (a) the engine for int8 / uint8
// and for ND != 128, (b) the exact recomputation for rows the tensor-core proposal path
// cannot certify.  Compile with -fmad=false: delta*delta and the add must not fuse, the
// reference (gcc, x86-64, no -mfma) rounds twice.
#include "match.cuh"

namespace pre3 {

constexpr int MT = 32;   // block tile: 32 L1 columns x 32 L2 columns
constexpr int MB = 32;   // bins per shared-memory chunk

template <typename ACC>
struct Top2 {
  ACC best, second;
  int bestk;
};

template <typename ACC>
__device__ __forceinline__ ACC maxval();
template <>
__device__ __forceinline__ double maxval<double>() { return INFINITY; }
template <>
__device__ __forceinline__ float maxval<float>() { return INFINITY; }
template <>
__device__ __forceinline__ int maxval<int>() { return 0x7fffffff; }

template <typename ACC>
__device__ __forceinline__ void top2_update(Top2<ACC>& s, ACC acc, int k2) {  // siftmatch.c:110-116
  if (acc < s.best) {
    s.second = s.best;
    s.best = acc;
    s.bestk = k2;
  } else if (acc < s.second) {
    s.second = acc;
  }
}

// Order-independent merge of two partial states over disjoint column sets: the sequential
// rule yields (min, first index of the min, second smallest with multiplicity).
template <typename ACC>
__device__ __forceinline__ void top2_merge(Top2<ACC>& a, const Top2<ACC>& b) {
  const bool a_first = (a.best < b.best) || (a.best == b.best && (unsigned)a.bestk < (unsigned)b.bestk);
  if (a_first) {
    a.second = b.best < a.second ? b.best : a.second;
  } else {
    const ACC s = a.best < b.second ? a.best : b.second;
    a.best = b.best;
    a.bestk = b.bestk;
    a.second = s;
  }
}

template <typename T, typename ACC>
__global__ void __launch_bounds__(256)
k_match_exact(const T* __restrict__ L1, const T* __restrict__ L2, int K1, int K2, int ND,
              const int32_t* __restrict__ k1c, const int32_t* __restrict__ k2c, float thresh,
              const int32_t* __restrict__ row_list, const int32_t* __restrict__ row_list_n,
              MatchRow* __restrict__ rows, int a_shared) {
  __shared__ T sa[MT][MB + 1];
  __shared__ T sb[MT][MB + 1];
  const int p = blockIdx.y;
  const int pa = a_shared ? 0 : p;  // one L1 set for every problem (find_consistent_sift_matches.m:39-65)
  const int n1 = k1c ? min(k1c[pa], K1) : K1;
  const int n2 = k2c ? min(k2c[p], K2) : K2;
  const int row0 = blockIdx.x * MT;
  if (row0 >= n1) return;
  const T* A = L1 + (size_t)pa * K1 * ND;
  const T* B = L2 + (size_t)p * K2 * ND;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  Top2<ACC> st[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    st[i].best = maxval<ACC>();
    st[i].second = maxval<ACC>();
    st[i].bestk = -1;
  }
  for (int col0 = 0; col0 < n2; col0 += MT) {
    ACC acc[2][2] = {{0, 0}, {0, 0}};
    for (int b0 = 0; b0 < ND; b0 += MB) {
      __syncthreads();
      for (int e = threadIdx.x; e < MT * MB; e += 256) {
        const int r = e / MB, c = e % MB;
        const int bin = b0 + c;
        T va = 0, vb = 0;
        if (bin < ND) {
          if (row0 + r < n1) va = A[(size_t)(row0 + r) * ND + bin];
          if (col0 + r < n2) vb = B[(size_t)(col0 + r) * ND + bin];
        }
        sa[r][c] = va;
        sb[r][c] = vb;
      }
      __syncthreads();
      const int nb = min(MB, ND - b0);
      for (int c = 0; c < nb; ++c) {
        const ACC a0 = (ACC)sa[2 * ty][c], a1 = (ACC)sa[2 * ty + 1][c];
        const ACC c0 = (ACC)sb[2 * tx][c], c1 = (ACC)sb[2 * tx + 1][c];
        ACC d;
        d = a0 - c0;
        acc[0][0] += d * d;
        d = a0 - c1;
        acc[0][1] += d * d;
        d = a1 - c0;
        acc[1][0] += d * d;
        d = a1 - c1;
        acc[1][1] += d * d;
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k2 = col0 + 2 * tx + j;
        if (k2 < n2) top2_update(st[i], acc[i][j], k2);
      }
  }
  // merge across the 16 threads (tx) that share a row pair: xor shuffles stay in the half-warp
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
      Top2<ACC> o;
      o.best = __shfl_xor_sync(0xffffffffu, st[i].best, off);
      o.second = __shfl_xor_sync(0xffffffffu, st[i].second, off);
      o.bestk = __shfl_xor_sync(0xffffffffu, st[i].bestk, off);
      top2_merge(st[i], o);
    }
    const int k1 = row0 + 2 * ty + i;
    if (tx == 0 && k1 < n1) {
      MatchRow r;
      r.best = (double)st[i].best;
      r.bestk = st[i].bestk;
      // thresh * (float) best <= (float) second_best && bestk != -1   (siftmatch.c:122-123)
      r.accept = (__fmul_rn(thresh, (float)st[i].best) <= (float)st[i].second && st[i].bestk != -1) ? 1 : 0;
      rows[(size_t)p * K1 + k1] = r;
    }
  }
  (void)row_list;
  (void)row_list_n;
}

constexpr int RX_ND = 128;
// Exact recomputation of the rows the proposal path flagged (row_list holds p*K1 + k1 entries, *row_list_n of
// them), one BLOCK per row: thread = column (strided by the block size), 128 sequential `acc += delta*delta`
// per column in registers (strictly in bin order, siftmatch.c:101-107), the A row in shared memory.  The rows the proposal
// cannot certify are few (a handful per thousand pairs), so what matters is the latency of ONE row: a warp per row
// needs ~100 us for 512 columns, a 256-thread block a few.
constexpr int RXB_THREADS = 128;  // 4 warps
constexpr int RXB_COLS = 32;      // columns per warp pass

template <typename T>
__device__ __forceinline__ void load_vec4(const T* src, T* v);
template <>
__device__ __forceinline__ void load_vec4<double>(const double* src, double* v) {
  const double2 x = __ldg(reinterpret_cast<const double2*>(src));
  const double2 y = __ldg(reinterpret_cast<const double2*>(src) + 1);
  v[0] = x.x, v[1] = x.y, v[2] = y.x, v[3] = y.y;
}
template <>
__device__ __forceinline__ void load_vec4<float>(const float* src, float* v) {
  const float4 x = __ldg(reinterpret_cast<const float4*>(src));
  v[0] = x.x, v[1] = x.y, v[2] = x.z, v[3] = x.w;
}

// A warp takes 32 columns at a time: every column is read as ONE coalesced row (lane = 4 consecutive bins), the
// products delta*delta go to shared memory, then lane L sums column L strictly in bin order (row stride 129: no bank
// conflicts).  The first version walked one column per THREAD (32 lanes x 1 KB apart per load): 47 us for a handful of
// rows, all of it exposed L2 latency (ncu r02_p: long scoreboard 49 warps per issue).
template <typename T, typename ACC>
__global__ void __launch_bounds__(RXB_THREADS)
k_match_rows_exact_blk(const T* __restrict__ L1, const T* __restrict__ L2, int K1, int K2, int ND,
                       const int32_t* __restrict__ k2c, float thresh, const int32_t* __restrict__ row_list,
                       const int32_t* __restrict__ row_list_n, int list_cap, MatchRow* __restrict__ rows,
                       int a_shared) {
  extern __shared__ __align__(16) unsigned char rx_smem[];
  ACC* sa = reinterpret_cast<ACC*>(rx_smem);                                 // the L1 row
  ACC(*sprod)[RX_ND + 1] = reinterpret_cast<ACC(*)[RX_ND + 1]>(sa + RX_ND);  // [warp * 32 + column][bin]
  __shared__ Top2<ACC> sred[RXB_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = min(*row_list_n, list_cap);
  for (int w = blockIdx.x; w < n; w += gridDim.x) {
    const int rid = row_list[w];
    const int p = rid / K1;
    const int n2 = k2c ? min(k2c[p], K2) : K2;
    const T* a = L1 + (size_t)(a_shared ? rid % K1 : rid) * ND;
    const T* B = L2 + (size_t)p * K2 * ND;
    __syncthreads();  // sa / sred of the previous row are no longer read
    for (int e = tid; e < RX_ND; e += RXB_THREADS) sa[e] = (ACC)a[e];
    __syncthreads();
    ACC av[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) av[e] = sa[4 * lane + e];
    Top2<ACC> st;
    st.best = maxval<ACC>();
    st.second = maxval<ACC>();
    st.bestk = -1;
    ACC(*mine)[RX_ND + 1] = sprod + warp * RXB_COLS;
    for (int c0 = warp * RXB_COLS; c0 < n2; c0 += (RXB_THREADS / 32) * RXB_COLS) {
      const int nc = min(RXB_COLS, n2 - c0);
      for (int j0 = 0; j0 < nc; j0 += 16) {  // sixteen columns' loads in flight together: a row is a latency chain
        T bv[16][4];
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (j0 + u < nc) load_vec4<T>(B + (size_t)(c0 + j0 + u) * ND + 4 * lane, bv[u]);
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (j0 + u < nc) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const ACC d = av[e] - (ACC)bv[u][e];
              mine[j0 + u][4 * lane + e] = d * d;
            }
          }
      }
      __syncwarp();
      if (lane < nc) {
        ACC acc = 0;
#pragma unroll 16
        for (int e = 0; e < RX_ND; ++e) acc += mine[lane][e];  // strictly in bin order (siftmatch.c:101-107)
        top2_update(st, acc, c0 + lane);  // ascending columns per lane: the first minimum is kept
      }
      __syncwarp();
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      Top2<ACC> o;
      o.best = __shfl_xor_sync(0xffffffffu, st.best, off);
      o.second = __shfl_xor_sync(0xffffffffu, st.second, off);
      o.bestk = __shfl_xor_sync(0xffffffffu, st.bestk, off);
      top2_merge(st, o);
    }
    if (lane == 0) sred[warp] = st;
    __syncthreads();
    if (tid == 0) {
      Top2<ACC> r = sred[0];
      for (int i = 1; i < RXB_THREADS / 32; ++i) top2_merge(r, sred[i]);
      MatchRow o;
      o.best = (double)r.best;
      o.bestk = r.bestk;
      o.accept = (__fmul_rn(thresh, (float)r.best) <= (float)r.second && r.bestk != -1) ? 1 : 0;
      rows[rid] = o;
    }
  }
}

template <typename ACC>
constexpr size_t rxb_smem_bytes() {
  return sizeof(ACC) * (RX_ND + (size_t)(RXB_THREADS / 32) * RXB_COLS * (RX_ND + 1));
}

// rows -> compact list in k1 order + gathered correspondences.  One block per pair.
__global__ void __launch_bounds__(256)
k_match_compact(const MatchRow* __restrict__ rows, int K1, const int32_t* __restrict__ k1c,
                int32_t* __restrict__ pairs, double* __restrict__ score, int32_t* __restrict__ n_out,
                const double* __restrict__ xyz1, const double* __restrict__ xyz2, int K2,
                double* __restrict__ Ya, double* __restrict__ Yb) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int p = blockIdx.x;
  const int n1 = k1c ? min(k1c[p], K1) : K1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int r0 = 0; r0 < n1; r0 += 256) {
    const int k1 = r0 + threadIdx.x;
    MatchRow r;
    r.accept = 0;
    if (k1 < n1) r = rows[(size_t)p * K1 + k1];
    const unsigned bal = __ballot_sync(0xffffffffu, r.accept != 0);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    off += __popc(bal & ((1u << lane) - 1u));
    if (r.accept) {
      if (pairs) {
        pairs[((size_t)p * K1 + off) * 2] = k1;
        pairs[((size_t)p * K1 + off) * 2 + 1] = r.bestk;
      }
      if (score) score[(size_t)p * K1 + off] = r.best;
      if (Ya) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          Ya[((size_t)p * K1 + off) * 3 + c] = xyz1[((size_t)p * K1 + k1) * 3 + c];
          Yb[((size_t)p * K1 + off) * 3 + c] = xyz2[((size_t)p * K2 + r.bestk) * 3 + c];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = s_base;
      for (int w = 0; w < 8; ++w) t += s_warp[w];
      s_base = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) n_out[p] = s_base;
}

size_t match_workspace_bytes(int cls, int P, int K1, int K2, int ND) {
  size_t b = align_up(sizeof(MatchRow) * (size_t)P * K1) + 4096;
  if (match_tc_supported(cls, K1, K2, ND)) b += match_tc_workspace_bytes(P, K1, K2);
  if (match_i8_supported(cls, K1, K2, ND)) b += match_i8_workspace_bytes(P, K1, K2);
  return b;
}

int launch_match_exact(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int P, int K1, int K2, int ND,
                       const int32_t* dk1, const int32_t* dk2, float thresh, MatchRow* drows) {
  Span span__(ctx, T_MATCH_EXACT);
  if (P <= 0 || K1 <= 0) return PRE3_OK;
  dim3 grid((K1 + MT - 1) / MT, P);
  switch (cls) {
    case PRE3_CLASS_DOUBLE:
      k_match_exact<double, double><<<grid, 256, 0, ctx->stream>>>((const double*)dL1, (const double*)dL2, K1, K2, ND,
                                                                    dk1, dk2, thresh, nullptr, nullptr, drows, ctx->l1_shared);
      break;
    case PRE3_CLASS_SINGLE:
      k_match_exact<float, float><<<grid, 256, 0, ctx->stream>>>((const float*)dL1, (const float*)dL2, K1, K2, ND, dk1,
                                                                  dk2, thresh, nullptr, nullptr, drows, ctx->l1_shared);
      break;
    case PRE3_CLASS_DOUBLE_F32:  // class double stored as float: double arithmetic
      k_match_exact<float, double><<<grid, 256, 0, ctx->stream>>>((const float*)dL1, (const float*)dL2, K1, K2, ND, dk1,
                                                                   dk2, thresh, nullptr, nullptr, drows, ctx->l1_shared);
      break;
    case PRE3_CLASS_INT8:
      k_match_exact<signed char, int><<<grid, 256, 0, ctx->stream>>>((const signed char*)dL1, (const signed char*)dL2,
                                                                      K1, K2, ND, dk1, dk2, thresh, nullptr, nullptr,
                                                                      drows, ctx->l1_shared);
      break;
    case PRE3_CLASS_UINT8:
      k_match_exact<unsigned char, int><<<grid, 256, 0, ctx->stream>>>((const unsigned char*)dL1,
                                                                        (const unsigned char*)dL2, K1, K2, ND, dk1,
                                                                        dk2, thresh, nullptr, nullptr, drows, ctx->l1_shared);
      break;
    default:
      return fail(ctx, PRE3_ERR_CLASS, "Unsupported numeric class");
  }
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_match_rows_exact(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int K1, int K2, int ND,
                            const int32_t* dk2, float thresh, const int32_t* drow_list, const int32_t* drow_list_n,
                            int list_cap, MatchRow* drows) {
  Span span__(ctx, T_RESCORE);
  if (ND != RX_ND) return fail(ctx, PRE3_ERR_ARG, "row recheck: ND must be 128");
  const int blocks = 2 * ctx->sm_count;
  // rows of L1 / L2 are read with 16-byte vector loads
  if ((((uintptr_t)dL1 | (uintptr_t)dL2) & 15u) != 0) return fail(ctx, PRE3_ERR_ARG, "row recheck: descriptors must be 16-byte aligned");
#define PRE3_RXB(T, ACC)                                                                                               \
  do {                                                                                                                 \
    static bool attr_done = false;                                                                                     \
    if (!attr_done) {                                                                                                  \
      PRE3_CUDA(cudaFuncSetAttribute(k_match_rows_exact_blk<T, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                     (int)rxb_smem_bytes<ACC>()));                                                     \
      attr_done = true;                                                                                                \
    }                                                                                                                  \
    k_match_rows_exact_blk<T, ACC><<<blocks, RXB_THREADS, rxb_smem_bytes<ACC>(), ctx->stream>>>(                       \
        (const T*)dL1, (const T*)dL2, K1, K2, ND, dk2, thresh, drow_list, drow_list_n, list_cap, drows, ctx->l1_shared); \
  } while (0)
  if (cls == PRE3_CLASS_DOUBLE)
    PRE3_RXB(double, double);
  else if (cls == PRE3_CLASS_SINGLE)
    PRE3_RXB(float, float);
  else if (cls == PRE3_CLASS_DOUBLE_F32)
    PRE3_RXB(float, double);
  else
    return fail(ctx, PRE3_ERR_CLASS, "row recheck: class must be double or single");
#undef PRE3_RXB
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

int launch_match_compact(pre3_ctx* ctx, const MatchRow* drows, int P, int K1, const int32_t* dk1,
                         int32_t* dpairs, double* dscore, int32_t* dn_out, const double* dxyz1,
                         const double* dxyz2, int K2, double* dYa, double* dYb) {
  Span span__(ctx, T_COMPACT);
  if (P <= 0) return PRE3_OK;
  k_match_compact<<<P, 256, 0, ctx->stream>>>(drows, K1, dk1, dpairs, dscore, dn_out, dxyz1, dxyz2, K2, dYa, dYb);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

// ------------------------------------------------------------------------------------------
// Search-region gate of the EKF's SIFT matcher (M/matching_sift_based.m:118-135), as a post-filter over the compacted
// match list of each problem: match i = (k1, k2) is accepted as individually compatible iff
//     norm(pos2(:,k2) - h(:,k1)) <= half_search_region_size_x,
// half_search_region_size_x = ceil(3*sqrt(S(1,1))) of the i-th PREDICTED feature (the reference indexes S with the
// loop counter over the matches, `features_info(index_in_info(i)).S`, :120 -- reproduced), 40 when that S is empty
// (NaN here).  Accepted: ic(k1) = 1, z(:,k1) = pos2(:,k2), match(k1) = k2; the others are counted as discarded.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_match_gate(const int32_t* __restrict__ pairs, const int32_t* __restrict__ n_match, int P, int F, int K2,
             const double* __restrict__ h, const double* __restrict__ S11, const double* __restrict__ pos2,
             uint8_t* __restrict__ ic, double* __restrict__ z, int32_t* __restrict__ match, int32_t* __restrict__ n_disc) {
  const int p = blockIdx.x;
  const double QNAN = __longlong_as_double(0x7ff8000000000000LL);
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    ic[(size_t)p * F + f] = 0;
    match[(size_t)p * F + f] = -1;
    z[((size_t)p * F + f) * 2] = QNAN;
    z[((size_t)p * F + f) * 2 + 1] = QNAN;
  }
  __syncthreads();
  const int n = n_match[p];
  int disc = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int k1 = pairs[((size_t)p * F + i) * 2], k2 = pairs[((size_t)p * F + i) * 2 + 1];
    const double S = S11[(size_t)p * F + i];
    const double radius = (S != S) ? 40.0 : ceil(3.0 * sqrt(S));
    const double dx = pos2[((size_t)p * K2 + k2) * 2] - h[((size_t)p * F + k1) * 2];
    const double dy = pos2[((size_t)p * K2 + k2) * 2 + 1] - h[((size_t)p * F + k1) * 2 + 1];
    const double dist = sqrt(dx * dx + dy * dy);
    if (dist <= radius) {
      ic[(size_t)p * F + k1] = 1;
      match[(size_t)p * F + k1] = k2;
      z[((size_t)p * F + k1) * 2] = pos2[((size_t)p * K2 + k2) * 2];
      z[((size_t)p * F + k1) * 2 + 1] = pos2[((size_t)p * K2 + k2) * 2 + 1];
    } else {
      ++disc;
    }
  }
  __shared__ int s_disc;
  if (threadIdx.x == 0) s_disc = 0;
  __syncthreads();
  if (disc) atomicAdd(&s_disc, disc);
  __syncthreads();
  if (threadIdx.x == 0) n_disc[p] = s_disc;
}

int launch_match_gate(pre3_ctx* ctx, const int32_t* dpairs, const int32_t* dn_match, int P, int F, int K2, const double* dh,
                      const double* dS11, const double* dpos2, uint8_t* dic, double* dz, int32_t* dmatch,
                      int32_t* dn_disc) {
  Span span__(ctx, T_COMPACT);
  if (P <= 0) return PRE3_OK;
  k_match_gate<<<P, 256, 0, ctx->stream>>>(dpairs, dn_match, P, F, K2, dh, dS11, dpos2, dic, dz, dmatch, dn_disc);
  count_launch(ctx);
  PRE3_CUDA(cudaGetLastError());
  return PRE3_OK;
}

}  // namespace pre3
