// match.cuh -- internal (C++) interface of the descriptor-matching kernels, used by capi.cu.
#pragma once
#include "common.cuh"

namespace pre3 {

// Per L1 descriptor (row): result of the nearest / second-nearest search before compaction.
struct MatchRow {
  double best;      // best squared distance in the class's accumulation type, widened
  int32_t bestk;    // 0-based index into L2, -1 if none
  int32_t accept;   // ratio test passed (siftmatch.c:122-123)
};

size_t match_workspace_bytes(int cls, int P, int K1, int K2, int ND);

// Exact brute force in the reference's arithmetic (siftmatch.c:98-116); any ND, any class.
int launch_match_exact(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int P, int K1, int K2, int ND,
                       const int32_t* dk1, const int32_t* dk2, float thresh, MatchRow* drows);

// Exact recomputation of listed rows (entries p*K1 + k1), one warp per row.
int launch_match_rows_exact(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int K1, int K2, int ND,
                            const int32_t* dk2, float thresh, const int32_t* drow_list, const int32_t* drow_list_n,
                            int list_cap, MatchRow* drows);

// Tensor-core proposal (tcgen05 split-fp16 GEMM with fused top-k) + exact rescore; ND == 128,
// classes double / single.  Rows whose candidates cannot be certified are recomputed exactly.
bool match_tc_supported(int cls, int K1, int K2, int ND);
size_t match_tc_workspace_bytes(int P, int K1, int K2);
// need_score == 0: MatchRow::best may be approximate for rows whose acceptance the brackets decide.
int launch_match_tc(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int P, int K1, int K2, int ND,
                    const int32_t* dk1, const int32_t* dk2, float thresh, int need_score, MatchRow* drows);

// Exact integer tensor-core matcher (tcgen05 kind::i8, s32 accumulators) for class int8 / uint8, ND == 128: writes the
// final rows, nothing to rescore.  Descriptor pointers must be 16-byte aligned.
bool match_i8_supported(int cls, int K1, int K2, int ND);
size_t match_i8_workspace_bytes(int P, int K1, int K2);
int launch_match_i8(pre3_ctx* ctx, const void* dL1, const void* dL2, int cls, int P, int K1, int K2, int ND,
                    const int32_t* dk1, const int32_t* dk2, float thresh, MatchRow* drows);

// rows -> compact (k1,k2) list in k1 order (siftmatch.c:238-246) and, optionally, the gathered
// correspondences Ya = xyz1(:,k1), Yb = xyz2(:,k2) (SIFT_match_save.m:53).
int launch_match_compact(pre3_ctx* ctx, const MatchRow* drows, int P, int K1, const int32_t* dk1,
                         int32_t* dpairs, double* dscore, int32_t* dn_out, const double* dxyz1,
                         const double* dxyz2, int K2, double* dYa, double* dYb);

// search-region gate of matching_sift_based.m:118-135 over the compacted match lists (P problems, F rows each)
int launch_match_gate(pre3_ctx* ctx, const int32_t* dpairs, const int32_t* dn_match, int P, int F, int K2, const double* dh,
                      const double* dS11, const double* dpos2, uint8_t* dic, double* dz, int32_t* dmatch,
                      int32_t* dn_disc);

}  // namespace pre3
