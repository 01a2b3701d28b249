// match_tc.cu -- placeholder until the tcgen05 proposal kernel lands (next commit).
#include "match.cuh"
namespace pre3 {
bool match_tc_supported(int, int, int, int) { return false; }
int launch_match_tc(pre3_ctx* ctx, const void*, const void*, int, int, int, int, int, const int32_t*, const int32_t*,
                    float, MatchRow*) {
  return fail(ctx, PRE3_ERR_ARG, "tensor-core matcher not built");
}
}  // namespace pre3
