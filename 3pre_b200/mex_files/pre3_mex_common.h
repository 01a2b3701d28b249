/* Shared by the MEX gateways: the lazily created libpre3 context (one per MATLAB process,
 * torn down by mexAtExit -- the lifetime MATLAB-Coder MEX files use,
 * M/mex_files/CorePar_Ver1/codegen/mex/corrcoef_partitioned/corrcoef_partitioned_mex.c:39-57)
 * and the error bridge: C-ABI error codes become mexErrMsgIdAndTxt("pre3:...") which does not
 * return (M/sift/siftmatch.c:155-190 relies on the same behaviour). */
#ifndef PRE3_MEX_COMMON_H
#define PRE3_MEX_COMMON_H
#include <string.h>

#include "mex.h"
#include "pre3.h"

static pre3_ctx *g_pre3 = NULL;

static void pre3_mex_atexit(void) {
  if (g_pre3) pre3_destroy(g_pre3);
  g_pre3 = NULL;
}

static pre3_ctx *pre3_mex_ctx(void) {
  if (!g_pre3) {
    int rc = pre3_create(&g_pre3, -1);
    if (rc != PRE3_OK) {
      static char msg[512];
      strncpy(msg, g_pre3 ? pre3_last_error(g_pre3) : "pre3_create failed", sizeof msg - 1);
      if (g_pre3) pre3_destroy(g_pre3);
      g_pre3 = NULL;
      mexErrMsgIdAndTxt("pre3:cuda", msg);
    }
    mexAtExit(pre3_mex_atexit);
  }
  return g_pre3;
}

static void pre3_mex_check(int rc) {
  if (rc == PRE3_OK) return;
  const char *id = rc == PRE3_ERR_CUDA ? "pre3:cuda" : rc == PRE3_ERR_ALLOC ? "pre3:alloc" : "pre3:arg";
  mexErrMsgIdAndTxt(id, pre3_last_error(g_pre3));
}
#endif
