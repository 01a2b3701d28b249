// cov.cuh -- covariance of the RANSAC pose (M/cov_est_RANSAC_deriv.m), internal interface.
#pragma once
#include "common.cuh"

namespace pre3 {

size_t cov_workspace_bytes(int P, int Nmax);
// dRT: per pair 9 doubles R (column-major) followed by 3 doubles T, `rt_stride` doubles apart.
int launch_cov_est(pre3_ctx* ctx, const double* dYa, const double* dYb, const int32_t* dn_corr, const uint8_t* dmasks,
                   int P, int Nmax, const double* dRT, int rt_stride, pre3_cov_result* dout);

}  // namespace pre3
