// C-ABI entry points for the natural / expectation parameter transforms
// (include/markovflow_b200.h; reference markovflow/ssm_gaussian_transformations.py).
#include "dispatch.cuh"
#include "mid_api.h"
#include "nat_kernels.cuh"
#include "ssm_sweep_api.h"

using namespace mf;

extern "C" {

int mf_nat_to_ssm(int dtype, const void* theta_lin, const void* theta_diag, const void* theta_sub,
                  void* out_a, void* out_offsets, void* out_chols, int32_t* info, int64_t B,
                  int64_t T, int64_t D, int smoothing, void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!theta_lin || !theta_diag || !out_offsets || !out_chols) return MF_ERR_BAD_ARG;
  if (T > 1 && (!theta_sub || !out_a)) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D))
    return mid_nat_to_ssm(dtype, theta_lin, theta_diag, theta_sub, out_a, out_offsets, out_chols, info, B, T, D,
                          smoothing, s);
  if (smoothing && D <= kSsmSweepMaxD && T > 1 && tuning(4) != 1) {
    const int rc = nat_sweep_to_ssm(dtype, D, theta_lin, theta_diag, theta_sub, out_a, out_offsets,
                                    out_chols, info, B, T, s);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    if (smoothing) {
      nat_to_ssm_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>(
          (const Tp*)theta_lin, (const Tp*)theta_diag, (const Tp*)theta_sub, (Tp*)out_a,
          (Tp*)out_offsets, (Tp*)out_chols, info, B, T);
    } else {
      if (info && cudaMemsetAsync(info, 0, sizeof(int32_t) * B, s) != cudaSuccess)
        return check_launch();
      nat_to_ssm_no_smoothing_kernel<Tp, kD><<<grid_for(B * T, 128), 128, 0, s>>>(
          (const Tp*)theta_lin, (const Tp*)theta_diag, (const Tp*)theta_sub, (Tp*)out_a,
          (Tp*)out_offsets, (Tp*)out_chols, info, B, T);
    }
    return check_launch();
  });
}

int mf_ssm_to_naturals(int dtype, const void* mu0, const void* chol_p0, const void* a,
                       const void* b, const void* chol_q, void* theta_lin, void* theta_diag,
                       void* theta_sub, int64_t B, int64_t T, int64_t D, int smoothing,
                       void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!mu0 || !chol_p0 || !theta_lin || !theta_diag) return MF_ERR_BAD_ARG;
  if (T > 1 && (!a || !b || !chol_q || !theta_sub)) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D))
    return mid_ssm_to_naturals(dtype, mu0, chol_p0, a, b, chol_q, theta_lin, theta_diag, theta_sub, B, T, D,
                               smoothing, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    ssm_to_naturals_kernel<Tp, kD><<<grid_for(B * T, 128), 128, 0, s>>>(
        (const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b, (const Tp*)chol_q,
        (Tp*)theta_lin, (Tp*)theta_diag, (Tp*)theta_sub, B, T, smoothing);
    return check_launch();
  });
}

int mf_ssm_to_expectations(int dtype, const void* mu0, const void* chol_p0, const void* a,
                           const void* b, const void* chol_q, void* eta_lin, void* eta_diag,
                           void* eta_sub, int64_t B, int64_t T, int64_t D, void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!mu0 || !chol_p0 || !eta_lin || !eta_diag) return MF_ERR_BAD_ARG;
  if (T > 1 && (!a || !b || !chol_q || !eta_sub)) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (mid_dim(D)) return mid_ssm_moments(dtype, 1, mu0, chol_p0, a, b, chol_q, eta_lin, eta_diag, eta_sub, B, T, D, s);
  if (D <= kSsmSweepMaxD && T > 1 && tuning(4) != 1) {
    const int rc = ssm_sweep_moments(dtype, D, 1, mu0, chol_p0, a, b, chol_q, eta_lin, eta_diag,
                                     eta_sub, B, T, s);
    if (rc != MF_ERR_UNSUPPORTED) return rc;
  }
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    ssm_to_expectations_kernel<Tp, kD><<<grid_for(B, 32), 32, 0, s>>>(
        (const Tp*)mu0, (const Tp*)chol_p0, (const Tp*)a, (const Tp*)b, (const Tp*)chol_q,
        (Tp*)eta_lin, (Tp*)eta_diag, (Tp*)eta_sub, B, T);
    return check_launch();
  });
}

int mf_expectations_to_ssm(int dtype, const void* eta_lin, const void* eta_diag,
                           const void* eta_sub, void* out_a, void* out_offsets, void* out_chols,
                           int32_t* info, int64_t B, int64_t T, int64_t D, void* stream) {
  if (B < 0 || T < 1 || D < 1) return MF_ERR_BAD_ARG;
  if (B == 0) return MF_OK;
  if (!eta_lin || !eta_diag || !out_offsets || !out_chols) return MF_ERR_BAD_ARG;
  if (T > 1 && (!eta_sub || !out_a)) return MF_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (info && cudaMemsetAsync(info, 0, sizeof(int32_t) * B, s) != cudaSuccess) return check_launch();
  if (mid_dim(D))
    return mid_expectations_to_ssm(dtype, eta_lin, eta_diag, eta_sub, out_a, out_offsets, out_chols, info, B, T, D, s);
  return dispatch_small(dtype, D, [&](auto tt, auto dd) {
    using Tp = typename decltype(tt)::type;
    constexpr int kD = decltype(dd)::value;
    expectations_to_ssm_kernel<Tp, kD><<<grid_for(B * T, 128), 128, 0, s>>>(
        (const Tp*)eta_lin, (const Tp*)eta_diag, (const Tp*)eta_sub, (Tp*)out_a, (Tp*)out_offsets,
        (Tp*)out_chols, info, B, T);
    return check_launch();
  });
}

}  // extern "C"
